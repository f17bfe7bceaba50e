"""The contractual fast-sweeping gate at scale, against the REFERENCE'S OWN BINARY: rupture
start-time library indices of the CUDA sweep identical to those of fast_sweep_ext.c (compiled unmodified into
oracle/_ref; /root/reference/beat/fast_sweeping/fast_sweep_ext.c:120-206, its own gate test/test_fastsweep.py:125-133)
on >= 1e6 random chains -- 10x20 and 10x15 patch grids, rough and smooth media -- for nearest-neighbour (rint) and
multilinear (ceil) index mapping, start times within 8 ulp (the reference's pow(x, 0.5) vs correctly rounded sqrt: 1 ulp per
update, a handful accumulated along a ray -- 5 ulp is the largest seen in 1e6 chains),
bit-identical to the sequential C restatement including the outer iteration counts."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import ffi_oracle as O  # noqa: E402
from oracle import parallel_check as PC  # noqa: E402

N_PER_CASE = 262144            # x 4 cases = 1 048 576 chains
CHUNK = 4096


def _media(kind, rng, B, nd, ns):
    if kind == "rough":                                  # iid velocities over the prior range
        return rng.uniform(2.2, 4.5, (B, nd * ns))
    if kind == "very_rough":                             # far outside the prior: more outer iterations
        return rng.uniform(0.3, 6.0, (B, nd * ns))
    # smooth: planar gradient in dip and strike plus 2 % perturbation
    r = np.arange(nd, dtype=np.float64)[:, None] / max(1, nd - 1)
    c = np.arange(ns, dtype=np.float64)[None, :] / max(1, ns - 1)
    g = rng.uniform(-1.0, 1.0, (B, 2))
    v0 = rng.uniform(2.6, 4.0, B)
    v = v0[:, None, None] + 0.4 * (g[:, 0, None, None] * r + g[:, 1, None, None] * c)
    v = v * (1.0 + 0.02 * rng.standard_normal((B, nd, ns)))
    return np.clip(v, 2.2, 4.5).reshape(B, nd * ns)


@pytest.fixture(scope="module")
def host_pool():
    with PC.pool() as ex:
        yield ex


@pytest.mark.parametrize("nd,ns,h,kind", [(10, 20, 2.0, "rough"), (10, 20, 2.0, "smooth"), (10, 15, 2.0, "rough"),
                                          (10, 15, 2.0, "very_rough")])
def test_sweep_indices_identical_to_reference_binary(host_pool, nd, ns, h, kind):
    from beat_b200.lib import Context
    ext = O.load_reference_ext()
    if ext is None:
        pytest.skip("oracle/_ref (the reference's compiled fast_sweep_ext) was not built in the dev container")
    rng = np.random.default_rng(nd * 100000 + ns * 100 + len(kind))
    B = N_PER_CASE
    slow = 1.0 / _media(kind, rng, B, nd, ns)
    hr, hc = rng.integers(0, nd, B), rng.integers(0, ns, B)
    time_ofs = rng.uniform(-5.0, 5.0, B)                 # `time` of the subfault (seismic.py:1269)

    ctx = Context(0)
    ctx.set_fault([nd], [ns], [h])
    got = np.empty_like(slow)
    it = np.empty(B, dtype=np.int32)
    step = 65536
    for b0 in range(0, B, step):
        got[b0:b0 + step], it[b0:b0 + step] = ctx.fast_sweep_batch(0, slow[b0:b0 + step], hr[b0:b0 + step], hc[b0:b0 + step],
                                                                   return_iters=True)
    ctx.close()

    tasks = [(slow[b0:b0 + CHUNK], h, hr[b0:b0 + CHUNK], hc[b0:b0 + CHUNK], nd, ns) for b0 in range(0, B, CHUNK)]
    res = list(host_pool.map(PC.sweep_chunk, tasks))
    ref = np.concatenate([r["ref"] for r in res])
    port = np.concatenate([r["port"] for r in res])
    it_port = np.concatenate([r["iters"] for r in res])

    # the reference's own binary, in this process too (so the loaded .so shows up under the pytest process itself)
    for i in (0, B // 2, B - 1):
        assert np.array_equal(ext.fast_sweep(np.ascontiguousarray(slow[i]), h, int(hr[i]), int(hc[i]), nd, ns), ref[i])

    # (1) bit-identical to the sequential C restatement, iteration counts included
    assert np.array_equal(got, port)
    assert np.array_equal(it, it_port)
    # (2) within 8 ulp of the reference binary everywhere
    assert np.all(np.abs(got - ref) <= 8 * np.spacing(np.abs(ref)))
    # (3) THE gate: library indices identical, nearest neighbour and multilinear, on the C3 start-time axis
    t_gpu, t_ref = got + time_ofs[:, None], ref + time_ofs[:, None]
    for interp in ("nearest_neighbor", "multilinear"):
        i_gpu, f_gpu = O.times2idxs(t_gpu, -5.0, 0.5, interp)
        i_ref, f_ref = O.times2idxs(t_ref, -5.0, 0.5, interp)
        assert np.array_equal(i_gpu, i_ref), (interp, int((i_gpu != i_ref).sum()))
        if f_gpu is not None:
            assert np.abs(f_gpu - f_ref).max() < 1e-12
    assert it.min() >= 2 and np.isfinite(got).all()
