"""L2 blocking of the stacking pass is derived from the library shape and the device's L2 size, not hard-coded:
libraries with larger per-patch blocks (more start times, longer traces, a third slip component) get smaller patch
chunks, results stay identical to the oracle whatever the chunking, and the row-index range is checked at upload."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from beat_b200 import synthetic  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402

SHAPES = {
    # per-patch library block (nvar * ndur * nst * row bytes): C3's rows; 4x larger block (nst 128, ns 240); 3 slip vars
    "c3_rows": dict(nt=3, subfaults=((6, 10, 2.0),), ns=120, ndur=17, nst=64),
    "big_block": dict(nt=2, subfaults=((6, 10, 2.0),), ns=240, ndur=6, nst=128),
    "three_vars_ragged_ns": dict(nt=3, subfaults=((5, 9, 2.0),), ns=102, ndur=5, nst=64, slip_vars=("uparr", "uperp", "utens")),
}


def _oracle(prob, Q):
    return np.array([O.ffi_seismic_eval(prob, synthetic.split_point(prob, q), impl="port") for q in Q])


@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("l2_frac", [None, "0.004", "0.0005"])
def test_parity_for_every_chunking(monkeypatch, shape, l2_frac):
    from beat_b200.engine import BatchedFFILogLike
    if l2_frac is None:
        monkeypatch.delenv("BEATGPU_L2_FRAC", raising=False)
    else:
        monkeypatch.setenv("BEATGPU_L2_FRAC", l2_frac)       # a small budget stands in for a library with a huge per-patch block
    monkeypatch.delenv("BEATGPU_CHUNK", raising=False)
    prob = synthetic.make_problem(interpolation="multilinear", seed=5, **SHAPES[shape])
    Q = synthetic.draw_chains(prob, 48, seed=6)
    ref = _oracle(prob, Q)
    seen = {}
    for store, rtol in (("float64", 1e-10), ("float32", 1e-5)):
        ev = BatchedFFILogLike.from_problem(prob, device=0, store_dtype=store)
        blk = ev.ctx.stack_blocking(ev.wmap_ids[0], len(prob["slip_vars"]))
        logpts, like = ev(Q)
        ev.close()
        np.testing.assert_allclose(logpts, ref, rtol=rtol)
        np.testing.assert_allclose(like, ref.sum(axis=1), rtol=rtol)
        npatch = prob["npatches"]
        assert blk["n_chunks"] == -(-npatch // blk["chunk_patches"]) and 1 <= blk["chunk_patches"] <= 64
        frac = float(l2_frac) if l2_frac else 0.6
        per_patch = blk["chunk_bytes"] // blk["chunk_patches"]
        # the chunk honours the L2 budget unless the floor (<= 24 chunks, bounded scratch) or one patch alone exceeds it
        floor = min(64, -(-npatch // 24))
        assert blk["chunk_bytes"] <= max(frac * blk["l2_bytes"], floor * per_patch, per_patch) + per_patch
        assert blk["l2_bytes"] > 32 << 20
        seen[store] = blk
    # f64 rows are twice as long: never more patches per chunk than f32
    assert seen["float64"]["chunk_patches"] <= seen["float32"]["chunk_patches"]


def test_forced_chunk_and_fused_kernel_agree(monkeypatch):
    """BEATGPU_CHUNK overrides the derivation; every chunking and the single-kernel variant give the same logpts
    (partials are summed in fixed chunk order, so a given chunking is deterministic; across chunkings only the rounding
    of the f64 partial sums differs)."""
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(interpolation="multilinear", seed=9, **SHAPES["c3_rows"])
    Q = synthetic.draw_chains(prob, 32, seed=10)
    outs = {}
    for name, env in (("derived", {}), ("chunk7", {"BEATGPU_CHUNK": "7"}), ("chunk32", {"BEATGPU_CHUNK": "32"}),
                      ("fused", {"BEATGPU_STACK_MODE": "fused"})):
        for k in ("BEATGPU_CHUNK", "BEATGPU_STACK_MODE", "BEATGPU_L2_FRAC"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        ev = BatchedFFILogLike.from_problem(prob, device=0, store_dtype="float64")
        if name == "chunk7":
            assert ev.ctx.stack_blocking(ev.wmap_ids[0], 2)["chunk_patches"] <= 7
        outs[name] = ev(Q)[0]
        again = ev(Q)[0]
        assert np.array_equal(outs[name], again)              # deterministic for a given chunking
        ev.close()
    for name in ("chunk7", "chunk32", "fused"):
        np.testing.assert_allclose(outs[name], outs["derived"], rtol=1e-12)


def test_row_index_range_checked_at_upload():
    """PatchPlan rows are 32-bit: a library with more than 2^31 rows is refused when it is declared, before any byte
    is allocated (nothing can overflow silently inside the kernels)."""
    from beat_b200.lib import Context, F32
    c = Context(0)
    c.set_fault([100], [100], [1.0])
    wid = c.add_wavemap(64, 16, "multilinear", None, np.zeros(64, np.int32), np.full(64, 16, np.int32))
    with pytest.raises(ValueError, match="2\\^31 rows"):
        c.alloc_gflib(wid, 0, F32, (64, 10000, 64, 64, 16), 0.5, 0.25, -5.0, 0.5)      # 2.6e9 rows
    c.close()
