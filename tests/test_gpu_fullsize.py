"""GPU tests at BASELINE.json's full C3 size (64 targets x 200 patches x 17 x 64 x 120, f32 library = 13.4 GB in HBM):
spot checks against the oracle on library blocks regenerated on the CPU, and size-independent properties."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from beat_b200 import synthetic  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402

C3 = dict(nt=64, subfaults=((10, 20, 2.0),), ns=120, ndur=17, nst=64)


@pytest.fixture(scope="module")
def c3():
    import torch
    from beat_b200.devlib import fill_library_on_device
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(interpolation="multilinear", seed=1234, build_library=False, **C3)
    ev = BatchedFFILogLike.from_problem(prob, device=0, store_dtype="float32", upload_libraries=False)
    dev = torch.device("cuda", 0)
    views = fill_library_on_device(ev, prob, torch, dev, "f32")
    yield prob, ev, views, torch, dev
    ev.close()


def _sub_problem(prob, t):
    """One-target problem with the library block regenerated on the CPU from the same recipe (f64)."""
    wm = prob["wavemaps"][0]
    G = {v: synthetic.library_block(wm["A"][v][t:t + 1], wm["k0"][v][t:t + 1], wm["ndur"], wm["nst"], wm["ns"], wm["st_step"], prob["dt"])
         for v in prob["slip_vars"]}
    wm1 = dict(wm, nt=1, G=G, data=wm["data"][t:t + 1], U=wm["U"][t:t + 1], slog_pdet=wm["slog_pdet"][t:t + 1],
               nsamples=wm["nsamples"][t:t + 1], hyper_idx=wm["hyper_idx"][t:t + 1])
    return dict(prob, wavemaps=[wm1])


def test_c3_spot_check_vs_oracle(c3):
    prob, ev, views, torch, dev = c3
    B = 512
    Q = synthetic.draw_chains(prob, B, seed=4321)
    logpts, like = ev(Q)
    assert logpts.shape == (B, 64) and np.isfinite(logpts).all()
    np.testing.assert_allclose(like, logpts.sum(axis=1), rtol=1e-13)
    for t in (0, 31, 63):
        sub = _sub_problem(prob, t)
        # the device library (f32) equals the CPU recipe to f32 rounding (torch exp vs numpy exp differ by <= 1 ulp f64)
        blk = views[0][t, 5].cpu().numpy()
        ref = sub["wavemaps"][0]["G"][prob["slip_vars"][0]][0, 5]
        np.testing.assert_allclose(blk, ref, rtol=2e-7, atol=1e-7 * np.abs(ref).max())
        for c in (0, 17, 300, 511):
            lp = O.ffi_seismic_eval(sub, synthetic.split_point(prob, Q[c]), impl="port")[0]
            assert abs(logpts[c, t] - lp) <= 1e-5 * abs(lp), (c, t, logpts[c, t], lp)     # north-star tolerance
    # start times of the fused path: bit-exact vs the sequential C restatement
    st = ev.starttimes(B)
    for c in (0, 17, 300):
        pt = synthetic.split_point(prob, Q[c])
        hr, hc = O.fault_locations2idxs(pt["nucleation_dip"][0], pt["nucleation_strike"][0], 2.0, 2.0)
        t0 = O.fast_sweep(1.0 / pt["velocities"], 2.0, hr, hc, 10, 20, impl="port") + pt["time"][0]
        assert np.array_equal(st[c], t0)


def test_c3_properties(c3):
    prob, ev, views, torch, dev = c3
    B = 256
    Q = synthetic.draw_chains(prob, B, seed=99)
    base, like = ev(Q)
    # order independence / determinism: a permuted batch gives bit-identical per-chain results
    perm = np.random.default_rng(0).permutation(B)
    got, _ = ev(Q[perm])
    assert np.array_equal(got, base[perm])
    # a chain evaluated alone equals the same chain inside a batch
    one, _ = ev(Q[7:8])
    assert np.array_equal(one[0], base[7])
    # duplicated chains -> duplicated results; device-resident entry == host entry
    dup, _ = ev(np.repeat(Q[:4], 3, axis=0))
    assert np.array_equal(dup[0::3], base[:4]) and np.array_equal(dup[1::3], base[:4])
    lp_dev, like_dev = ev.eval_device(torch.from_numpy(Q).to(dev))
    torch.cuda.synchronize()
    assert np.array_equal(lp_dev.cpu().numpy(), base)
    # hyper-parameter identity: logpt(h) - logpt(0) = -M h - (exp(-2h) - 1) quad/2 ; check via two evaluations
    Qh = Q.copy()
    oh = prob["offsets"]["hypers"]
    Qh[:, oh] = 0.0
    l0, _ = ev(Qh)
    Qh[:, oh] = 1.0
    l1, _ = ev(Qh)
    wm = prob["wavemaps"][0]
    M = wm["nsamples"].astype(float)
    quad = -2.0 * l0 - wm["slog_pdet"] - M * np.log(2 * np.pi)
    pred = -0.5 * (wm["slog_pdet"] + M * (2.0 + np.log(2 * np.pi)) + np.exp(-2.0) * quad)
    np.testing.assert_allclose(l1, pred, rtol=1e-12)


def test_c3_stack_linearity(c3):
    """stack_all is linear in the slips: doubling is exact in binary floating point, additivity to rounding."""
    prob, ev, views, torch, dev = c3
    wm = prob["wavemaps"][0]
    rng = np.random.default_rng(5)
    B, nt, npatch, ns = 4, 64, 200, 120
    d = rng.uniform(0.6, 4.4, (B, npatch))
    st = rng.uniform(-4.0, 25.0, (B, nt, npatch))
    u1 = rng.uniform(0, 3, (2, B, npatch))
    u2 = rng.uniform(0, 3, (2, B, npatch))
    wid = ev.wmap_ids[0]
    s1 = ev.ctx.stack_batch(wid, d, st, u1, nt, ns)
    s2 = ev.ctx.stack_batch(wid, d, st, u2, nt, ns)
    s12 = ev.ctx.stack_batch(wid, d, st, u1 + u2, nt, ns)
    sd = ev.ctx.stack_batch(wid, d, st, 2.0 * u1, nt, ns)
    assert np.array_equal(sd, 2.0 * s1)
    np.testing.assert_allclose(s12, s1 + s2, rtol=0, atol=2e-6 * np.abs(s12).max())
    # zero slip -> exactly zero synthetics
    assert not ev.ctx.stack_batch(wid, d, st, np.zeros_like(u1), nt, ns).any()
