"""GPU tests at BASELINE.json's full C3 size (64 targets x 200 patches x 17 x 64 x 120, f32 library = 13.4 GB in HBM,
f64 = 26.8 GB): EVERY per-dataset logpt of 1024 chains against the oracle (fanned out over the host cores, library
blocks regenerated on the CPU) for both storage modes, near-MAP populations at two noise levels, and size-independent
properties.  Tolerances: f32 storage rtol 1e-5 (the north star's; the reference's own stack tolerance is 5e-6,
test/test_ffi_gfstacking.py:49-58), f64 storage rtol 1e-10."""
import json
import os

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


@pytest.fixture(scope="module")
def c3_f64(c3):
    """Strict mode: the same problem with the library stored in float64 (26.8 GB), every operation in f64."""
    prob, _, _, torch, dev = c3
    from beat_b200.devlib import fill_library_on_device
    from beat_b200.engine import BatchedFFILogLike
    ev = BatchedFFILogLike.from_problem(prob, device=0, store_dtype="float64", upload_libraries=False)
    fill_library_on_device(ev, prob, torch, dev, "f64")
    yield ev
    ev.close()


@pytest.fixture(scope="module")
def host_pool():
    from oracle import parallel_check as PC
    with PC.pool() as ex:
        yield ex


N_ALL = 1024


@pytest.fixture(scope="module")
def c3_oracle(c3, host_pool):
    """Oracle logpts [1024, 64] of 1024 prior draws at C3 (float64 CPU path, one chain and one target at a time)."""
    from oracle import parallel_check as PC
    prob = c3[0]
    Q = synthetic.draw_chains(prob, N_ALL, seed=20261017)
    return Q, PC.full_size_logpts(prob, Q, host_pool)


def _report(name, **kw):
    """Measured parity margins, kept next to the bench output (gpurun_out/ travels back from the GPU box)."""
    d = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        path = os.path.join(d, "fullsize_parity.json")
        cur = json.load(open(path)) if os.path.exists(path) else {}
        cur[name] = kw
        json.dump(cur, open(path, "w"), indent=1)
    except Exception:
        pass


def test_c3_all_chains_all_logpts_f32(c3, c3_oracle):
    """SURVEY 8d parity gate at full size: llk rtol 1e-5 on ALL chains, per-dataset logpts too (f32 library storage vs the
    float64 CPU path on the float64 library)."""
    prob, ev, views, torch, dev = c3
    Q, ref = c3_oracle
    logpts, like = ev(Q)
    assert logpts.shape == ref.shape == (N_ALL, 64) and np.isfinite(ref).all()
    err = np.abs(logpts / ref - 1.0)
    _report("c3_prior_draws_f32", chains=N_ALL, logpts=int(ref.size), max_rel_err=float(err.max()), max_rel_err_like=float(np.abs(like / ref.sum(axis=1) - 1).max()))
    np.testing.assert_allclose(logpts, ref, rtol=1e-5)                   # north-star tolerance
    np.testing.assert_allclose(like, ref.sum(axis=1), rtol=1e-5)
    # the device-pointer entry gives the same numbers as the host-pointer entry
    lp_dev, _ = ev.eval_device(torch.from_numpy(Q).to(dev))
    torch.cuda.synchronize()
    assert np.array_equal(lp_dev.cpu().numpy(), logpts)


def test_c3_all_chains_all_logpts_f64(c3, c3_f64, c3_oracle):
    """Strict mode at full size: f64 library, rtol 1e-10 on all chains and datasets (torch's exp in the device fill and
    numpy's in the CPU recipe differ by <= 1 ulp per library value: ~1e-13 on a logpt)."""
    Q, ref = c3_oracle
    logpts, like = c3_f64(Q)
    err = np.abs(logpts / ref - 1.0)
    _report("c3_prior_draws_f64", chains=N_ALL, logpts=int(ref.size), max_rel_err=float(err.max()))
    np.testing.assert_allclose(logpts, ref, rtol=1e-10)
    np.testing.assert_allclose(like, ref.sum(axis=1), rtol=1e-10)
    # identical rupture onset times in both modes, bit-exact vs the sequential C restatement
    st = c3_f64.starttimes(N_ALL)
    for c in (0, 511, 1023):
        pt = synthetic.split_point(c3[0], Q[c])
        hr, hc = O.fault_locations2idxs(pt["nucleation_dip"][0], pt["nucleation_strike"][0], 2.0, 2.0)
        assert np.array_equal(st[c], O.fast_sweep(1.0 / pt["velocities"], 2.0, hr, hc, 10, 20, impl="port") + pt["time"][0])


def test_c3_all_chains_nearest_neighbor(c3, host_pool):
    """The other interpolation mode of the reference (`nearest_neighbor`: one tap per patch, rint index mapping,
    ffi/base.py:649-660) at full C3 size: every logpt of 512 chains, f32 library, against the oracle."""
    import copy
    from oracle import parallel_check as PC
    from beat_b200.devlib import fill_library_on_device
    from beat_b200.engine import BatchedFFILogLike
    prob, _, _, torch, dev = c3
    prob_nn = copy.copy(prob)
    prob_nn["wavemaps"] = [dict(prob["wavemaps"][0], interpolation="nearest_neighbor")]
    ev = BatchedFFILogLike.from_problem(prob_nn, device=0, store_dtype="float32", upload_libraries=False)
    try:
        fill_library_on_device(ev, prob_nn, torch, dev, "f32")
        Q = synthetic.draw_chains(prob_nn, 512, seed=515)
        ref = PC.full_size_logpts(prob_nn, Q, host_pool)
        logpts, like = ev(Q)
        err = np.abs(logpts / ref - 1.0)
        _report("c3_prior_draws_f32_nearest_neighbor", chains=512, logpts=int(ref.size), max_rel_err=float(err.max()))
        np.testing.assert_allclose(logpts, ref, rtol=1e-5)
        np.testing.assert_allclose(like, ref.sum(axis=1), rtol=1e-5)
    finally:
        ev.close()


@pytest.mark.parametrize("noise_frac", [0.05, 0.01])
def test_c3_near_map_population(c3, c3_f64, host_pool, noise_frac):
    """Where the likelihood is most sensitive to synthetics errors: data = synth(q_true) + noise, chains in a small
    neighbourhood of q_true, so that the residual is the noise itself.  The gate is the north star's: the chain's
    log-likelihood (`like`) within rtol 1e-5 for f32 library storage, 1e-10 for f64.  Per-dataset logpts are held to the same
    tolerances relative to the size of the terms they are made of (log|C|, M(2h + ln 2pi) and the quadratic form): at 1 %
    noise log|C| is strongly negative and some logpts cancel to nearly zero, where a plain relative error says nothing
    about the arithmetic (the strict f64 path shows ~1e-10 there, too).  The measured margins at both noise levels go to
    gpurun_out/fullsize_parity.json."""
    from oracle import parallel_check as PC
    from beat_b200.covariance import Covariance, exponential_data_covariance
    prob, ev32, views, torch, dev = c3
    wm = prob["wavemaps"][0]
    nt, ns = wm["nt"], wm["ns"]
    rng = np.random.default_rng(int(noise_frac * 1e4))
    q_true = synthetic.draw_chains(prob, 1, seed=777)[0]
    clean = c3_f64.get_synthetics(q_true)                                  # [nt, ns], the f64-library forward model
    data, U, lpd = np.empty((nt, ns)), np.empty((nt, ns, ns)), np.empty(nt)
    base = exponential_data_covariance(ns, prob["dt"], 2.0)
    for t in range(nt):
        Ct = base * (noise_frac * np.abs(clean[t]).max()) ** 2
        cov = Covariance(data=Ct)
        U[t], lpd[t] = cov.chol_inverse, cov.log_pdet
        data[t] = clean[t] + np.linalg.cholesky(Ct).dot(rng.standard_normal(ns))
    B = 256
    lo = np.concatenate([prob["priors"][n][0] for n, _ in prob["var_order"]])
    hi = np.concatenate([prob["priors"][n][1] for n, _ in prob["var_order"]])
    Q = np.clip(q_true + rng.standard_normal((B, q_true.size)) * (hi - lo) * 0.002, lo, hi)
    Q[0] = q_true
    oh = prob["offsets"]["hypers"]
    Q[:, oh] = np.clip(rng.normal(0.0, 0.05, B), 0.0, 4.0)                 # near the true noise scale (h = 0)
    try:
        for ev in (ev32, c3_f64):
            ev.ctx.upload_data(ev.wmap_ids[0], data)
            ev.update_weights(0, U, lpd)
        ref = PC.full_size_logpts(prob, Q, host_pool, data=data, U=U, slog_pdet=lpd)
        l32, _ = ev32(Q)
        l64, _ = c3_f64(Q)
    finally:
        for ev in (ev32, c3_f64):
            ev.ctx.upload_data(ev.wmap_ids[0], wm["data"])
            ev.update_weights(0, wm["U"], wm["slog_pdet"])
    # the population really sits where residual ~ noise: chi^2 per sample of the true point is ~1
    quad0 = -2.0 * ref[0] - lpd - ns * (2.0 * Q[0, oh] + np.log(2 * np.pi))
    assert 0.5 < np.median(quad0 * np.exp(2.0 * Q[0, oh])) / ns < 1.5
    h = Q[:, oh][:, None]
    scale = 0.5 * (np.abs(lpd)[None, :] + ns * np.abs(2.0 * h + np.log(2 * np.pi)) + np.abs(-2.0 * ref - lpd[None, :] - ns * (2.0 * h + np.log(2 * np.pi))))
    e32, e64 = np.abs(l32 - ref) / scale, np.abs(l64 - ref) / scale
    like_ref = ref.sum(axis=1)
    k32, k64 = np.abs(l32.sum(axis=1) / like_ref - 1.0), np.abs(l64.sum(axis=1) / like_ref - 1.0)
    plain32 = np.abs(l32 / ref - 1.0)
    _report("c3_near_map_noise_%g" % noise_frac, chains=B, like_max_rel_err_f32=float(k32.max()), like_max_rel_err_f64=float(k64.max()),
            logpts_max_err_over_term_size_f32=float(e32.max()), logpts_max_err_over_term_size_f64=float(e64.max()),
            logpts_plain_rel_err_f32_max=float(plain32.max()), logpts_plain_rel_err_f32_median=float(np.median(plain32)),
            logpts_plain_rel_err_f32_frac_above_1e_5=float((plain32 > 1e-5).mean()), min_abs_logpt=float(np.abs(ref).min()),
            median_abs_logpt=float(np.median(np.abs(ref))))
    assert k64.max() <= 1e-10 and e64.max() <= 1e-10
    assert k32.max() <= 1e-5 and e32.max() <= 1e-5


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
