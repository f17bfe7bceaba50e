"""GPU parity tests (-m gpu): every CUDA path, called through the C-ABI, against the CPU oracle and the
golden vectors produced by the reference's own code.  Integer/index work is bit-exact; float tolerances are
written next to each assertion (north star: llk rtol 1e-5; the reference's own stack tolerance is 5e-6)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from beat_b200 import synthetic  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402


@pytest.fixture(scope="module")
def ctx():
    from beat_b200.lib import Context
    c = Context(0)
    yield c
    c.close()


def _oracle_logpts(prob, Q, impl="port"):
    return np.array([O.ffi_seismic_eval(prob, synthetic.split_point(prob, q), impl=impl) for q in Q])


# ------------------------------------------------------------------------------------------ fast sweeping
def test_sweep_golden_reference_c(golden):
    """CUDA sweep vs the reference's compiled C on the golden cases: <= 4 ulp (pow(x,.5) vs sqrt), same indices."""
    from beat_b200.lib import Context
    for i in range(int(golden["fs_ncases"])):
        nd, ns, nuc_dip, nuc_strike = (int(v) for v in golden[f"fs{i}_meta"])
        c = Context(0)
        c.set_fault([nd], [ns], [float(golden[f"fs{i}_h"])])
        t = c.fast_sweep_batch(0, golden[f"fs{i}_slow"][None, :], [nuc_dip], [nuc_strike])[0]
        ref = golden[f"fs{i}_t_c"]
        assert np.all(np.abs(t - ref) <= 4 * np.spacing(np.abs(ref))), i
        np.testing.assert_allclose(t, golden[f"fs{i}_t_numpy"], rtol=0, atol=1e-6)   # reference's own gate
        for interp in ("nearest_neighbor", "multilinear"):
            assert np.array_equal(O.times2idxs(t, -5.0, 0.5, interp)[0], O.times2idxs(ref, -5.0, 0.5, interp)[0])
        c.close()


@pytest.mark.parametrize("pack", ["1", "0"])
@pytest.mark.parametrize("nd,ns,h", [(10, 20, 2.0), (10, 15, 2.5), (1, 7, 1.0), (9, 1, 3.0), (24, 40, 1.0), (40, 33, 0.5), (16, 16, 1.0),
                                     (5, 30, 2.0), (3, 3, 4.0)])
def test_sweep_bitexact_vs_port(nd, ns, h, pack, monkeypatch):
    """Anti-diagonal wavefront schedule == sequential Gauss-Seidel, bit for bit, incl. the iteration count -- with one
    chain per warp and with several chains sharing a warp (lane groups of min(nd, ns) lanes: 3 chains on a 10-row fault,
    2 on 16 x 16, 6 on 5 x 30, 10 on 3 x 3, 32 on a single row; chains of a warp converge after different iteration counts)."""
    monkeypatch.setenv("BEATGPU_SWEEP_PACK", pack)
    from beat_b200.lib import Context
    rng = np.random.default_rng(nd * 100 + ns)
    B = 3000
    slow = 1.0 / rng.uniform(2.2, 4.5, (B, nd * ns))
    slow[: B // 10] = 1.0 / rng.uniform(0.3, 6.0, (B // 10, nd * ns))      # rough media -> more outer iterations
    hr, hc = rng.integers(0, nd, B), rng.integers(0, ns, B)
    c = Context(0)
    c.set_fault([nd], [ns], [h])
    got, it = c.fast_sweep_batch(0, slow, hr, hc, return_iters=True)
    ref, it_ref = O.fast_sweep_batch_port(slow, h, hr, hc, nd, ns)
    assert np.array_equal(got, ref)
    assert np.array_equal(it, it_ref)
    c.close()


def test_sweep_packed_subfaults_of_different_shape():
    """Fused path, two subfaults with different grids, enough chains for the packed sweep: the lane groups of one warp
    then relax grids of different shape (different diagonal counts); start times bit-identical to the sequential C."""
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=2, subfaults=((4, 7, 2.0), (6, 5, 1.5)), ns=16, ndur=3, seed=77)
    B = 1200                                         # 2400 (chain, subfault) items >= 12 per SM: packed
    Q = synthetic.draw_chains(prob, B, seed=9)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    logpts, _ = ev(Q)
    st = ev.starttimes(B)
    ev.close()
    for b in list(range(0, B, 97)) + [B - 1]:
        lp, _, t0 = O.ffi_seismic_eval(prob, synthetic.split_point(prob, Q[b]), impl="port", return_synth=True)
        assert np.array_equal(st[b], t0), b
        np.testing.assert_allclose(logpts[b], lp, rtol=1e-10)


def test_sweeper_op_reference_test_case():
    """Mirror of test/test_fastsweep.py:84-112 (`_pytensor_c_wrapper`) with the GPU Op."""
    from beat_b200.ops import Sweeper
    patch_size, nuc_x, nuc_y, n_patch_strike, n_patch_dip = 10.0, 2, 3, 4, 6
    velocities = np.concatenate((np.ones((n_patch_dip, 2)), np.ones((n_patch_dip, 2)) * 3.5), axis=1)
    slownesses = 1.0 / velocities
    cleanup = Sweeper(patch_size, n_patch_dip, n_patch_strike, "cuda")
    out = [[None]]
    cleanup.perform(None, [slownesses.flatten(), nuc_y, nuc_x], out)
    c_i = O.fast_sweep(slownesses.flatten(), patch_size, nuc_y, nuc_x, n_patch_dip, n_patch_strike, impl="port")
    np.testing.assert_allclose(out[0][0], c_i, rtol=0.0, atol=1e-6)
    assert np.array_equal(out[0][0], c_i)
    np.testing.assert_allclose(out[0][0].reshape(6, 4)[3], [20.0, 10.0, 0.0, 2.8571428571428568], atol=1e-13)
    with pytest.raises(NotImplementedError):
        Sweeper(patch_size, n_patch_dip, n_patch_strike, "fortran").perform(None, [slownesses.flatten(), 3, 2], [[None]])
    with pytest.raises(IndexError):
        cleanup.perform(None, [slownesses.flatten(), 6, 2], [[None]])     # nucleation outside the grid


# ------------------------------------------------------------------------------------------ stacking
@pytest.mark.parametrize("name", ["rand", "recipe"])
@pytest.mark.parametrize("tag,interp", [("nn", "nearest_neighbor"), ("ml", "multilinear")])
@pytest.mark.parametrize("store,rtol", [("float64", 1e-12), ("float32", 5e-6)])
def test_stack_all_golden(golden, name, tag, interp, store, rtol):
    """SeismicGFLibrary.stack_all on the GPU vs the reference's own numpy stack_all output."""
    from beat_b200.ops import SeismicGFLibrary
    st_min, st_step, dur_min, dur_step = golden["stack_axes"]
    G = golden[f"stack_{name}_G"]
    gfs = SeismicGFLibrary(G, dur_min, dur_step, st_min, st_step, store_dtype=store)
    d, s, u = golden[f"stack_{name}_durations"], golden[f"stack_{name}_starttimes"], golden[f"stack_{name}_slips"]
    tidx = np.atleast_2d(np.arange(G.shape[0])).T
    got = gfs.stack_all(d, s, u, targetidxs=tidx, interpolation=interp)
    ref = golden[f"stack_{name}_{tag}"]
    np.testing.assert_allclose(got, ref, rtol=rtol, atol=rtol * np.abs(ref).max())
    with pytest.raises(ValueError):
        gfs.stack_all(d, s, u, interpolation=interp)                       # targetidxs mandatory (base.py:630-631)
    with pytest.raises(NotImplementedError):
        gfs.stack_all(d, s, u, targetidxs=tidx, interpolation="cubic")


def test_stack_batch_random_vs_oracle():
    """Batched stacking, two components summed, ragged sizes (ns not a multiple of 4, > one 128-sample window)."""
    from beat_b200.ops import SeismicGFLibrary
    rng = np.random.default_rng(8)
    for (nt, npatch, ndur, nst, ns) in [(3, 7, 3, 6, 10), (2, 300, 2, 5, 8), (2, 5, 3, 4, 150), (1, 1, 2, 2, 1)]:
        G = {"uparr": rng.standard_normal((nt, npatch, ndur, nst, ns)), "uperp": rng.standard_normal((nt, npatch, ndur, nst, ns))}
        axes = dict(dur_min=0.5, dur_step=0.25, st_min=-1.0, st_step=0.5)
        B = 5
        d = rng.uniform(0.5 + 1e-3, 0.5 + (ndur - 1) * 0.25, (B, npatch))
        s = rng.uniform(-1.0 + 1e-3, -1.0 + (nst - 1) * 0.5, (B, nt, npatch))
        u = {k: rng.uniform(0, 3, (B, npatch)) for k in G}
        tidx = np.atleast_2d(np.arange(nt)).T
        for store, rtol in (("float64", 1e-12), ("float32", 5e-6)):
            gfs = SeismicGFLibrary(G, 0.5, 0.25, -1.0, 0.5, store_dtype=store)
            for interp in ("nearest_neighbor", "multilinear"):
                got = gfs.stack_all(d, s, u, targetidxs=tidx, interpolation=interp)
                for b in range(B):
                    ref = sum(O.stack_all(G[k], d[b], s[b], u[k][b], axes["dur_min"], axes["dur_step"], axes["st_min"],
                                          axes["st_step"], interp) for k in G)
                    np.testing.assert_allclose(got[b], ref, rtol=rtol, atol=rtol * np.abs(ref).max())


def test_stack_out_of_library_raises():
    from beat_b200.ops import SeismicGFLibrary
    rng = np.random.default_rng(0)
    G = rng.standard_normal((2, 3, 3, 4, 8))
    gfs = SeismicGFLibrary(G, 0.5, 0.25, 0.0, 0.5, store_dtype="float64")
    tidx = np.atleast_2d(np.arange(2)).T
    d, u = np.full(3, 0.7), np.ones(3)
    with pytest.raises(IndexError):
        gfs.stack_all(d, np.full((2, 3), 5.0), u, targetidxs=tidx, interpolation="multilinear")   # beyond last start time
    with pytest.raises(IndexError):
        gfs.stack_all(d, np.full((2, 3), -0.7), u, targetidxs=tidx, interpolation="nearest_neighbor")
    # exactly on the first grid node: the zero-weight floor tap must not count as a violation
    got = gfs.stack_all(np.full(3, 0.5), np.zeros((2, 3)), u, targetidxs=tidx, interpolation="multilinear")
    np.testing.assert_allclose(got, G[:, :, 0, 0, :].sum(axis=1), rtol=1e-13)


# ------------------------------------------------------------------------------------------ misfit
def test_mvn_chol_golden(golden):
    """multivariate_normal_chol on the GPU vs the reference's own output (diag, banded and dense weights mixed)."""
    import types
    from beat_b200.ops import multivariate_normal_chol
    C, U, lp, res = golden["mvn_C"], golden["mvn_U"], golden["mvn_logpdet"], golden["mvn_res"]
    n_t, ns = res.shape
    datasets = [types.SimpleNamespace(samples=ns, typ="any_P_T", covariance=types.SimpleNamespace(slog_pdet=lp[i], log_pdet=lp[i]))
                for i in range(n_t)]
    got = multivariate_normal_chol(datasets, list(U), {"h_any_P_T": float(golden["mvn_h_scalar"])}, res)
    np.testing.assert_allclose(got, golden["mvn_logpts_scalar"], rtol=1e-12)
    got = multivariate_normal_chol(datasets, list(U), {"h_any_P_T": golden["mvn_h_vec"]}, res, hp_specific=True)
    np.testing.assert_allclose(got, golden["mvn_logpts_vec"], rtol=1e-12)
    # each structure on its own (diag-only -> DIAG path, exponential-only -> BAND path)
    for sel in ([0, 1], [2, 3], [4, 5]):
        ds = [datasets[i] for i in sel]
        got = multivariate_normal_chol(ds, [U[i] for i in sel], {"h_any_P_T": float(golden["mvn_h_scalar"])}, res[sel])
        np.testing.assert_allclose(got, golden["mvn_logpts_scalar"][sel], rtol=1e-12)


def test_mvn_chol_vs_scipy_batched():
    """Property the reference tests (test/test_models.py:149-222), batched over chains."""
    import scipy.stats
    import types
    from beat_b200.covariance import Covariance, exponential_data_covariance
    from beat_b200.ops import multivariate_normal_chol
    rng = np.random.default_rng(2)
    n_t, ns, B = 4, 50, 7
    Cs = [exponential_data_covariance(ns, 0.5, 2.0 + i) * 0.2 ** 2 for i in range(n_t)]
    covs = [Covariance(data=c) for c in Cs]
    datasets = [types.SimpleNamespace(samples=ns, typ="any_P_Z", covariance=c) for c in covs]
    res = rng.standard_normal((B, n_t, ns))
    h = rng.uniform(0, 2, B)
    got = multivariate_normal_chol(datasets, [c.chol_inverse for c in covs], {"h_any_P_Z": h}, res)
    for b in range(B):
        for i in range(n_t):
            ref = scipy.stats.multivariate_normal.logpdf(res[b, i], mean=np.zeros(ns), cov=Cs[i] * np.exp(2 * h[b]))
            assert abs(got[b, i] - ref) <= 1e-9 * abs(ref)


# ------------------------------------------------------------------------------------------ fused path
CASES = {
    "ml_exp": dict(),
    "nn_exp": dict(interpolation="nearest_neighbor"),
    "ml_variance": dict(noise="variance"),
    "ml_dense": dict(noise="dense"),
    "station_corr_hp_specific": dict(station_corrections=True, hp_specific=True),
    "two_subfaults": dict(subfaults=((4, 5, 2.0), (3, 6, 2.5))),
    "three_slipvars": dict(slip_vars=("uparr", "uperp", "utens")),
    "one_slipvar": dict(slip_vars=("uparr",)),
    "two_wavemaps": dict(n_wavemaps=2),
    "long_traces": dict(ns=300, nt=3),                  # > 256 samples: CTA-per-item misfit pass
    "mid_traces": dict(ns=200, nt=3),                   # 129..256 samples: warp-per-item misfit pass, 8 samples per lane
    "mid_traces_variance": dict(ns=161, nt=3, noise="variance"),
    "many_patches": dict(subfaults=((12, 25, 1.0),), nt=2, ns=16),
    "odd_ns": dict(ns=37),
}


@pytest.mark.parametrize("case", sorted(CASES))
@pytest.mark.parametrize("store,rtol", [("float64", 1e-10), ("float32", 1e-5)])
def test_fused_loglike_vs_oracle(case, store, rtol):
    """q -> per-dataset logpts and like, all chains, vs the oracle's restatement of the reference graph."""
    from beat_b200.engine import BatchedFFILogLike
    args = dict(nt=5, subfaults=((4, 6, 2.0),), ns=32, ndur=4, seed=100 + sorted(CASES).index(case))
    args.update(CASES[case])
    prob = synthetic.make_problem(**args)
    B = 24
    Q = synthetic.draw_chains(prob, B, seed=5)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype=store)
    logpts, like = ev(Q)
    ref = _oracle_logpts(prob, Q)
    assert logpts.shape == ref.shape
    np.testing.assert_allclose(logpts, ref, rtol=rtol)
    np.testing.assert_allclose(like, ref.sum(axis=1), rtol=rtol)
    # rupture start times of the fused path are bit-identical to the sequential C restatement
    st = ev.starttimes(B)
    for b in range(0, B, 7):
        _, _, t0 = O.ffi_seismic_eval(prob, synthetic.split_point(prob, Q[b]), impl="port", return_synth=True)
        assert np.array_equal(st[b], t0)
    # single-chain call, reference return convention
    out = ev.logp_forw_func(Q[3])
    np.testing.assert_allclose(out[0], ref[3], rtol=rtol)
    assert ev.ctx.launch_count() > 0
    ev.close()


@pytest.mark.parametrize("geo_mode", ["mma", "simple"])
@pytest.mark.parametrize("nobs,B", [([60, 45], 16), ([150, 70, 5], 70), ([64], 129), ([260, 131], 150)])
def test_fused_joint_geodetic_laplacian(geo_mode, nobs, B, monkeypatch):
    """Config-4 shape: seismic + geodetic static (dense non-Toeplitz C) + laplacian prior, all in one call.
    geo_mode "mma" = FP64 tensor-core GEMM tiles over all chains, "simple" = one CTA per (chain, dataset)."""
    from beat_b200.engine import BatchedFFILogLike
    monkeypatch.setenv("BEATGPU_GEO_MODE", geo_mode)
    prob = synthetic.make_problem(nt=4, subfaults=((5, 7, 2.0),), ns=24, ndur=4, geodetic=dict(nobs=nobs), laplacian=True, seed=3)
    Q = synthetic.draw_chains(prob, B, seed=6)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    logpts, like = ev(Q)
    assert logpts.shape == (B, 4 + len(nobs) + 1)
    for b in range(B):
        pt = synthetic.split_point(prob, Q[b])
        ref = np.concatenate([O.ffi_seismic_eval(prob, pt, impl="port"), O.ffi_geodetic_eval(prob["geodetic"], pt),
                              [O.ffi_laplacian_eval(prob["laplacian"], pt, prob["slip_vars"])]])
        np.testing.assert_allclose(logpts[b], ref, rtol=1e-10)
        np.testing.assert_allclose(like[b], ref.sum(), rtol=1e-10, atol=1e-10 * np.abs(ref).sum())
    ev.close()


def test_fused_fixed_variables():
    """Variables absent from q (the reference's fixed_rvs) come from the `fixed` vector."""
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=3, subfaults=((4, 5, 2.0),), ns=20, ndur=4, seed=9)
    B = 6
    Q = synthetic.draw_chains(prob, B, seed=1)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    full, _ = ev(Q)
    # fix durations and time to chain 0's values; drop them from q
    npatch = prob["npatches"]
    keep = [n for n, _ in prob["var_order"] if n not in ("durations", "time")]
    sizes = dict(prob["var_order"])
    prob2 = dict(prob)
    offs, o = {}, 0
    for n in keep:
        offs[n] = o
        o += sizes[n]
    prob2["offsets"], prob2["n_params"] = offs, o
    nsl, nsf = len(prob["slip_vars"]), len(prob["subfaults"])
    fixed = np.zeros(nsl * npatch + 2 * npatch + 3 * nsf + prob["n_hypers"] + prob["n_time_shifts"])
    fixed[nsl * npatch: (nsl + 1) * npatch] = Q[0, prob["offsets"]["durations"]: prob["offsets"]["durations"] + npatch]
    fixed[(nsl + 2) * npatch + 2 * nsf: (nsl + 2) * npatch + 3 * nsf] = Q[0, prob["offsets"]["time"]: prob["offsets"]["time"] + nsf]
    prob2["fixed"] = fixed
    Q2 = np.concatenate([Q[:, prob["offsets"][n]: prob["offsets"][n] + sizes[n]] for n in keep], axis=1)
    ev2 = BatchedFFILogLike.from_problem(prob2, store_dtype="float64")
    got, _ = ev2(Q2)
    Qref = Q.copy()
    Qref[:, prob["offsets"]["durations"]: prob["offsets"]["durations"] + npatch] = Q[0, prob["offsets"]["durations"]: prob["offsets"]["durations"] + npatch]
    Qref[:, prob["offsets"]["time"]: prob["offsets"]["time"] + nsf] = Q[0, prob["offsets"]["time"]: prob["offsets"]["time"] + nsf]
    ref, _ = ev(Qref)
    assert np.array_equal(got, ref)
    assert np.array_equal(got[0], full[0])
    ev.close()
    ev2.close()


def test_fused_violation_and_update_weights():
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=3, subfaults=((4, 5, 2.0),), ns=20, ndur=4, seed=21)
    Q = synthetic.draw_chains(prob, 8, seed=2)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    base, _ = ev(Q)
    # weights update between SMC stages (seismic.py:1527-1534): scaling C by 4 -> U/2, log_pdet + ns*log(4)
    wm = prob["wavemaps"][0]
    ev.update_weights(0, wm["U"] / 2.0, wm["slog_pdet"] + wm["ns"] * np.log(4.0))
    prob2 = dict(prob)
    wm2 = dict(wm, U=wm["U"] / 2.0, slog_pdet=wm["slog_pdet"] + wm["ns"] * np.log(4.0))
    prob2["wavemaps"] = [wm2]
    got, _ = ev(Q)
    np.testing.assert_allclose(got, _oracle_logpts(prob2, Q), rtol=1e-10)
    assert not np.allclose(got, base)
    # a chain whose origin time pushes start times off the library axis -> IndexError, like the reference
    Qbad = Q.copy()
    Qbad[2, prob["offsets"]["time"]] = 1e4
    with pytest.raises(IndexError):
        ev(Qbad)
    # nucleation point outside the fault -> IndexError as well
    Qbad = Q.copy()
    Qbad[1, prob["offsets"]["nucleation_dip"]] = 1e3
    with pytest.raises(IndexError):
        ev(Qbad)
    # the context stays usable
    got2, _ = ev(Q)
    assert np.array_equal(got, got2)
    ev.close()


def test_device_resident_entry_matches_host_entry():
    import torch
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=4, subfaults=((4, 6, 2.0),), ns=32, ndur=4, seed=33)
    Q = synthetic.draw_chains(prob, 32, seed=3)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float32")
    host_lp, host_like = ev(Q)
    q_dev = torch.from_numpy(Q).cuda()
    lp, like = ev.eval_device(q_dev)
    torch.cuda.synchronize()
    assert np.array_equal(lp.cpu().numpy(), host_lp)
    assert np.array_equal(like.cpu().numpy(), host_like)
    assert ev.ctx.last_stack_ms() > 0
    ev.close()


def test_near_map_accuracy_f32_storage():
    """The hard regime for f32 libraries: chains close to the data-generating model, where the residual is a small
    difference of large synthetics.  logpts must still match the f64 oracle to the north-star rtol 1e-5."""
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=6, subfaults=((6, 10, 2.0),), ns=64, ndur=6, seed=77)
    q_true = synthetic.draw_chains(prob, 1, seed=8)[0]
    _, synths, _ = O.ffi_seismic_eval(prob, synthetic.split_point(prob, q_true), impl="port", return_synth=True)
    rng = np.random.default_rng(4)
    wm = prob["wavemaps"][0]
    sigma = 0.05 * np.abs(synths[0]).max(axis=1)
    from beat_b200.covariance import Covariance, exponential_data_covariance
    for t in range(wm["nt"]):
        C = exponential_data_covariance(wm["ns"], 0.5, 2.0) * sigma[t] ** 2
        cov = Covariance(data=C)
        wm["U"][t], wm["slog_pdet"][t] = cov.chol_inverse, cov.log_pdet
        wm["data"][t] = synths[0][t] + np.linalg.cholesky(C).dot(rng.standard_normal(wm["ns"]))
    B = 64
    Q = np.tile(q_true, (B, 1))
    for name in prob["slip_vars"]:
        o = prob["offsets"][name]
        Q[1:, o:o + prob["npatches"]] += rng.normal(0, 0.01, (B - 1, prob["npatches"]))   # small slip perturbations
    ref = _oracle_logpts(prob, Q)
    for store, rtol in (("float64", 1e-10), ("float32", 1e-5)):
        ev = BatchedFFILogLike.from_problem(prob, store_dtype=store)
        logpts, like = ev(Q)
        np.testing.assert_allclose(logpts, ref, rtol=rtol)
        np.testing.assert_allclose(like, ref.sum(axis=1), rtol=rtol)
        ev.close()


@pytest.mark.parametrize("env", [
    {"BEATGPU_STACK_MODE": "fused", "BEATGPU_PERSISTENT": "0"},
    {"BEATGPU_STACK_MODE": "fused", "BEATGPU_PERSISTENT": "1"},
    {"BEATGPU_STACK_MODE": "chunked", "BEATGPU_CHUNK": "7"},
    {"BEATGPU_STACK_MODE": "chunked", "BEATGPU_CHUNK": "32"},
    {"BEATGPU_STACK_MODE": "chunked", "BEATGPU_CHUNK": "7", "BEATGPU_MISFIT_WARP": "0"},      # CTA-per-item misfit pass for short traces too
])
@pytest.mark.parametrize("case", ["ml_exp", "nn_exp", "station_corr_hp_specific", "two_subfaults", "long_traces", "mid_traces", "odd_ns", "ml_dense"])
def test_stack_execution_modes_vs_oracle(env, case, monkeypatch):
    """Every scheduling variant of the stacking kernel (one CTA per item, persistent CTAs, patch-chunked warps with a
    separate misfit pass) must give the oracle's answer; violations surface as IndexError in all of them."""
    from beat_b200.engine import BatchedFFILogLike
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    args = dict(nt=4, subfaults=((5, 9, 2.0),), ns=32, ndur=4, seed=300 + sorted(CASES).index(case))
    args.update(CASES[case])
    prob = synthetic.make_problem(**args)
    Q = synthetic.draw_chains(prob, 20, seed=15)
    ref = _oracle_logpts(prob, Q)
    for store, rtol in (("float64", 1e-10), ("float32", 1e-5)):
        ev = BatchedFFILogLike.from_problem(prob, store_dtype=store)
        logpts, like = ev(Q)
        np.testing.assert_allclose(logpts, ref, rtol=rtol)
        np.testing.assert_allclose(like, ref.sum(axis=1), rtol=rtol)
        Qbad = Q.copy()
        Qbad[5, prob["offsets"]["time"]] = 1e4
        with pytest.raises(IndexError):
            ev(Qbad)
        got, _ = ev(Q)
        assert np.array_equal(got, logpts)
        ev.close()


def test_get_synthetics_and_stage_weight_update():
    """Stage boundary (seismic.py:1509-1534): synthetics at the MAP point -> residuals -> non-Toeplitz covariance ->
    weights on the GPU -> uploaded -> next stage's llk equals the oracle evaluated with the host-computed weights."""
    import torch
    from beat_b200 import covariance as cv
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=5, subfaults=((4, 6, 2.0),), ns=60, ndur=4, seed=41, station_corrections=True)
    Q = synthetic.draw_chains(prob, 12, seed=3)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    syn = ev.get_synthetics(Q)
    for b in (0, 5, 11):
        _, ref, _ = O.ffi_seismic_eval(prob, synthetic.split_point(prob, Q[b]), impl="port", return_synth=True)
        np.testing.assert_allclose(syn[b], ref[0], rtol=1e-12, atol=1e-12 * np.abs(ref[0]).max())
    np.testing.assert_array_equal(ev.get_synthetics(Q[5]), syn[5])
    logpts, like = ev(Q)
    imap = int(np.argmax(like))
    wm = prob["wavemaps"][0]
    resid = wm["data"] - syn[imap]
    U_dev, lp_dev, _ = cv.weights_from_residuals_device(resid, device=torch.device("cuda", 0))
    U_host, lp_host = cv.weights_from_residuals_host(resid)
    np.testing.assert_allclose(lp_dev.cpu().numpy(), lp_host, rtol=1e-9)
    np.testing.assert_allclose(U_dev.cpu().numpy(), U_host, rtol=1e-6, atol=1e-7 * np.abs(U_host).max())
    ev.update_weights(0, U_dev.cpu().numpy(), lp_dev.cpu().numpy())
    got, _ = ev(Q)
    prob2 = dict(prob, wavemaps=[dict(wm, U=U_host, slog_pdet=lp_host)])
    np.testing.assert_allclose(got, _oracle_logpts(prob2, Q), rtol=1e-7)
    # the same update without leaving the device (structure detection + repack in kernels): identical results, for
    # dense, banded and diagonal weights
    ev.update_weights_device(0, U_dev.contiguous(), lp_dev.contiguous())
    got_dev, _ = ev(Q)
    np.testing.assert_array_equal(got_dev, got)
    for kind in ("band", "diag"):
        Uk = torch.triu(U_dev) if kind == "band" else torch.diag_embed(torch.diagonal(U_dev, dim1=1, dim2=2))
        if kind == "band":
            Uk = Uk - torch.triu(U_dev, diagonal=3)                    # main + two super-diagonals
        ev.update_weights(0, Uk.cpu().numpy(), lp_dev.cpu().numpy())
        a, _ = ev(Q)
        ev.update_weights_device(0, Uk.contiguous(), lp_dev.contiguous())
        b, _ = ev(Q)
        np.testing.assert_array_equal(a, b)
    bad = U_dev.clone()
    bad[1, 2, 3] = float("nan")
    with pytest.raises(ValueError, match="NaN"):
        ev.update_weights_device(0, bad, lp_dev.contiguous())
    ev.close()


def test_fault_geometry_helpers(golden):
    """FaultGeometry index helpers (fault.py:610-632,722-752,866-894) + library index helpers vs golden / oracle."""
    from beat_b200.fault import FaultGeometry, FaultOrdering, positions2idxs
    from beat_b200.ops import SeismicGFLibrary
    for cs in (1.0, 2.0, 2.5):
        got = positions2idxs(golden["pos_in"], cs)
        assert got.dtype == np.int16 and np.array_equal(got, golden[f"pos_idx_{cs}"])
    fault = FaultGeometry(FaultOrdering(npls=[20, 15], npws=[10, 10], patch_sizes_strike=[2.0, 2.5], patch_sizes_dip=[2.0, 2.5]))
    assert fault.npatches == 350 and fault.ordering.get_subfault_discretization(1) == (10, 15)
    rng = np.random.default_rng(0)
    B = 9
    point = dict(velocities=rng.uniform(2.2, 4.5, (B, 350)), nucleation_dip=rng.uniform(0, 19.9, (B, 2)),
                 nucleation_strike=np.stack([rng.uniform(0, 39.9, B), rng.uniform(0, 37.4, B)], axis=1), time=rng.uniform(-5, 5, (B, 2)))
    for index, (nd, ns_, h) in enumerate(((10, 20, 2.0), (10, 15, 2.5))):
        st = fault.point2starttimes(point, index=index)
        assert st.shape == (B, nd, ns_)
        for b in range(B):
            di, si = O.fault_locations2idxs(point["nucleation_dip"][b, index], point["nucleation_strike"][b, index], h, h)
            vel = point["velocities"][b, fault.cum_subfault_npatches[index]: fault.cum_subfault_npatches[index + 1]]
            ref = O.fast_sweep(1.0 / vel, h, di, si, nd, ns_, impl="port").reshape(nd, ns_) + point["time"][b, index]
            assert np.array_equal(st[b], ref)
        one = fault.point2starttimes({k: v[3] for k, v in point.items()}, index=index)
        assert np.array_equal(one, st[3])
    st_min, st_step, dur_min, dur_step = golden["stack_axes"]
    gfs = SeismicGFLibrary(golden["stack_rand_G"], dur_min, dur_step, st_min, st_step)
    for tag, interp in (("nn", "nearest_neighbor"), ("ml", "multilinear")):
        si, sf = gfs.starttimes2idxs(golden["stack_rand_starttimes"], interpolation=interp)
        di, df = gfs.durations2idxs(golden["stack_rand_durations"], interpolation=interp)
        assert np.array_equal(si, golden[f"stack_rand_{tag}_si"]) and np.array_equal(di, golden[f"stack_rand_{tag}_di"])
        if sf is not None:
            assert np.array_equal(sf, golden[f"stack_rand_{tag}_sf"]) and np.array_equal(df, golden[f"stack_rand_{tag}_df"])


def test_eval_device_is_stream_ordered_with_torch():
    """eval_device must run on torch's current stream (the default stream has handle 0): inputs produced by queued
    torch work are seen, and torch work queued afterwards sees the outputs -- without any host synchronisation."""
    import torch
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=4, subfaults=((4, 6, 2.0),), ns=32, ndur=4, seed=55)
    Q = synthetic.draw_chains(prob, 64, seed=9)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    ref_lp, ref_like = ev(Q)
    dev = torch.device("cuda", 0)
    q_host = torch.from_numpy(Q).pin_memory()
    for use_side_stream in (False, True):
        stream = torch.cuda.Stream(dev) if use_side_stream else torch.cuda.current_stream(dev)
        with torch.cuda.stream(stream):
            q_dev = torch.zeros(Q.shape, dtype=torch.float64, device=dev)
            a = torch.randn(4096, 4096, device=dev)
            for _ in range(30):                       # keep the stream busy so the copy below is still pending
                a = (a @ a).clamp_(-1, 1)
            q_dev.copy_(q_host, non_blocking=True)
            lp, like = ev.eval_device(q_dev)
            total = like.sum()                        # torch work queued behind our kernels
            got_like, got_total = like.cpu().numpy(), float(total.cpu())
            got_lp = lp.cpu().numpy()
        assert np.array_equal(got_lp, ref_lp), use_side_stream
        assert np.array_equal(got_like, ref_like)
        np.testing.assert_allclose(got_total, ref_like.sum(), rtol=1e-12)
    ev.close()


@pytest.mark.parametrize("mode", ["mma", "simple"])
@pytest.mark.parametrize("ns,B", [(130, 70), (40, 33), (64, 128), (125, 70), (301, 96)])     # even / odd trace lengths: 16- and 8-byte staging copies
def test_dense_covariance_misfit_gemm_path(mode, ns, B, monkeypatch):
    """Full non-Toeplitz covariance (dense upper-triangular U): in the chunked path |U r|^2 of all chains is a batched
    FP64 tensor-core GEMM per target ("mma") or one matvec per (chain, target) ("simple"); both equal the oracle."""
    from beat_b200.engine import BatchedFFILogLike
    monkeypatch.setenv("BEATGPU_GEO_MODE", mode)
    monkeypatch.setenv("BEATGPU_STACK_MODE", "chunked")
    monkeypatch.setenv("BEATGPU_CHUNK", "8")
    prob = synthetic.make_problem(nt=3, subfaults=((4, 7, 2.0),), ns=ns, ndur=4, noise="dense", hp_specific=True, seed=71)
    Q = synthetic.draw_chains(prob, B, seed=2)
    ref = _oracle_logpts(prob, Q)
    for store, rtol in (("float64", 1e-10), ("float32", 1e-5)):
        ev = BatchedFFILogLike.from_problem(prob, store_dtype=store)
        logpts, like = ev(Q)
        np.testing.assert_allclose(logpts, ref, rtol=rtol)
        Qbad = Q.copy()
        Qbad[4, prob["offsets"]["time"]] = 1e4
        Qbad[9, prob["offsets"]["nucleation_strike"]] = 1e3
        with pytest.raises(IndexError):
            ev(Qbad)
        ev.close()
    # a general (non-triangular) weight matrix, e.g. the QR fallback of chol_inverse, goes through the same path
    rng = np.random.default_rng(0)
    wm = prob["wavemaps"][0]
    Ufull = wm["U"] + 0.01 * rng.standard_normal(wm["U"].shape) * np.abs(wm["U"]).max()
    prob2 = dict(prob, wavemaps=[dict(wm, U=Ufull)])
    ev = BatchedFFILogLike.from_problem(prob2, store_dtype="float64")
    np.testing.assert_allclose(ev(Q)[0], _oracle_logpts(prob2, Q), rtol=1e-10)
    ev.close()


def test_mvn_chol_dense_weights_batched_gemm():
    """multivariate_normal_chol with full (dense) weights and many chains: the tensor-core GEMM route of misfit_batch."""
    import scipy.stats
    import types
    from beat_b200.covariance import Covariance
    from beat_b200.ops import multivariate_normal_chol
    rng = np.random.default_rng(12)
    n_t, ns, B = 3, 150, 80
    Cs = []
    for i in range(n_t):
        a = rng.random((ns, ns))
        Cs.append((a.T.dot(a) + np.eye(ns) * 0.3) * 0.01)                   # recipe of test/test_covariance.py:72-74
    covs = [Covariance(data=c) for c in Cs]
    datasets = [types.SimpleNamespace(samples=ns, typ="any_P_Z", covariance=c) for c in covs]
    res = rng.standard_normal((B, n_t, ns)) * 0.2
    h = rng.uniform(0, 1, B)
    got = multivariate_normal_chol(datasets, [c.chol_inverse for c in covs], {"h_any_P_Z": h}, res)
    for b in range(0, B, 9):
        for i in range(n_t):
            ref = scipy.stats.multivariate_normal.logpdf(res[b, i], mean=np.zeros(ns), cov=Cs[i] * np.exp(2 * h[b]))
            assert abs(got[b, i] - ref) <= 1e-8 * abs(ref)
    ref_all = np.array([O.mvn_chol_logpts(res[b], [c.chol_inverse for c in covs], [c.log_pdet for c in covs], [ns] * n_t, h[b]) for b in range(B)])
    np.testing.assert_allclose(got, ref_all, rtol=1e-11)


def test_misfit_batch_dev_long_traces():
    """Config-2 trace length (2048 samples, Toeplitz covariance -> banded weights): device-pointer misfit entry equals
    the host entry and the oracle."""
    import torch
    from beat_b200.covariance import Covariance, exponential_data_covariance
    from beat_b200.lib import Context
    rng = np.random.default_rng(3)
    nt, ns, B = 3, 2048, 5
    covs = [Covariance(data=exponential_data_covariance(ns, 0.5, 2.0 + i) * 0.05 ** 2) for i in range(nt)]
    U = np.stack([c.chol_inverse for c in covs])
    lp = np.array([c.log_pdet for c in covs])
    ctx = Context(0)
    wid = ctx.add_wavemap(nt, ns, "nearest_neighbor", None, np.arange(nt, dtype=np.int32), np.full(nt, ns, np.int32))
    ctx.update_weights(wid, U, lp)
    res = rng.standard_normal((B, nt, ns)) * 0.05
    hyp = rng.uniform(0, 1, (B, nt))
    host = ctx.misfit_batch(wid, res, hyp)
    for b in range(B):
        ref = O.mvn_chol_logpts(res[b], U, lp, [ns] * nt, hyp[b])
        np.testing.assert_allclose(host[b], ref, rtol=1e-9)
    dev = torch.device("cuda", 0)
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream, external=True)
    r_d, h_d = torch.from_numpy(res).to(dev), torch.from_numpy(hyp).to(dev)
    out = torch.empty((B, nt), dtype=torch.float64, device=dev)
    ctx.misfit_batch_dev(wid, B, r_d.data_ptr(), h_d.data_ptr(), nt, out.data_ptr())
    assert np.array_equal(out.cpu().numpy(), host)
    ctx.close()


@pytest.mark.parametrize("name", ["one_fault_ml", "one_fault_nn_corr", "two_faults_ml_corr"])
def test_cuda_synthetics_match_reference_composite_golden(name):
    """CUDA forward model (sweep + stacking, f64 library) against synthetics produced by the reference's own
    SeismicDistributerComposite.get_synthetics / FaultGeometry.point2starttimes (committed fixture
    tests/golden/ffi_composite_golden.npz, generated by tests/golden/make_ffi_composite_golden.py); no oracle in the loop."""
    from test_oracle_golden import FFI_COMPOSITE_CASES, load_ffi_composite_golden
    from beat_b200.engine import BatchedFFILogLike
    g = load_ffi_composite_golden()
    prob = synthetic.make_problem(**FFI_COMPOSITE_CASES[name])
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    got = ev.get_synthetics(g[name + "_Q"])
    ev.close()
    ref = g[name + "_synths"]
    np.testing.assert_allclose(got, ref, rtol=1e-8, atol=1e-8 * np.abs(ref).max())


@pytest.mark.parametrize("name,hp_specific", [(n, h) for n in ("one_fault_ml", "one_fault_nn_corr", "two_faults_ml_corr") for h in (False, True)])
def test_cuda_loglike_matches_reference_get_formula_golden(name, hp_specific):
    """The fused CUDA evaluation (f64 library) against per-dataset logpts produced by the reference's own production
    graph SeismicDistributerComposite.get_formula, run eagerly (committed fixture tests/golden/ffi_composite_golden.npz,
    tests/golden/make_ffi_composite_golden.py); no oracle in the loop."""
    from test_oracle_golden import FFI_COMPOSITE_CASES, load_ffi_composite_golden
    from beat_b200.engine import BatchedFFILogLike
    g = load_ffi_composite_golden()
    tag = "formula_" + name + ("_hps" if hp_specific else "")
    prob = synthetic.make_problem(hp_specific=hp_specific, **FFI_COMPOSITE_CASES[name])
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    logpts, like = ev(g[tag + "_Q"])
    ev.close()
    np.testing.assert_allclose(logpts, g[tag + "_logpts"], rtol=1e-8)
    np.testing.assert_allclose(like, g[tag + "_logpts"].sum(axis=1), rtol=1e-8)


@pytest.mark.parametrize("geo_mode", ["mma", "simple"])
@pytest.mark.parametrize("name", ["one_dataset", "three_datasets", "two_datasets_one_slipvar"])
def test_cuda_geodetic_loglike_matches_reference_get_formula_golden(name, geo_mode, monkeypatch):
    """The geodetic columns of the fused CUDA evaluation (DMMA GEMM path and the matvec path) against per-dataset logpts
    produced by the reference's own GeodeticDistributerComposite.get_formula, run eagerly (committed fixture
    tests/golden/geodetic_composite_golden.npz, tests/golden/make_geodetic_composite_golden.py); no oracle in the loop.
    The six chains of the fixture are tiled to 36 so that the batched GEMM path (B >= 32) is the one that runs."""
    from test_oracle_golden import GEODETIC_COMPOSITE_CASES, load_geodetic_composite_golden
    from beat_b200.engine import BatchedFFILogLike
    monkeypatch.setenv("BEATGPU_GEO_MODE", geo_mode)
    g = load_geodetic_composite_golden()
    prob = synthetic.make_problem(**GEODETIC_COMPOSITE_CASES[name])
    Q, ref = np.tile(g[name + "_Q"], (6, 1)), np.tile(g[name + "_logpts"], (6, 1))
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    logpts, like = ev(Q)
    ev.close()
    nt, nd = prob["wavemaps"][0]["nt"], ref.shape[1]
    assert logpts.shape == (36, nt + nd)
    np.testing.assert_allclose(logpts[:, nt:nt + nd], ref, rtol=1e-10)
    np.testing.assert_allclose(like, logpts.sum(axis=1), rtol=1e-10, atol=1e-10 * np.abs(logpts).sum(axis=1).max())
