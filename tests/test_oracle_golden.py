"""CPU tests: the oracle (oracle/) against the golden vectors produced by the reference's own code."""
import numpy as np
import pytest

from oracle import ffi_oracle as O


def _fs_cases(golden):
    for i in range(int(golden["fs_ncases"])):
        nd, ns, nuc_dip, nuc_strike = (int(v) for v in golden[f"fs{i}_meta"])
        yield i, nd, ns, nuc_dip, nuc_strike, float(golden[f"fs{i}_h"]), golden[f"fs{i}_slow"]


def test_fast_sweep_port_vs_reference_c(golden):
    """Plain-C restatement vs the reference's compiled C: <= 2 ulp (glibc pow(x,.5) vs sqrt), identical indices."""
    for i, nd, ns, nuc_dip, nuc_strike, h, slow in _fs_cases(golden):
        t = O.fast_sweep(slow, h, nuc_dip, nuc_strike, nd, ns, impl="port")
        ref = golden[f"fs{i}_t_c"]
        assert np.all(np.abs(t - ref) <= 4 * np.spacing(np.abs(ref))), i
        # the reference's own cross-implementation gate (test/test_fastsweep.py:131-133)
        np.testing.assert_allclose(t, golden[f"fs{i}_t_numpy"], rtol=0, atol=1e-6)
        for interp in ("nearest_neighbor", "multilinear"):
            a, _ = O.times2idxs(t, -5.0, 0.5, interp)
            b, _ = O.times2idxs(ref, -5.0, 0.5, interp)
            assert np.array_equal(a, b)


def test_fast_sweep_kat_reference_test_case(golden):
    """Known-answer: reference test inputs (test/test_fastsweep.py:21-31); values as in BASELINE.md."""
    t = O.fast_sweep(golden["fs0_slow"], 10.0, 3, 2, 6, 4, impl="port").reshape(6, 4)
    np.testing.assert_allclose(t[3], [20.0, 10.0, 0.0, 2.8571428571428568], rtol=0, atol=1e-13)
    np.testing.assert_allclose(t[0], [27.757645028504108, 18.1593235621289, 8.57142857142857, 9.834944019440146],
                               rtol=0, atol=1e-13)


def test_fast_sweep_ref_binary_if_present(golden):
    if O.load_reference_ext() is None:
        pytest.skip("oracle/_ref not built (no /root/reference on this box)")
    for i, nd, ns, nuc_dip, nuc_strike, h, slow in _fs_cases(golden):
        t = O.fast_sweep(slow, h, nuc_dip, nuc_strike, nd, ns, impl="ref")
        assert np.array_equal(t, golden[f"fs{i}_t_c"])


def test_port_vs_ref_random_grids():
    """Wider sweep of random grids, only where the reference binary is available."""
    if O.load_reference_ext() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    nbit = 0
    for _ in range(400):
        nd, ns = int(rng.integers(1, 24)), int(rng.integers(1, 40))
        slow = 1.0 / rng.uniform(2.2, 4.5, nd * ns)
        hr, hc = int(rng.integers(0, nd)), int(rng.integers(0, ns))
        h = float(rng.choice([1.0, 2.0, 2.5, 5.0]))
        a = O.fast_sweep(slow, h, hr, hc, nd, ns, impl="ref")
        b = O.fast_sweep(slow, h, hr, hc, nd, ns, impl="port")
        assert np.all(np.abs(a - b) <= 4 * np.spacing(np.abs(a)))
        nbit += np.array_equal(a, b)
        ia, _ = O.times2idxs(a, -5.0, 0.5, "nearest_neighbor")
        ib, _ = O.times2idxs(b, -5.0, 0.5, "nearest_neighbor")
        assert np.array_equal(ia, ib)
    assert nbit >= 380


def test_positions2idxs(golden):
    for cs in (1.0, 2.0, 2.5):
        got = O.positions2idxs(golden["pos_in"], cs)
        assert got.dtype == np.int16
        assert np.array_equal(got, golden[f"pos_idx_{cs}"])


@pytest.mark.parametrize("name", ["rand", "recipe"])
@pytest.mark.parametrize("tag,interp", [("nn", "nearest_neighbor"), ("ml", "multilinear")])
def test_stack_all(golden, name, tag, interp):
    st_min, st_step, dur_min, dur_step = golden["stack_axes"]
    G = golden[f"stack_{name}_G"]
    d, s, u = golden[f"stack_{name}_durations"], golden[f"stack_{name}_starttimes"], golden[f"stack_{name}_slips"]
    got = O.stack_all(G, d, s, u, dur_min, dur_step, st_min, st_step, interp)
    ref = golden[f"stack_{name}_{tag}"]
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=1e-13 * np.abs(ref).max())
    loops = O.stack_all_loops(G, d, s, u, dur_min, dur_step, st_min, st_step, interp)
    np.testing.assert_allclose(loops, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    si, sf = O.times2idxs(s, st_min, st_step, interp)
    di, df = O.times2idxs(d, dur_min, dur_step, interp)
    assert np.array_equal(si, golden[f"stack_{name}_{tag}_si"]) and si.dtype == np.int16
    assert np.array_equal(di, golden[f"stack_{name}_{tag}_di"])
    if interp == "multilinear":
        assert np.array_equal(sf, golden[f"stack_{name}_{tag}_sf"])
        assert np.array_equal(df, golden[f"stack_{name}_{tag}_df"])


def test_geodetic_stack(golden):
    np.testing.assert_allclose(O.geodetic_stack_all(golden["geo_G"], golden["geo_slips"]), golden["geo_mu"],
                               rtol=1e-14, atol=1e-14)


def test_covariance_and_mvn(golden):
    C, U, lp = golden["mvn_C"], golden["mvn_U"], golden["mvn_logpdet"]
    n_t, ns = C.shape[0], C.shape[1]
    for i in range(n_t):
        np.testing.assert_allclose(O.chol_inverse(C[i]), U[i], rtol=1e-10, atol=1e-10 * np.abs(U[i]).max())
        np.testing.assert_allclose(O.log_pdet(C[i]), lp[i], rtol=1e-13)
    res = golden["mvn_res"]
    got = O.mvn_chol_logpts(res, U, lp, [ns] * n_t, float(golden["mvn_h_scalar"]))
    np.testing.assert_allclose(got, golden["mvn_logpts_scalar"], rtol=1e-13)
    got = O.mvn_chol_logpts(res, U, lp, [ns] * n_t, golden["mvn_h_vec"])
    np.testing.assert_allclose(got, golden["mvn_logpts_vec"], rtol=1e-13)
    np.testing.assert_array_equal(O.exponential_data_covariance(16, 0.5, 2.0), golden["cov_exp_n16"])


def test_mvn_vs_scipy():
    """Property the reference tests (test/test_models.py:149-222): llk == scipy logpdf of N(0, C e^{2h})."""
    import scipy.stats
    rng = np.random.default_rng(1)
    ns = 40
    C = O.exponential_data_covariance(ns, 0.5, 2.0) * 0.3 ** 2
    U, lp = O.chol_inverse(C), O.log_pdet(C)
    r = rng.standard_normal(ns)
    for h in (0.0, 0.7, -0.4):
        got = O.mvn_chol_logpts([r], [U], [lp], [ns], h)[0]
        ref = scipy.stats.multivariate_normal.logpdf(r, mean=np.zeros(ns), cov=C * np.exp(2 * h))
        assert abs(got - ref) < 1e-9 * abs(ref)


def test_exponential_U_is_bidiagonal():
    """chol(inv(C_exponential)).T is upper-bidiagonal to ~1e-14 -- basis of the banded misfit path."""
    for n, dt, t0 in ((120, 0.5, 2.0), (128, 0.5, 2.0), (64, 1.0, 5.0)):
        U = O.chol_inverse(O.exponential_data_covariance(n, dt, t0) * 0.01)
        off = U - np.diag(np.diag(U)) - np.diag(np.diag(U, 1), 1)
        assert np.abs(off).max() < 1e-12 * np.abs(U).max()


def test_laplacian(golden):
    nstr, ndip, hs, hd = golden["lap_dims"]
    L = O.smoothing_operator_nearest_neighbor(int(nstr), int(ndip), hs, hd)
    np.testing.assert_array_equal(L, golden["lap_L"])
    got = O.laplacian_logpt(L, float(golden["lap_sdet"]), golden["lap_u"], float(golden["lap_h"]))
    np.testing.assert_allclose(got, float(golden["lap_logpt"]), rtol=1e-14)


def test_noise_estimators_host_mirrors(golden):
    """Host mirrors of the per-stage covariance update (product side, set-up time) vs the reference's own output."""
    from beat_b200 import covariance as cv
    resid, ws = golden["nt_resid"], int(golden["nt_window"])
    np.testing.assert_allclose(cv.running_window_rms(resid, ws, mode="same"), golden["nt_rms_same"], rtol=1e-14)
    np.testing.assert_allclose(cv.autocovariance(resid / golden["nt_rms_same"]), golden["nt_autocov"], rtol=1e-11, atol=1e-14)
    toe, stds = cv.toeplitz_covariance(resid, ws)
    np.testing.assert_allclose(toe, golden["nt_toeplitz"], rtol=1e-11, atol=1e-14)
    np.testing.assert_allclose(cv.non_toeplitz_covariance(resid, ws), golden["nt_cov"], rtol=1e-11, atol=1e-14)
    U, lp = cv.weights_from_residuals_host(resid[None, :])
    np.testing.assert_allclose(U[0], golden["nt_U"], rtol=1e-8, atol=1e-9 * np.abs(golden["nt_U"]).max())
    np.testing.assert_allclose(lp[0], float(golden["nt_logpdet"]), rtol=1e-11)
    # Covariance mirror: chol_inverse / log_pdet on the mvn golden matrices
    for C, Ug, lg in zip(golden["mvn_C"], golden["mvn_U"], golden["mvn_logpdet"]):
        c = cv.Covariance(data=C)
        np.testing.assert_allclose(c.chol_inverse, Ug, rtol=1e-10, atol=1e-10 * np.abs(Ug).max())
        np.testing.assert_allclose(c.log_pdet, lg, rtol=1e-13)
    np.testing.assert_array_equal(cv.smoothing_operator_nearest_neighbor(5, 4, 2.0, 2.0), golden["lap_L"])
    np.testing.assert_allclose(cv.log_determinant(golden["lap_L"].T * golden["lap_L"]), float(golden["lap_sdet"]), rtol=1e-13)


def test_weights_from_residuals_torch_cpu(golden):
    """The batched torch implementation (run here on CPU tensors) equals the host mirror."""
    import torch
    from beat_b200 import covariance as cv
    rng = np.random.default_rng(3)
    t = np.arange(60)
    res = np.stack([golden["nt_resid"], np.cos(t / 3.0) * (1 + t / 40.0) + 0.2 * rng.standard_normal(60),
                    rng.standard_normal(60) * np.linspace(0.5, 2.0, 60)])
    U, lp, C = cv.weights_from_residuals_device(res, device=torch.device("cpu"))
    Uh, lph = cv.weights_from_residuals_host(res)
    np.testing.assert_allclose(C[0].numpy(), golden["nt_cov"], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(lp.numpy(), lph, rtol=1e-9)
    for i in range(3):
        # U is unique up to nothing (Cholesky with positive diagonal): compare directly and through U^T U = C^-1
        np.testing.assert_allclose(U[i].numpy(), Uh[i], rtol=1e-6, atol=1e-7 * np.abs(Uh[i]).max())


FFI_COMPOSITE_CASES = {
    "one_fault_ml": dict(nt=4, subfaults=((4, 6, 2.0),), ns=24, ndur=4, seed=301, interpolation="multilinear"),
    "one_fault_nn_corr": dict(nt=5, subfaults=((3, 5, 2.5),), ns=20, ndur=3, seed=302, interpolation="nearest_neighbor",
                              station_corrections=True),
    "two_faults_ml_corr": dict(nt=3, subfaults=((3, 4, 2.0), (2, 5, 2.0)), ns=16, ndur=4, seed=303, interpolation="multilinear",
                               station_corrections=True),
}


def load_ffi_composite_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ffi_composite_golden.npz"))


@pytest.mark.parametrize("name", sorted(FFI_COMPOSITE_CASES))
def test_oracle_matches_reference_composite_synthetics(name):
    """tests/golden/ffi_composite_golden.npz: synthetics from the reference's OWN
    SeismicDistributerComposite.get_synthetics (beat/models/seismic.py:1351-1507) with FaultGeometry.point2starttimes
    (numpy fast sweep), station corrections and SeismicGFLibrary.stack_all per slip component
    (tests/golden/make_ffi_composite_golden.py).  The oracle's composite evaluation must reproduce them."""
    from beat_b200 import synthetic
    g = load_ffi_composite_golden()
    prob = synthetic.make_problem(**FFI_COMPOSITE_CASES[name])
    for q, ref in zip(g[name + "_Q"], g[name + "_synths"]):
        _, mine, _ = O.ffi_seismic_eval(prob, synthetic.split_point(prob, q), impl="auto", return_synth=True)
        np.testing.assert_allclose(mine[0], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())


FORMULA_CASES = [(n, hps) for n in sorted(FFI_COMPOSITE_CASES) for hps in (False, True)]


@pytest.mark.parametrize("name,hp_specific", FORMULA_CASES)
def test_oracle_matches_reference_get_formula(name, hp_specific):
    """Golden per-dataset logpts from the reference's OWN production graph, SeismicDistributerComposite.get_formula
    (beat/models/seismic.py:1210-1349), executed eagerly through the numpy-backed pytensor shim: Sweeper Op -> compiled
    fast_sweep_ext, station corrections, stack_all in pytensor mode (batched_dot), residuals,
    multivariate_normal_chol with scalar and dataset-specific hyperparameters (make_ffi_composite_golden.py)."""
    from beat_b200 import synthetic
    g = load_ffi_composite_golden()
    tag = "formula_" + name + ("_hps" if hp_specific else "")
    prob = synthetic.make_problem(hp_specific=hp_specific, **FFI_COMPOSITE_CASES[name])
    for q, ref in zip(g[tag + "_Q"], g[tag + "_logpts"]):
        mine = O.ffi_seismic_eval(prob, synthetic.split_point(prob, q), impl="auto")
        np.testing.assert_allclose(mine, ref, rtol=1e-10)


GEODETIC_COMPOSITE_CASES = {
    "one_dataset": dict(nt=2, subfaults=((4, 6, 2.0),), ns=16, ndur=3, seed=401, geodetic=dict(nobs=[40])),
    "three_datasets": dict(nt=2, subfaults=((3, 5, 2.5),), ns=16, ndur=3, seed=402, geodetic=dict(nobs=[30, 17, 5])),
    "two_datasets_one_slipvar": dict(nt=2, subfaults=((3, 4, 2.0),), ns=16, ndur=3, seed=403, geodetic=dict(nobs=[12, 21]),
                                     slip_vars=("uparr",)),
}


def load_geodetic_composite_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geodetic_composite_golden.npz"))


@pytest.mark.parametrize("name", sorted(GEODETIC_COMPOSITE_CASES))
def test_oracle_matches_reference_geodetic_get_formula(name):
    """Golden per-dataset logpts ("geo_like") from the reference's OWN GeodeticDistributerComposite.get_formula
    (beat/models/geodetic.py:1030-1084), executed eagerly through the numpy-backed pytensor shim: GeodeticGFLibrary.stack_all
    per slip component, (sdata - mu) * sodws, Bij.srmap, multivariate_normal_chol
    (tests/golden/make_geodetic_composite_golden.py).  Pins rows a6 / a7 / a8 of the oracle at composite level."""
    from beat_b200 import synthetic
    g = load_geodetic_composite_golden()
    prob = synthetic.make_problem(**GEODETIC_COMPOSITE_CASES[name])
    for q, ref in zip(g[name + "_Q"], g[name + "_logpts"]):
        mine = O.ffi_geodetic_eval(prob["geodetic"], synthetic.split_point(prob, q))
        np.testing.assert_allclose(mine, ref, rtol=1e-10)
