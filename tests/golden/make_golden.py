"""
tests/golden/make_golden.py -- regenerates tests/golden/*.npz by running the REFERENCE'S OWN code.

Run in the dev container only (needs /root/reference and oracle/_ref built by ``make -C oracle``):

    python tests/golden/make_golden.py

What runs is the reference's source, imported from /root/reference:
  * fast sweeping: ``beat.pytensorf.Sweeper.perform`` with implementation "c" (-> the reference's compiled
    ``fast_sweep_ext``) and "numpy" (``beat.fast_sweeping.fast_sweep.get_rupture_times_numpy``);
    first case = the reference's own test inputs (test/test_fastsweep.py:21-31);
  * ``beat.utility.positions2idxs``;
  * ``beat.ffi.base.SeismicGFLibrary.stack_all`` (numpy mode AND the pytensor branch through the numpy-backed
    shim, nearest_neighbor + multilinear), ``starttimes2idxs`` / ``durations2idxs``,
    ``GeodeticGFLibrary.stack_all``; library recipe after test/test_ffi.py:22-89 plus a random library;
  * ``beat.models.distributions.multivariate_normal_chol`` (hp_specific False/True) over
    ``beat.heart.Covariance`` objects (``chol_inverse``, ``log_pdet``), covariance structures from
    ``beat.covariance.exponential_data_covariance`` and the random-SPD recipe of
    test/test_covariance.py:71-75; cross-checked here against scipy's logpdf as test/test_models.py does;
  * ``beat.models.laplacian.get_smoothing_operator_nearest_neighbor``, ``beat.heart.log_determinant`` and
    ``LaplacianDistributerComposite._eval_prior``.

pytensor / pyrocko / pymc are not installed in this container; ``_refshim`` supplies inert stand-ins for
them (see its docstring).  The generated vectors are small and committed; the GPU box never runs this.
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _refshim  # noqa: E402
from oracle import ffi_oracle  # noqa: E402

warnings.simplefilter("ignore")
ext = ffi_oracle.load_reference_ext()
assert ext is not None, "build oracle/_ref first: make -C oracle"
_refshim.install(fast_sweep_ext=ext)

from beat import covariance as bcov  # noqa: E402
from beat import heart, pytensorf, utility  # noqa: E402
from beat.config import GeodeticGFLibraryConfig, SeismicGFLibraryConfig  # noqa: E402
from beat.fast_sweeping import fast_sweep as bfs  # noqa: E402
from beat.ffi import base as ffibase  # noqa: E402
from beat.models import distributions, laplacian  # noqa: E402


def gen_fast_sweep(out):
    rng = np.random.default_rng(20240901)
    cases = []
    # reference's own test case: test/test_fastsweep.py:21-31
    v = np.concatenate((np.ones((6, 2)), np.ones((6, 2)) * 3.5), axis=1)
    cases.append(dict(nd=6, ns=4, h=10.0, nuc_dip=3, nuc_strike=2, slow=(1.0 / v).flatten()))
    for nd, ns, h in [(10, 20, 2.0), (10, 15, 2.5), (4, 7, 1.0), (1, 9, 2.0), (12, 1, 3.0), (16, 33, 1.5),
                      (2, 2, 5.0), (24, 40, 1.0), (10, 20, 2.0), (10, 20, 2.0)]:
        cases.append(dict(nd=nd, ns=ns, h=h, nuc_dip=int(rng.integers(0, nd)), nuc_strike=int(rng.integers(0, ns)),
                          slow=1.0 / rng.uniform(2.2, 4.5, nd * ns)))
    # strongly heterogeneous medium (more outer iterations)
    cases.append(dict(nd=10, ns=20, h=2.0, nuc_dip=9, nuc_strike=0, slow=1.0 / rng.uniform(0.3, 6.0, 200)))
    for i, c in enumerate(cases):
        sw_c = pytensorf.Sweeper(c["h"], c["nd"], c["ns"], "c")
        sw_n = pytensorf.Sweeper(c["h"], c["nd"], c["ns"], "numpy")
        oc, on = [[None]], [[None]]
        sw_c.perform(None, [c["slow"], np.int64(c["nuc_dip"]), np.int64(c["nuc_strike"])], oc)
        sw_n.perform(None, [c["slow"], np.int64(c["nuc_dip"]), np.int64(c["nuc_strike"])], on)
        t_c, t_n = oc[0][0], on[0][0]
        np.testing.assert_allclose(t_c, t_n, rtol=0, atol=1e-6)   # the reference's own gate
        out[f"fs{i}_meta"] = np.array([c["nd"], c["ns"], c["nuc_dip"], c["nuc_strike"]], dtype=np.int64)
        out[f"fs{i}_h"] = np.float64(c["h"])
        out[f"fs{i}_slow"] = c["slow"]
        out[f"fs{i}_t_c"] = t_c
        out[f"fs{i}_t_numpy"] = t_n
    out["fs_ncases"] = np.int64(len(cases))


def gen_positions(out):
    pos = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 2.5, 3.0, 3.49999, 3.5, 5.0, 7.0, 9.0, 11.0, 19.99, 39.99, 12.345, 1e-9])
    for cs in (1.0, 2.0, 2.5):
        out[f"pos_idx_{cs}"] = utility.positions2idxs(pos, cs)
    out["pos_in"] = pos


def _mk_seis_lib(G, st_min, st_step, dur_min, dur_step):
    cfg = SeismicGFLibraryConfig(dimensions=G.shape, starttime_min=st_min, starttime_sampling=st_step,
                                 duration_min=dur_min, duration_sampling=dur_step)
    lib = ffibase.SeismicGFLibrary(config=cfg)
    lib._gfmatrix = G
    lib._tmins = np.zeros(G.shape[0])
    lib._stack_switch = {"numpy": G}
    lib.set_stack_mode("numpy")
    return lib


def gen_stack(out):
    rng = np.random.default_rng(77)
    nt, npatch, ndur, nst, ns = 5, 12, 4, 9, 16
    st_min, st_step, dur_min, dur_step = -1.0, 0.5, 0.5, 0.25
    G = rng.standard_normal((nt, npatch, ndur, nst, ns))
    # second library: the reference's test recipe (test/test_ffi.py:80-89): arange(ns) * target index
    G2 = np.tile(np.arange(ns, dtype=float), nt * npatch * ndur * nst).reshape(nt, npatch, ndur, nst, ns)
    G2 = G2 * np.arange(nt)[:, None, None, None, None]
    for name, lib_arr in (("rand", G), ("recipe", G2)):
        lib = _mk_seis_lib(lib_arr, st_min, st_step, dur_min, dur_step)
        # strictly interior so every multilinear floor tap exists
        durations = rng.uniform(dur_min + 1e-3, dur_min + (ndur - 1) * dur_step - 1e-3, npatch)
        starttimes = rng.uniform(st_min + 1e-3, st_min + (nst - 1) * st_step - 1e-3, (nt, npatch))
        # a few values exactly on the grid / on .5 boundaries
        durations[0] = dur_min + dur_step            # integral -> ceil == x, factor 0
        starttimes[:, 1] = st_min + 2 * st_step
        starttimes[0, 2] = st_min + 2.5 * st_step    # rint tie
        slips = rng.uniform(-0.3, 6.0, npatch)
        tidx = np.atleast_2d(np.arange(nt)).T
        out[f"stack_{name}_G"] = lib_arr
        out[f"stack_{name}_durations"] = durations
        out[f"stack_{name}_starttimes"] = starttimes
        out[f"stack_{name}_slips"] = slips
        for interp, tag in (("nearest_neighbor", "nn"), ("multilinear", "ml")):
            lib.set_stack_mode("numpy")
            r_np = lib.stack_all(durations, starttimes, slips, targetidxs=tidx, interpolation=interp)
            # pytensor branch (tt.batched_dot) through the numpy-backed shim
            lib._sgfmatrix = lib_arr.view(_refshim._ND)
            lib.spatchidxs = lib.patchidxs
            lib._stack_switch["pytensor"] = lib._sgfmatrix
            lib.set_stack_mode("pytensor")
            r_pt = lib.stack_all(durations, starttimes, slips, targetidxs=tidx, interpolation=interp)
            np.testing.assert_allclose(r_np, np.asarray(r_pt).reshape(r_np.shape), rtol=0, atol=1e-9)
            lib.set_stack_mode("numpy")
            out[f"stack_{name}_{tag}"] = r_np
            si, sf = lib.starttimes2idxs(starttimes, interpolation=interp)
            di, df = lib.durations2idxs(durations, interpolation=interp)
            out[f"stack_{name}_{tag}_si"] = si
            out[f"stack_{name}_{tag}_di"] = di
            if sf is not None:
                out[f"stack_{name}_{tag}_sf"] = sf
                out[f"stack_{name}_{tag}_df"] = df
    out["stack_axes"] = np.array([st_min, st_step, dur_min, dur_step])

    # geodetic
    Gg = rng.standard_normal((npatch, 37))
    cfg = GeodeticGFLibraryConfig(dimensions=Gg.shape)
    glib = ffibase.GeodeticGFLibrary(config=cfg)
    glib._gfmatrix = Gg
    glib._stack_switch = {"numpy": Gg}
    glib.set_stack_mode("numpy")
    u = rng.uniform(0, 3, npatch)
    out["geo_G"], out["geo_slips"], out["geo_mu"] = Gg, u, glib.stack_all(slips=u)


def gen_mvn(out):
    import scipy.stats
    rng = np.random.default_rng(5)
    n_t, ns = 6, 24
    covs, datasets = [], []
    for i in range(n_t):
        if i < 2:        # "variance"
            C = np.eye(ns) * (0.001 * (i + 1))
        elif i < 4:      # "exponential" toeplitz, beat/covariance.py:24-51 scaled by a variance
            C = bcov.exponential_data_covariance(ns, 0.5, 2.0 + i) * (0.05 * (i + 1)) ** 2
        else:            # random SPD, test/test_covariance.py:72-74
            a = rng.random((ns, ns))
            C = a.T.dot(a) + np.eye(ns) * 0.3
        cov = heart.Covariance(data=C)
        ds = types.SimpleNamespace(samples=ns, typ="any_P_T", covariance=cov)
        covs.append(C)
        datasets.append(ds)
    U = [ds.covariance.chol_inverse for ds in datasets]
    lp = np.array([float(ds.covariance.log_pdet) for ds in datasets])
    for i in range(n_t):
        np.testing.assert_allclose(U[i].T.dot(U[i]), np.linalg.inv(covs[i]), rtol=0, atol=1e-6 * np.abs(np.linalg.inv(covs[i])).max())
    res = rng.standard_normal((n_t, ns)) * 0.1
    h_scalar = 0.37
    h_vec = rng.uniform(-1, 2, n_t)
    lp_scalar = distributions.multivariate_normal_chol(datasets, U, {"h_any_P_T": h_scalar}, res, hp_specific=False)
    lp_vec = distributions.multivariate_normal_chol(datasets, U, {"h_any_P_T": h_vec}, res, hp_specific=True)
    # sanity as in test/test_models.py:149-222: equals scipy logpdf with C scaled by exp(2h)
    for i in range(n_t):
        ref = scipy.stats.multivariate_normal.logpdf(res[i], mean=np.zeros(ns), cov=covs[i] * np.exp(2 * h_scalar))
        assert abs(ref - lp_scalar[i]) < 1e-6 * max(1, abs(ref)), (i, ref, lp_scalar[i])
    out["mvn_C"] = np.array(covs)
    out["mvn_U"] = np.array(U)
    out["mvn_logpdet"] = lp
    out["mvn_res"] = res
    out["mvn_h_scalar"] = np.float64(h_scalar)
    out["mvn_h_vec"] = h_vec
    out["mvn_logpts_scalar"] = np.asarray(lp_scalar)
    out["mvn_logpts_vec"] = np.asarray(lp_vec)
    out["cov_exp_n16"] = bcov.exponential_data_covariance(16, 0.5, 2.0)


def gen_laplacian(out):
    rng = np.random.default_rng(9)
    nstr, ndip, hs, hd = 5, 4, 2.0, 2.0
    L = laplacian.get_smoothing_operator_nearest_neighbor(nstr, ndip, hs, hd)
    sdet = heart.log_determinant(L.T * L, inverse=False)      # as laplacian.py:57-60 (elementwise product)
    u = rng.uniform(0, 4, nstr * ndip)
    fake = types.SimpleNamespace(sdet_shared_smoothing_op=sdet, spatches=nstr * ndip)
    Ls = L.dot(u)
    val = laplacian.LaplacianDistributerComposite._eval_prior(fake, 0.8, Ls.T.dot(Ls))
    out["lap_L"], out["lap_sdet"], out["lap_u"], out["lap_h"], out["lap_logpt"] = L, np.float64(sdet), u, np.float64(0.8), np.float64(val)
    out["lap_dims"] = np.array([nstr, ndip, hs, hd])


def gen_noise_estimators(out):
    """Per-stage covariance update operands (beat/covariance.py:716-771, beat/utility.py:1141-1161)."""
    rng = np.random.default_rng(21)
    n = 60
    t = np.arange(n)
    resid = np.sin(t / 5.0) * (0.5 + t / n) + 0.3 * rng.standard_normal(n)      # non-stationary residual
    ws = n // 5                                                                   # covariance.py:316
    out["nt_resid"] = resid
    out["nt_window"] = np.int64(ws)
    out["nt_rms_same"] = utility.running_window_rms(resid, window_size=ws, mode="same")
    out["nt_autocov"] = bcov.autocovariance(resid / out["nt_rms_same"])
    toe, stds = bcov.toeplitz_covariance(resid, ws)
    out["nt_toeplitz"], out["nt_stds"] = toe, stds
    C = bcov.non_toeplitz_covariance(resid, ws)
    out["nt_cov"] = C
    Cpsd = utility.ensure_cov_psd(C)
    cov = heart.Covariance(data=Cpsd)
    out["nt_cov_psd"] = Cpsd
    out["nt_U"] = cov.chol_inverse
    out["nt_logpdet"] = np.float64(cov.log_pdet)


def main():
    out = {}
    gen_fast_sweep(out)
    gen_positions(out)
    gen_stack(out)
    gen_mvn(out)
    gen_laplacian(out)
    gen_noise_estimators(out)
    path = os.path.join(HERE, "reference_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays,", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
