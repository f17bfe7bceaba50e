"""
tests/golden/make_geodetic_composite_golden.py -- regenerates tests/golden/geodetic_composite_golden.npz.

Runs the REFERENCE'S OWN production graph of the geodetic finite-fault composite, imported from /root/reference in the
dev container and executed eagerly through the numpy-backed pytensor shim (tests/golden/_refshim.py):

    GeodeticDistributerComposite.get_formula                                  beat/models/geodetic.py:1030-1084
      -> GeodeticGFLibrary.stack_all per slip component (pytensor mode)        beat/ffi/base.py:292-305
      -> residuals = Bij.srmap((sdata - mu) * sodws)                           geodetic.py:1072-1074, utility.py:167-350
      -> multivariate_normal_chol (one hyperparameter per dataset type)        beat/models/distributions.py:72-140

for seeded problems of ``beat_b200.synthetic.make_problem(geodetic=...)`` (one and three datasets, dense non-Toeplitz
covariance as test/test_covariance.py:72-74) and stores chain parameters + per-dataset logpts ("geo_like").  Objects the
method reads from ``self`` that need a project directory (config tree, GF files, noise analysis) are attribute bags;
the library class, the list <-> array bijection and the likelihood are the reference's.  These vectors pin
``oracle.ffi_oracle.ffi_geodetic_eval`` (rows a6 / a7 / a8 at composite level); the CUDA path is tested against them.

    python tests/golden/make_geodetic_composite_golden.py
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
from beat_b200 import synthetic as S  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402

warnings.simplefilter("ignore")

CASES = {
    "one_dataset": dict(nt=2, subfaults=((4, 6, 2.0),), ns=16, ndur=3, seed=401, geodetic=dict(nobs=[40])),
    "three_datasets": dict(nt=2, subfaults=((3, 5, 2.5),), ns=16, ndur=3, seed=402, geodetic=dict(nobs=[30, 17, 5])),
    "two_datasets_one_slipvar": dict(nt=2, subfaults=((3, 4, 2.0),), ns=16, ndur=3, seed=403, geodetic=dict(nobs=[12, 21]),
                                     slip_vars=("uparr",)),
}


def main():
    _refshim.install(fast_sweep_ext=O.load_reference_ext())
    from beat import utility
    from beat.config import GeodeticGFLibraryConfig
    from beat.ffi import base as ffibase
    from beat.models import geodetic as rgeodetic

    captured = {}

    def deterministic(name, var):
        captured[name] = np.asarray(var)
        return var
    rgeodetic.Deterministic = deterministic

    out = {}
    for name, kw in CASES.items():
        prob = S.make_problem(**kw)
        geo = prob["geodetic"]
        nobs_list = [hi - lo for lo, hi in geo["slices"]]
        libs = {}
        for v in prob["slip_vars"]:
            G = np.ascontiguousarray(geo["G"][v])
            lib = ffibase.GeodeticGFLibrary(config=GeodeticGFLibraryConfig(dimensions=G.shape))
            lib._gfmatrix = G
            lib._sgfmatrix = G.view(_refshim._ND)                  # what init_optimization would share (ffi/base.py:220-235)
            lib._stack_switch = {"numpy": G, "pytensor": lib._sgfmatrix}
            lib.set_stack_mode("pytensor")
            lib.init_optimization = lambda: None
            libs[v] = lib
        # one hyperparameter per dataset (distinct dataset types), log-determinant carried like heart.Covariance does
        datasets = [types.SimpleNamespace(samples=int(geo["nsamples"][d]), typ="SAR%d" % d,
                                          covariance=types.SimpleNamespace(slog_pdet=np.float64(geo["slog_pdet"][d])))
                    for d in range(len(nobs_list))]
        lists = [np.zeros(n) for n in nobs_list]
        ordering = utility.ListArrayOrdering(lists, intype="numpy")
        Bij = utility.ListToArrayBijection(ordering, lists)
        assert ordering.size == geo["data"].size
        Q = S.draw_chains(prob, 6, seed=700)
        ref_logpts, ref_like = [], []
        for q in Q:
            p = S.split_point(prob, q)
            comp = types.SimpleNamespace(
                name="geodetic", _like_name="geo_like", gfs=libs, slip_varnames=list(prob["slip_vars"]), Bij=Bij,
                sdata=geo["data"].view(_refshim._ND), sodws=geo["odw"].view(_refshim._ND), datasets=datasets, weights=list(geo["U"]),
                load_gfs=lambda **k: None, analyse_noise=lambda *a, **k: None, init_weights=lambda: None,
                get_gflibrary_key=lambda crust_ind, wavename, component: component,
                config=types.SimpleNamespace(dataset_specific_residual_noise_estimation=False,
                                             gf_config=types.SimpleNamespace(reference_model_idx=0, n_variations=(0, 1)),
                                             corrections_config=types.SimpleNamespace(has_enabled_corrections=False)))
            input_rvs = {v: np.array(p[v], dtype=np.float64) for v in prob["slip_vars"]}
            hyper = {"h_SAR%d" % d: np.float64(p["hypers"][geo["hyper_idx"][d]]) for d in range(len(nobs_list))}
            total = rgeodetic.GeodeticDistributerComposite.get_formula(comp, input_rvs, {}, hyper,
                                                                       types.SimpleNamespace(get_test_point=lambda: {}))
            ref_logpts.append(captured["geo_like"].copy())
            ref_like.append(float(total))
        ref_logpts = np.array(ref_logpts)
        for q, ref in zip(Q, ref_logpts):
            np.testing.assert_allclose(O.ffi_geodetic_eval(geo, S.split_point(prob, q)), ref, rtol=1e-10)
        np.testing.assert_allclose(ref_like, ref_logpts.sum(axis=1), rtol=1e-12)
        out[name + "_Q"], out[name + "_logpts"] = Q, ref_logpts
        print(name, ref_logpts.shape, float(ref_logpts.min()), float(ref_logpts.max()))
    np.savez_compressed(os.path.join(HERE, "geodetic_composite_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "geodetic_composite_golden.npz"))


if __name__ == "__main__":
    main()
