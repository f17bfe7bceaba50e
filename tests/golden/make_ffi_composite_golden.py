"""
tests/golden/make_ffi_composite_golden.py -- regenerates tests/golden/ffi_composite_golden.npz.

Runs the REFERENCE'S OWN composite-level forward model of the finite-fault path -- the numpy twin of the graph that
``get_formula`` builds -- imported from /root/reference in the dev container:

    SeismicDistributerComposite.get_synthetics(point, outmode="array")      beat/models/seismic.py:1351-1507
      -> FaultGeometry.point2starttimes / fault_locations2idxs / get_subfault_starttimes   beat/ffi/fault.py:614-632,722-752,866-894
         -> fast_sweep.get_rupture_times_numpy                                beat/fast_sweeping/fast_sweep.py:67-230
      -> station corrections (:1408-1427), SeismicGFLibrary.stack_all per slip component (:1432-1459)

for seeded problems of ``beat_b200.synthetic.make_problem`` (one and two subfaults, with and without station
corrections, nearest-neighbour and multilinear) and stores chain parameters + synthetics.  Objects the method reads
from ``self`` that need the un-installed pyrocko / a project directory (config tree, wavemaps, the fault's geometry
bookkeeping) are attribute bags defined here; FaultGeometry's own methods, FaultOrdering, SeismicGFLibrary and the sweep
are the reference's.  These vectors pin ``oracle.ffi_oracle.ffi_seismic_eval(return_synth=True)`` at composite level.

    python tests/golden/make_ffi_composite_golden.py
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
    "one_fault_ml": dict(nt=4, subfaults=((4, 6, 2.0),), ns=24, ndur=4, seed=301, interpolation="multilinear"),
    "one_fault_nn_corr": dict(nt=5, subfaults=((3, 5, 2.5),), ns=20, ndur=3, seed=302, interpolation="nearest_neighbor",
                              station_corrections=True),
    "two_faults_ml_corr": dict(nt=3, subfaults=((3, 4, 2.0), (2, 5, 2.0)), ns=16, ndur=4, seed=303, interpolation="multilinear",
                               station_corrections=True),
}


def main():
    ext = O.load_reference_ext()
    _refshim.install(fast_sweep_ext=ext)
    from beat.config import SeismicGFLibraryConfig
    from beat.ffi import base as ffibase
    from beat.ffi import fault as rfault
    from beat.models import seismic as rseismic

    class Fault(rfault.FaultGeometry):
        """The reference's FaultGeometry methods on the few attributes they read (its constructor wants a config tree)."""

        def __init__(self, subfaults):
            self.ordering = rfault.FaultOrdering(npls=[s[1] for s in subfaults], npws=[s[0] for s in subfaults],
                                                 patch_sizes_strike=[s[2] for s in subfaults],
                                                 patch_sizes_dip=[s[2] for s in subfaults])
            self._nsf = len(subfaults)
            self._cum = np.concatenate([[0], np.cumsum([s[0] * s[1] for s in subfaults])]).astype(int)

        nsubfaults = property(lambda self: self._nsf)
        npatches = property(lambda self: int(self._cum[-1]))
        cum_subfault_npatches = property(lambda self: self._cum)

        def _check_index(self, index):
            assert 0 <= index < self._nsf

    out = {}
    for name, kw in CASES.items():
        prob = S.make_problem(**kw)
        wm = prob["wavemaps"][0]
        libs = {}
        for v in prob["slip_vars"]:
            cfg = SeismicGFLibraryConfig(dimensions=wm["G"][v].shape, starttime_min=wm["st_min"], starttime_sampling=wm["st_step"],
                                         duration_min=wm["dur_min"], duration_sampling=wm["dur_step"])
            lib = ffibase.SeismicGFLibrary(config=cfg)
            lib._gfmatrix = wm["G"][v]
            lib._tmins = np.zeros(wm["nt"])
            lib._stack_switch = {"numpy": wm["G"][v]}
            lib.set_stack_mode("numpy")
            libs[v] = lib
        corr = bool(prob.get("n_time_shifts"))
        wmap = types.SimpleNamespace(
            n_t=wm["nt"], _mapid="any_P_0", time_shifts_id="time_shifts_any_P_0", is_prepared=True, _prepared_data=[],
            station_correction_idxs=wm["station_idx"], targets=[None] * wm["nt"],
            config=types.SimpleNamespace(interpolation=wm["interpolation"], event_idx=0,
                                         arrival_taper=types.SimpleNamespace(nsamples=lambda sample_rate, ns=wm["ns"]: ns)))
        fault = object.__new__(Fault)            # FaultGeometry's pyrocko base class is a stand-in whose metaclass swallows arguments
        Fault.__init__(fault, prob["subfaults"])
        comp = types.SimpleNamespace(
            fault=fault, wavemaps=[wmap], gfs=libs, slip_varnames=list(prob["slip_vars"]),
            correction_name="time_shift", hierarchicals={},
            get_gflibrary_key=lambda crust_ind, wavename, component: component,
            config=types.SimpleNamespace(station_corrections=corr, get_hypernames=lambda: ["h_any_P_0_Z"],
                                         gf_config=types.SimpleNamespace(reference_model_idx=0, sample_rate=1.0 / prob["dt"])))
        Q = S.draw_chains(prob, 6, seed=500)
        synths = []
        for q in Q:
            p = S.split_point(prob, q)
            point = {k: np.array(v, dtype=np.float64) for k, v in p.items() if k not in ("hypers", "time_shifts")}
            if corr:
                point["time_shifts_any_P_0"] = np.array(p["time_shifts"])
            syn, _ = rseismic.SeismicDistributerComposite.get_synthetics(comp, point, outmode="array")
            synths.append(np.vstack(syn))
        synths = np.array(synths)
        # the oracle's restatement (compiled reference sweep + numpy stack) against the reference's numpy twin
        for q, ref in zip(Q, synths):
            _, mine, _ = O.ffi_seismic_eval(prob, S.split_point(prob, q), impl="auto", return_synth=True)
            np.testing.assert_allclose(mine[0], ref, rtol=1e-9, atol=1e-9 * np.abs(ref).max())
        out[name + "_Q"], out[name + "_synths"] = Q, synths
        print(name, synths.shape, float(np.abs(synths).max()))
    # ------------------------------------------------------------------------------------------------------------
    # The production graph itself: SeismicDistributerComposite.get_formula (beat/models/seismic.py:1210-1349), executed
    # eagerly through the numpy-backed pytensor shim: the reference's Sweeper Op -> compiled fast_sweep_ext, station
    # corrections with tt.tile / tt.repeat, stack_all in PYTENSOR mode (tt.batched_dot), residuals, and
    # multivariate_normal_chol -> per-dataset logpts ("seis_like") and their sum.
    from beat import pytensorf
    captured = {}

    def deterministic(name, var):
        captured[name] = np.asarray(var)
        return var
    rseismic.Deterministic = deterministic
    for name, kw in CASES.items():
        for hp_specific in (False, True):
            prob = S.make_problem(hp_specific=hp_specific, **kw)
            wm = prob["wavemaps"][0]
            libs = {}
            for v in prob["slip_vars"]:
                cfg = SeismicGFLibraryConfig(dimensions=wm["G"][v].shape, starttime_min=wm["st_min"], starttime_sampling=wm["st_step"],
                                             duration_min=wm["dur_min"], duration_sampling=wm["dur_step"])
                lib = ffibase.SeismicGFLibrary(config=cfg)
                lib._gfmatrix = wm["G"][v]
                lib._tmins = np.zeros(wm["nt"])
                lib._sgfmatrix = wm["G"][v].view(_refshim._ND)            # what init_optimization would share (ffi/base.py:387-404)
                lib.spatchidxs = lib.patchidxs
                lib._stack_switch = {"numpy": wm["G"][v], "pytensor": lib._sgfmatrix}
                lib.set_stack_mode("pytensor")
                lib.init_optimization = lambda: None
                libs[v] = lib
            corr = bool(prob.get("n_time_shifts"))
            fault = object.__new__(Fault)
            Fault.__init__(fault, prob["subfaults"])
            # datasets carry the log-determinant the way heart.Covariance does (shared scalar ``slog_pdet``, heart.py:130,247-253)
            datasets = [types.SimpleNamespace(samples=int(wm["nsamples"][t]), typ="any_P_0_Z",
                                              covariance=types.SimpleNamespace(slog_pdet=np.float64(wm["slog_pdet"][t])))
                        for t in range(wm["nt"])]
            Q = S.draw_chains(prob, 5, seed=600)
            logpts_ref, like_ref = [], []
            for q in Q:
                p = S.split_point(prob, q)
                wmap = types.SimpleNamespace(
                    n_t=wm["nt"], _mapid="any_P_0", time_shifts_id="time_shifts_any_P_0", station_correction_idxs=wm["station_idx"],
                    prepare_data=lambda **k: None, shared_data_array=wm["data"], datasets=datasets, weights=list(wm["U"]),
                    config=types.SimpleNamespace(interpolation=wm["interpolation"], event_idx=0, domain="time", name="any_P",
                                                 arrival_taper=types.SimpleNamespace(nsamples=lambda sample_rate, ns=wm["ns"]: ns)))
                comp = types.SimpleNamespace(
                    name="seismic", _like_name="seis_like", fault=fault, wavemaps=[wmap], gfs=libs, slip_varnames=list(prob["slip_vars"]),
                    sweepers=[pytensorf.Sweeper(h, nd, nstr, "c") for nd, nstr, h in prob["subfaults"]],
                    hierarchicals={"time_shifts_any_P_0": np.array(p["time_shifts"])} if corr else {},
                    events=[None], engine=None, load_gfs=lambda **k: None, analyse_noise=lambda *a, **k: None,
                    init_weights=lambda: None, get_all_station_names=lambda: [None] * wm["nt"],
                    get_gflibrary_key=lambda crust_ind, wavename, component: component,
                    config=types.SimpleNamespace(station_corrections=corr, dataset_specific_residual_noise_estimation=hp_specific,
                                                 gf_config=types.SimpleNamespace(reference_model_idx=0, sample_rate=1.0 / prob["dt"])))
                input_rvs = {k: np.array(v, dtype=np.float64) for k, v in p.items() if k not in ("hypers", "time_shifts")}
                hyper = {"h_any_P_0_Z": np.array(p["hypers"][wm["hyper_idx"]]) if hp_specific else np.float64(p["hypers"][wm["hyper_idx"][0]])}
                total = rseismic.SeismicDistributerComposite.get_formula(
                    comp, input_rvs, {}, hyper, types.SimpleNamespace(get_test_point=lambda: {}))
                logpts_ref.append(captured["seis_like"].copy())
                like_ref.append(float(total))
            logpts_ref = np.array(logpts_ref)
            for q, ref in zip(Q, logpts_ref):
                mine = O.ffi_seismic_eval(prob, S.split_point(prob, q), impl="ref")
                np.testing.assert_allclose(mine, ref, rtol=1e-10)
            np.testing.assert_allclose(like_ref, logpts_ref.sum(axis=1), rtol=1e-12)
            tag = name + ("_hps" if hp_specific else "")
            out["formula_" + tag + "_Q"], out["formula_" + tag + "_logpts"] = Q, logpts_ref
            print("formula", tag, logpts_ref.shape, float(logpts_ref.min()), float(logpts_ref.max()))
    np.savez_compressed(os.path.join(HERE, "ffi_composite_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "ffi_composite_golden.npz"))


if __name__ == "__main__":
    main()
