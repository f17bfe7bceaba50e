"""
tests/golden/make_geometry_golden.py -- regenerates tests/golden/geometry_golden.npz.

Runs the REFERENCE'S OWN ``beat.heart.seis_synthetics`` (beat/heart.py:3564-3762) -- and with it
``get_phase_taperer`` (:2590-2625), ``ArrivalTaper.get_pyrocko_taper`` (:316-336),
``DynamicTarget.update_target_times`` (:457-477), ``post_process_trace`` (:3466-3525) and ``Filter.apply`` /
``BandstopFilter.apply`` (:377-412) -- imported from /root/reference in the dev container, for seeded geometry-mode
problems, and stores inputs + outputs as golden vectors for ``oracle/geom_oracle.py``.

What is real and what is a stand-in: pyrocko is not installed, so the two pyrocko objects that code touches are
stand-ins defined HERE: ``pyrocko.trace.Trace`` / ``CosTaper`` (methods highpass / lowpass / bandpass / bandstop /
extend / taper / chop restated from pyrocko's published source; the filters are the scipy.signal.butter + lfilter calls
pyrocko makes) and an engine whose ``process()`` returns, for the window the REFERENCE code put on each target
(``target.tmin / tmax``), the oracle's restated pyrocko seismogram.  So these vectors pin the BEAT-side control flow
of the geometry-mode forward model (window, order of filter / extend / taper / chop, arguments the reference passes,
stacking, tmins) against the reference's source; they do NOT pin the restated pyrocko arithmetic against pyrocko.

    python tests/golden/make_geometry_golden.py
"""
import math
import os
import sys
import types
import warnings

import numpy as np
from scipy import signal

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _refshim  # noqa: E402
from beat_b200 import synthetic as S  # noqa: E402
from oracle import ffi_oracle  # noqa: E402
from oracle import geom_oracle as O  # noqa: E402

warnings.simplefilter("ignore")


# ---------------------------------------------------------------- stand-in pyrocko.trace
class NoData(Exception):
    pass


def t2ind(t, tdelta, snap=round):
    return int(snap(t / tdelta))


class Taper(_refshim.GutsObject):
    pass


class CosTaper(Taper):
    """pyrocko.trace.CosTaper: __call__(y, x0, dx) applies the raised-cosine flanks in place."""

    def __init__(self, a, b, c, d):
        Taper.__init__(self)
        self.a, self.b, self.c, self.d = a, b, c, d

    def __call__(self, y, x0, dx):
        O.cos_taper_inplace(y, x0, dx, self.a, self.b, self.c, self.d)


class Trace(object):
    """The part of pyrocko.trace.Trace that post_process_trace uses."""

    def __init__(self, tmin, deltat, ydata):
        self.tmin, self.deltat, self.ydata = float(tmin), float(deltat), ydata
        self.calls = []

    @property
    def tmax(self):
        return self.tmin + (self.ydata.size - 1) * self.deltat

    def data_len(self):
        return self.ydata.size

    def copy(self):
        return Trace(self.tmin, self.deltat, self.ydata.copy())

    def _filter(self, btype, corners, order, demean):
        b, a = signal.butter(order, [c * 2.0 * self.deltat for c in corners], btype=btype)
        data = self.ydata.astype(np.float64)
        if demean:
            data -= np.mean(data)
        self.ydata = signal.lfilter(b, a, data)

    def highpass(self, order, corner, nyquist_warn=True, nyquist_exception=False, demean=True):
        self.calls.append(("highpass", order, corner, demean))
        self._filter("high", [corner], order, demean)

    def lowpass(self, order, corner, nyquist_warn=True, nyquist_exception=False, demean=True):
        self.calls.append(("lowpass", order, corner, demean))
        self._filter("low", [corner], order, demean)

    def bandpass(self, order, corner_hp, corner_lp, demean=True):
        self.calls.append(("bandpass", order, corner_hp, corner_lp, demean))
        self._filter("band", [corner_hp, corner_lp], order, demean)

    def bandstop(self, order, corner_hp, corner_lp, demean=True):
        self.calls.append(("bandstop", order, corner_hp, corner_lp, demean))
        self._filter("bandstop", [corner_hp, corner_lp], order, demean)

    def extend(self, tmin=None, tmax=None, fillmethod="zeros"):
        self.calls.append(("extend", fillmethod))
        nl = max(0, int(round((self.tmin - tmin) / self.deltat))) if tmin is not None else 0
        nh = max(0, int(round((tmax - self.tmax) / self.deltat))) if tmax is not None else 0
        if nl or nh:
            assert fillmethod == "zeros"
            self.ydata = np.concatenate((np.zeros(nl, self.ydata.dtype), self.ydata, np.zeros(nh, self.ydata.dtype)))
            self.tmin -= nl * self.deltat

    def taper(self, taperer, inplace=True, chop=False):
        self.calls.append(("taper", inplace))
        assert inplace
        taperer(self.ydata, self.tmin, self.deltat)

    def chop(self, tmin, tmax, inplace=True, include_last=False, snap=(round, round), want_incomplete=True):
        self.calls.append(("chop", snap[0].__name__, snap[1].__name__, include_last))
        if tmax <= self.tmin or self.tmax < tmin:
            raise NoData()
        ibeg = max(0, t2ind(tmin - self.tmin, self.deltat, snap[0]))
        iend = min(self.data_len(), t2ind(tmax - self.tmin, self.deltat, snap[1]) + (1 if include_last else 0))
        if ibeg >= iend:
            raise NoData()
        self.tmin += ibeg * self.deltat
        self.ydata = self.ydata[ibeg:iend].copy()
        return self


def make_trace_module():
    m = _refshim._StubModule("pyrocko.trace")
    m.Taper, m.CosTaper, m.Trace, m.NoData, m.t2ind = Taper, CosTaper, Trace, NoData, t2ind
    m.PoleZeroResponse = type("PoleZeroResponse", (_refshim.GutsObject,), {})
    return m


class GfTarget(_refshim.GutsObject):
    """Stand-in base class for heart.DynamicTarget (pyrocko.gf.Target): plain attribute bag."""
    tmin = None
    tmax = None


# ---------------------------------------------------------------- stand-in engine
class Engine(object):
    """process() = the oracle's restated pyrocko synthesis over the window the reference put on each target.
    ``t_event``: absolute origin time of the reference event (a multiple of deltat); the oracle works relative to it."""

    def __init__(self, gprob, wm, t_event=0.0):
        self.gprob, self.wm, self.t_event = gprob, wm, t_event

    def get_store(self, store_id):
        return types.SimpleNamespace(config=types.SimpleNamespace(sample_rate=1.0 / self.gprob["store"]["deltat"]))

    def close_cashed_stores(self):
        pass

    def process(self, sources, targets, nthreads=1):
        gprob, wm, dt = self.gprob, self.wm, self.gprob["store"]["deltat"]
        i_event = int(round(self.t_event / dt))
        results = []
        for src_obj in sources:
            params = src_obj.params if hasattr(src_obj, "params") else src_obj.as_oracle_source(self.t_event)
            for t, target in enumerate(targets):
                # pyrocko seismosizer: itmin = floor(tmin/deltat), nsamples = ceil(tmax/deltat) - itmin + 1
                itmin = int(math.floor(target.tmin / dt))
                n = int(math.ceil(target.tmax / dt)) - itmin + 1
                assert (itmin - i_event, n) == O.target_window(wm, t), "reference window != oracle window"
                raw, it0 = O.seismogram(gprob, wm, t, params)
                assert it0 == itmin - i_event and raw.size == n
                results.append((src_obj, target, Trace(itmin * dt, dt, raw)))
        return types.SimpleNamespace(iter_results=lambda: iter(results))


class FakeSTF(dict):
    """pyrocko STF stand-in: utility.update_source writes ``source.stf[k] = v`` for non-source attributes."""


class FakeSource(object):
    """pyrocko gf.DCSource stand-in with the mapping interface utility.update_source uses (keys / __setitem__ / stf)."""
    _keys = ("east_shift", "north_shift", "depth", "strike", "dip", "rake", "magnitude", "time")

    def __init__(self):
        self.stf = FakeSTF(duration=0.0)
        for k in self._keys:
            setattr(self, k, 0.0)

    def keys(self):
        return list(self._keys)

    def __setitem__(self, k, v):
        assert k in self._keys
        setattr(self, k, v)

    def as_oracle_source(self, t_event):
        d = {k: getattr(self, k) for k in self._keys}           # already in metres (adjust_point_units ran)
        d["time"] = self.time - t_event                           # the oracle's times are relative to the event origin
        d["duration"] = self.stf["duration"]
        return d


class FakeMapping(object):
    """beat.config.DatatypeParameterMapping stand-in (utility.split_point, beat/utility.py:678-733)."""

    def __init__(self, n_sources):
        self.n_sources = n_sources
        self._names = FakeSource._keys + ("duration",)

    def point_to_sources_mapping(self):
        return {k: list(range(self.n_sources)) for k in self._names}

    def point_variable_names(self):
        return list(self._names)


def main():
    ext = ffi_oracle.load_reference_ext()
    finder = _refshim.install(fast_sweep_ext=ext)
    import pyrocko  # the stub package
    trace_mod = make_trace_module()
    finder.fixed["pyrocko.trace"] = trace_mod
    sys.modules["pyrocko.trace"] = trace_mod
    pyrocko.trace = trace_mod
    import pyrocko.gf as gf
    gf.Target = GfTarget
    from beat import heart  # noqa: E402  (the reference's own module)

    assert heart.trace is trace_mod
    out = {}
    cases = [
        dict(name="stepwise_ml", kw=dict(n_stations=3, seed=201), chop=("b", "c")),
        dict(name="bandpass_nn", kw=dict(n_stations=2, seed=202, interpolation="nearest_neighbor",
                                         filterer=[dict(kind="bandpass", order=3, lower_corner=0.02, upper_corner=0.5)]), chop=("b", "c")),
        dict(name="bandstop_ad", kw=dict(n_stations=2, seed=203, channels=("Z", "E"),
                                         filterer=[dict(kind="stepwise", order=2, lower_corner=0.05, upper_corner=0.6),
                                                   dict(kind="bandstop", order=2, lower_corner=0.12, upper_corner=0.25)]), chop=("a", "d")),
        # station corrections: SeisSynthesizer.perform hands arrival_times + time_shifts to seis_synthetics (pytensorf.py:248-252)
        dict(name="station_corr", kw=dict(n_stations=3, seed=204, station_corrections=True), chop=("b", "c")),
        # two sources: seis_synthetics post-processes every (source, target) trace and stacks them (heart.py:3719-3724)
        dict(name="two_sources", kw=dict(n_stations=2, seed=205, n_sources=2), chop=("b", "c")),
    ]
    for case in cases:
        gprob = S.make_geometry_problem(**case["kw"])
        wm = gprob["wavemaps"][0]
        a, b, c, d = wm["taper"]
        ataper = heart.ArrivalTaper(a=a, b=b, c=c, d=d)
        ataper.check_sample_rate_consistency(wm["deltat"])
        filterer = []
        for f in wm["filterer"]:
            if f["kind"] == "bandstop":
                filterer.append(heart.BandstopFilter(lower_corner=f["lower_corner"], upper_corner=f["upper_corner"], order=f["order"]))
            else:
                filterer.append(heart.Filter(lower_corner=f["lower_corner"], upper_corner=f["upper_corner"], order=f["order"],
                                             stepwise=f["kind"] == "stepwise"))
        targets = [heart.DynamicTarget(lat=wm["lats"][t], lon=wm["lons"][t], azimuth=wm["azimuths"][t], dip=wm["dips"][t],
                                       store_id="synthetic") for t in range(wm["nt"])]
        assert all(tg.response is None for tg in targets)
        Q = S.draw_chains(gprob, 4, seed=300 + len(out))
        synths_ref, tmins_ref = [], []
        for q in Q:
            point = S.split_point(gprob, q)
            arrival_times = np.array(wm["arrival_times"])
            if wm.get("station_idx") is not None:
                arrival_times = arrival_times + point["time_shifts"][wm["station_idx"]]
            engine = Engine(gprob, dict(wm, arrival_times=arrival_times))
            srcs = [types.SimpleNamespace(params=sp) for sp in O.point_to_sources(gprob, point)]
            synths, tmins = heart.seis_synthetics(
                engine=engine, sources=srcs, targets=targets, arrival_taper=ataper, wavename="any_P", filterer=filterer,
                pre_stack_cut=True, arrival_times=arrival_times, outmode="array", chop_bounds=list(case["chop"]))
            synths_ref.append(synths)
            tmins_ref.append(tmins)
        synths_ref, tmins_ref = np.array(synths_ref), np.array(tmins_ref)
        # the oracle's own composition must reproduce what the reference's control flow produced
        for i, q in enumerate(Q):
            if case["chop"] == ("b", "c"):
                mine = O.geometry_synthetics(gprob, S.split_point(gprob, q))
            else:
                srcp = O.point_to_source(gprob, S.split_point(gprob, q))
                mine = np.vstack([O.post_process(wm, t, *O.seismogram(gprob, wm, t, srcp), chop_bounds=case["chop"]) for t in range(wm["nt"])])
            np.testing.assert_array_equal(mine, synths_ref[i])
        if wm.get("station_idx") is None:
            np.testing.assert_allclose(tmins_ref[0], wm["arrival_times"] + dict(a=a, b=b, c=c, d=d)[case["chop"][0]])
        assert ataper.nsamples(1.0 / wm["deltat"], list(case["chop"])) == synths_ref.shape[2]
        n = case["name"]
        out[n + "_Q"], out[n + "_synths"], out[n + "_tmins"] = Q, synths_ref, tmins_ref
        print(n, synths_ref.shape, float(np.abs(synths_ref).max()))
    # ---- the reference Op itself: pytensorf.SeisSynthesizer.perform (beat/pytensorf.py:241-302) with an absolute event
    # time, km -> m units (utility.adjust_point_units), split_point / update_source, station corrections
    from beat import pytensorf
    t_event = 1.0e6                                              # a multiple of deltat
    for name, kw in (("op_two_sources", dict(n_stations=2, seed=206, n_sources=2)),
                     ("op_station_corr", dict(n_stations=3, seed=207, station_corrections=True))):
        gprob = S.make_geometry_problem(**kw)
        wm = gprob["wavemaps"][0]
        a, b, c, d = wm["taper"]
        n_src = gprob["n_sources"]
        targets = [heart.DynamicTarget(lat=wm["lats"][t], lon=wm["lons"][t], azimuth=wm["azimuths"][t], dip=wm["dips"][t],
                                       store_id="synthetic") for t in range(wm["nt"])]
        f = wm["filterer"][0]
        Q = S.draw_chains(gprob, 3, seed=400)
        got = []
        for q in Q:
            point = S.split_point(gprob, q)
            shifts = point["time_shifts"][wm["station_idx"]] if wm.get("station_idx") is not None else None
            arr_rel = np.array(wm["arrival_times"]) + (shifts if shifts is not None else 0.0)
            engine = Engine(gprob, dict(wm, arrival_times=arr_rel), t_event=t_event)
            op = pytensorf.SeisSynthesizer(
                engine=engine, sources=[FakeSource() for _ in range(n_src)], mapping=FakeMapping(n_src), targets=targets,
                events=[types.SimpleNamespace(time=t_event)], event_idx=0, arrival_taper=heart.ArrivalTaper(a=a, b=b, c=c, d=d),
                arrival_times=np.array(wm["arrival_times"]) + t_event, wavename="any_P",
                filterer=[heart.Filter(lower_corner=f["lower_corner"], upper_corner=f["upper_corner"], order=f["order"])],
                pre_stack_cut=True, station_corrections=shifts is not None, domain="time")
            inputs = {k: np.atleast_1d(v) for k, v in point.items() if k not in ("hypers", "time_shifts")}
            if shifts is not None:
                inputs["time_shift"] = shifts
            op.varnames = list(inputs.keys())
            output = [[None], [None]]
            op.perform(None, list(inputs.values()), output)
            synths, tmins = output[0][0], output[1][0]
            assert op.infer_shape() == [synths.shape, (wm["nt"],)]
            # not bit-equal: (t_event + t) - t_event differs from t by ~1e-10 s, which moves STF bin weights in the last
            # float32 digit -- the oracle's relative-time convention against the reference's absolute one
            mine = O.geometry_synthetics(gprob, point)
            np.testing.assert_allclose(synths, mine, rtol=2e-6, atol=2e-6 * np.abs(mine).max())
            np.testing.assert_allclose(np.asarray(tmins) - t_event, arr_rel + b, atol=1e-6)
            got.append(synths)
        out[name + "_Q"], out[name + "_synths"] = Q, np.array(got)
        print(name, np.array(got).shape)

    # the sequence of trace operations the reference issued for the default (stepwise) filter, for the record
    tr = Trace(0.0, 0.5, np.random.default_rng(0).standard_normal(200).astype(np.float32))
    heart.post_process_trace(tr, CosTaper(20.0, 25.0, 60.0, 65.0), [heart.Filter(lower_corner=0.01, upper_corner=0.4, order=4)],
                             chop_bounds=["b", "c"])
    print(tr.calls)
    assert tr.calls == [("highpass", 4, 0.01, True), ("lowpass", 4, 0.4, False), ("extend", "zeros"), ("taper", True),
                        ("chop", "floor", "floor", False)]
    np.savez_compressed(os.path.join(HERE, "geometry_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "geometry_golden.npz"))


if __name__ == "__main__":
    main()
