"""CPU tests of the geometry-mode oracle (oracle/geom_oracle.py) and of the host-side mirrors in beat_b200/geometry.py.

pyrocko is absent, so the synthesis part of the oracle is UNPINNED against the reference (see the oracle's header);
what can be checked here is checked: two independent published formulations of the double-couple moment tensor agree,
known-answer values for the geodesy and the STF, the algebraic properties of the delay-and-sum (linearity, shift
invariance on the time grid, constant continuation), taper/chop index rules, and that the llk is the golden-pinned one.
"""
import math

import numpy as np
import pytest

from beat_b200 import synthetic as S
from oracle import geom_oracle as O


def test_moment_tensor_two_formulations_agree():
    rng = np.random.default_rng(0)
    for _ in range(200):
        s, d, r = rng.uniform(0, 360), rng.uniform(0, 90), rng.uniform(-180, 180)
        np.testing.assert_allclose(O.dc_m6(s, d, r, 3.0), O.dc_m6_aki_richards(s, d, r, 3.0), atol=1e-14)
    # known answers: vertical strike-slip along north -> only m_ne; 45-degree thrust striking north -> mee = -mdd... (A&R 4.91)
    np.testing.assert_allclose(O.dc_m6(0.0, 90.0, 0.0), [0, 0, 0, 1, 0, 0], atol=1e-15)
    np.testing.assert_allclose(O.dc_m6(0.0, 45.0, 90.0), [0, -1, 1, 0, 0, 0], atol=1e-15)
    m = O.dc_m6(123.0, 37.0, -70.0)
    assert abs(m[0] + m[1] + m[2]) < 1e-15                                  # a double couple has no trace
    assert O.magnitude_to_moment(6.0) == pytest.approx(1.1220184543019653e18, rel=1e-12)


def test_stf_discretisation():
    t, a = O.halfsinusoid_discretize_t(0.0, -1.0, 0.5, 1.3)
    assert len(t) == 1 and a[0] == 1.0 and t[0] == pytest.approx(1.5)
    t, a = O.halfsinusoid_discretize_t(4.0, -1.0, 0.5, 0.0)
    assert len(t) == 9 and t[0] == 0.0 and t[-1] == 4.0
    assert a.sum() == pytest.approx(1.0) and np.all(a > 0)
    np.testing.assert_allclose(a, a[::-1], rtol=1e-12)                      # symmetric pulse on a symmetric grid
    # first bin integrates the half sinusoid from 0 to 0.25 s: (1 - cos(pi*0.25/4)) / 2
    assert a[0] == pytest.approx((1.0 - math.cos(math.pi * 0.25 / 4.0)) / 2.0, rel=1e-12)
    t, a = O.halfsinusoid_discretize_t(4.0, 0.0, 0.5, 10.0)
    assert t[0] == 8.0 and t[-1] == 12.0                                    # anchor 0: centred on the reference time


def test_boxcar_and_triangular_stf_discretisation():
    """The other two entries of the reference's stf_catalog (beat/sources.py:723-729): known answers of the restated
    [pyrocko] BoxcarSTF / TriangularSTF.discretize_t."""
    # piecewise-linear bin integrals: exact areas, jumps contribute nothing, end values continue
    e = np.linspace(-1.0, 4.0, 11)
    assert O.plf_integrate_piecewise(e, [0, 0, 2, 2], [0, 1, 1, 0]).sum() == pytest.approx(2.0, rel=1e-14)
    assert O.plf_integrate_piecewise(e, [0, 1.2, 3], [0, 1, 0]).sum() == pytest.approx(1.5, rel=1e-14)
    np.testing.assert_allclose(O.plf_integrate_piecewise([0.0, 0.5, 1.0], [0.0, 1.0], [0.0, 1.0]), [0.125, 0.375], rtol=1e-14)
    # boxcar on the grid: 4 s from 0 s at 0.5 s sampling = half-weight end points, centroid already in place
    t, a = O.boxcar_discretize_t(4.0, -1.0, 0.5, 0.0)
    assert a.sum() == pytest.approx(1.0) and (t * a).sum() == pytest.approx(2.0, abs=1e-12)
    k = np.flatnonzero(a > 1e-9)
    np.testing.assert_allclose(a[k], np.r_[0.5, np.ones(7), 0.5] / 8.0, rtol=1e-9)
    # off the grid: the sub-sample shift puts the discrete centroid on tref + duration / 2 (anchor -1)
    for dur, tref in ((2.3, 0.7), (0.2, -1.13), (7.77, 3.01)):
        t, a = O.boxcar_discretize_t(dur, -1.0, 0.5, tref)
        assert a.sum() == pytest.approx(1.0) and np.all(a >= 0)
        assert (t * a).sum() == pytest.approx(tref + 0.5 * dur, abs=1e-12)
        np.testing.assert_allclose(np.diff(t), 0.5, rtol=1e-12)
        np.testing.assert_allclose(t / 0.5, np.rint(t / 0.5), atol=1e-9)           # still on the store's time grid
    t, a = O.boxcar_discretize_t(0.0, -1.0, 0.5, 1.3)                               # a spike between two grid points
    np.testing.assert_allclose((t * a).sum(), 1.3, atol=1e-12)
    assert len(t) == 2
    # triangle: centroid ratio (1 + peak_ratio) / 3, symmetric for 0.5, onset at tref for anchor -1
    assert O.triangular_centroid_ratio(0.5) == pytest.approx(0.5)
    assert O.triangular_centroid_ratio(0.2) == pytest.approx(1.2 / 3.0)
    t, a = O.triangular_discretize_t(4.0, 0.5, -1.0, 0.5, 0.0)
    assert t[0] == 0.0 and t[-1] == 4.0 and a.sum() == pytest.approx(1.0)
    np.testing.assert_allclose(a, a[::-1], rtol=1e-12)
    assert a[0] == pytest.approx(0.5 * 0.25 * (0.25 / 2.0) / 2.0, rel=1e-12)        # area under the ramp up to 0.25 s / total area 2
    t, a = O.triangular_discretize_t(4.0, 0.25, -1.0, 0.5, 0.0)
    assert int(np.argmax(a)) == 2                                                    # apex at 1 s
    t, a = O.triangular_discretize_t(3.0, 0.5, 0.0, 0.5, 10.0)                      # anchor 0: centroid at tref
    assert (t * a).sum() == pytest.approx(10.0, abs=1e-12)
    for pr in (0.0, 1.0):                                                            # degenerate apex: a saw tooth
        t, a = O.triangular_discretize_t(2.0, pr, -1.0, 0.5, 0.3)
        assert a.sum() == pytest.approx(1.0) and np.all(a >= 0)


def test_stf_type_selects_the_catalogue_entry():
    """gprob['stf_type'] routes the sampled duration / peak_ratio like config.py:2058-2060 + utility.update_source."""
    src = dict(duration=3.0, time=0.4, peak_ratio=0.3)
    for kind, ref in (("HalfSinusoid", O.halfsinusoid_discretize_t(3.0, -1.0, 0.5, 0.4)), ("Boxcar", O.boxcar_discretize_t(3.0, -1.0, 0.5, 0.4)),
                      ("Triangular", O.triangular_discretize_t(3.0, 0.3, -1.0, 0.5, 0.4))):
        t, a = O.stf_discretize_t(dict(stf_type=kind, stf_anchor=-1.0), src, 0.5)
        np.testing.assert_array_equal(t, ref[0])
        np.testing.assert_array_equal(a, ref[1])
    with pytest.raises(ValueError):
        O.stf_discretize_t(dict(stf_type="Resonator", stf_anchor=-1.0), src, 0.5)
    gprob = S.make_geometry_problem(n_stations=2, stf_type="Triangular", seed=5)
    assert "peak_ratio" in gprob["offsets"] and gprob["n_params"] == len(S.GEOM_VARS) + 2
    q = S.draw_chains(gprob, 1, seed=2)[0]
    assert O.point_to_source(gprob, S.split_point(gprob, q))["peak_ratio"] == q[gprob["offsets"]["peak_ratio"]]


def test_geodesy_known_answers():
    lat, lon = O.ne_to_latlon(10.0, 20.0, 0.0, 0.0)
    assert (lat, lon) == pytest.approx((10.0, 20.0), abs=1e-9)
    lat, lon = O.ne_to_latlon(0.0, 0.0, 111194.92664455873, 0.0)           # one degree of arc on the 6371 km sphere
    assert lat == pytest.approx(1.0, abs=1e-9) and lon == pytest.approx(0.0, abs=1e-9)
    assert O.azimuth(0.0, 0.0, 1.0, 0.0) == pytest.approx(0.0, abs=1e-12)
    assert O.azimuth(0.0, 0.0, 0.0, 1.0) == pytest.approx(90.0, abs=1e-12)
    assert O.azimuth(0.0, 0.0, -1.0, 0.0) == pytest.approx(180.0, abs=1e-12)
    # one degree of longitude on the WGS84 equator: a * pi / 180
    assert O.distance_accurate50m(0.0, 0.0, 0.0, 1.0) == pytest.approx(6378140.0 * math.pi / 180.0, rel=1e-6)
    # a meridian degree at the equator is shorter by ~ (1 - e^2): 110.57 km
    assert O.distance_accurate50m(0.0, 0.0, 1.0, 0.0) == pytest.approx(110574.0, rel=2e-4)
    d, azi, bazi = O.source_receiver_geometry(30.0, 40.0, 3000.0, -4000.0, 30.0, 40.0)
    assert d == pytest.approx(5000.0) and bazi == pytest.approx(azi + 180.0)


def _tiny_store(rng, nrec=40):
    nz, nx = 2, 3
    tr = rng.standard_normal((nz, nx, 10, nrec)).astype(np.float32)
    return dict(deltat=0.5, nz=nz, nx=nx, z0=1000.0, dz=1000.0, x0=5000.0, dx=1000.0, traces=tr,
                itmin=rng.integers(-4, 4, (nz, nx, 10)).astype(np.int32), nsamples=rng.integers(30, nrec + 1, (nz, nx, 10)).astype(np.int32))


def test_store_sum_properties():
    rng = np.random.default_rng(1)
    st = _tiny_store(rng)
    e1 = [(0, 1, 3, 1.0, 0.7), (1, 2, 5, 2.5, -1.2)]
    e2 = [(1, 0, 9, 0.5, 2.0)]
    a = O.store_sum(st, e1, -10, 80)
    b = O.store_sum(st, e2, -10, 80)
    np.testing.assert_allclose(O.store_sum(st, e1 + e2, -10, 80), a + b, rtol=1e-6, atol=1e-6)        # linear
    # a delay of k samples shifts the output by k samples
    shifted = O.store_sum(st, [(iz, ix, g, d + 1.5, w) for iz, ix, g, d, w in e1], -10, 80)
    np.testing.assert_array_equal(shifted[3:], a[:-3])
    # constant continuation with the record's first / last value
    iz, ix, g = 0, 1, 3
    n = int(st["nsamples"][iz, ix, g]); it0 = int(st["itmin"][iz, ix, g])
    one = O.store_sum(st, [(iz, ix, g, 0.0, 1.0)], it0 - 5, n + 10)
    np.testing.assert_array_equal(one[:5], np.full(5, st["traces"][iz, ix, g, 0]))
    np.testing.assert_array_equal(one[5:5 + n], st["traces"][iz, ix, g, :n])
    np.testing.assert_array_equal(one[5 + n:], np.full(5, st["traces"][iz, ix, g, n - 1]))
    # a delay between two grid points is the linear mix of the neighbours
    half = O.store_sum(st, [(iz, ix, g, 0.25, 1.0)], it0, n)
    full0 = O.store_sum(st, [(iz, ix, g, 0.0, 1.0)], it0, n)
    full1 = O.store_sum(st, [(iz, ix, g, 0.5, 1.0)], it0, n)
    np.testing.assert_allclose(half, 0.5 * full0 + 0.5 * full1, rtol=1e-6, atol=1e-7)


def test_store_nodes():
    st = dict(z0=1000.0, dz=1000.0, nz=3, x0=5000.0, dx=1000.0, nx=4)
    assert O.store_nodes(st, 2000.0, 6000.0, "multilinear") == [(1, 1, 1.0)]                          # on a node: one node
    nodes = O.store_nodes(st, 1250.0, 6500.0, "multilinear")
    assert sorted(nodes) == [(0, 1, 0.375), (0, 2, 0.375), (1, 1, 0.125), (1, 2, 0.125)]
    assert O.store_nodes(st, 1500.0, 6500.0, "nearest_neighbor") == [(0, 2, 1.0)]                     # rint: ties to even
    with pytest.raises(O.OutOfBounds):
        O.store_nodes(st, 3500.0, 6000.0, "multilinear")


def test_component_weights_sensor_projection():
    m6 = O.dc_m6(30.0, 60.0, 45.0, 2.0)
    Wn = O.component_weights(m6, 40.0, 220.0, 0.0, 0.0)
    We = O.component_weights(m6, 40.0, 220.0, 90.0, 0.0)
    Wz = O.component_weights(m6, 40.0, 220.0, 0.0, -90.0)
    assert np.all(Wn[list(O.G_D)] == 0.0) and np.all(Wz[list(O.G_NE)] == 0.0)
    # a horizontal sensor at azimuth phi is cos(phi) N + sin(phi) E
    W30 = O.component_weights(m6, 40.0, 220.0, 30.0, 0.0)
    np.testing.assert_allclose(W30, math.cos(math.radians(30)) * Wn + math.sin(math.radians(30)) * We, atol=1e-15)
    # up = -down
    Wd = O.component_weights(m6, 40.0, 220.0, 0.0, 90.0)
    np.testing.assert_allclose(Wz, -Wd, atol=1e-15)


def test_taper_and_chop_rules():
    y = np.ones(71)
    O.cos_taper_inplace(y, 90.0, 0.5, 97.5, 100.0, 120.0, 122.5)
    assert np.all(y[:15] == 0.0) and y[15] == 0.0 and np.all(y[20:60] == 1.0) and np.all(y[65:] == 0.0)
    assert 0.0 < y[17] < 1.0 and 0.0 < y[62] < 1.0
    assert O.chop_indices(90.0, 0.5, 71, 100.0, 120.0) == (20, 60)          # [b, c): c itself excluded
    assert O.chop_indices(90.0, 0.5, 71, 100.2, 120.2) == (20, 60)          # floor snapping
    # between b and c the taper is exactly one: with chop bounds (b, c) it cannot change the likelihood window
    gprob = S.make_geometry_problem(n_stations=1, seed=3)
    wm = gprob["wavemaps"][0]
    raw = np.random.default_rng(2).standard_normal(71).astype(np.float32)
    itmin, n = O.target_window(wm, 0)
    assert n == 71
    wm_nofilt = dict(wm, filterer=[])
    out = O.post_process(wm_nofilt, 0, raw, itmin)
    ibeg = int(round((wm["arrival_times"][0] + wm["taper"][1]) / 0.5)) - itmin
    np.testing.assert_array_equal(out, raw[ibeg:ibeg + 40].astype(np.float64))


def test_filter_sections_are_what_scipy_runs():
    from scipy import signal
    from beat_b200.geometry import BandstopFilter, Filter
    secs = O.filter_sections([dict(kind="stepwise", order=4, lower_corner=0.01, upper_corner=0.4)], 0.5)
    assert len(secs) == 2 and secs[0][2] is True and secs[1][2] is False and len(secs[0][0]) == 5
    mine = Filter(0.01, 0.4, 4, stepwise=True).sections(0.5)
    for (b0, a0, d0), (b1, a1, d1) in zip(secs, mine):
        np.testing.assert_array_equal(b0, b1); np.testing.assert_array_equal(a0, a1); assert d0 == d1
    bp = Filter(0.02, 0.5, 4, stepwise=False).sections(0.5)
    assert len(bp) == 1 and len(bp[0][0]) == 9 and bp[0][2] is True
    bs = BandstopFilter(0.12, 0.25, 2).sections(0.5)
    assert len(bs) == 1 and len(bs[0][0]) == 5 and bs[0][2] is False
    # direct form II transposed restated in plain Python == scipy.signal.lfilter (what the CUDA kernel evaluates)
    x = np.random.default_rng(3).standard_normal(300)
    b, a, _ = secs[0]
    z = np.zeros(len(b) - 1)
    y = np.empty_like(x)
    for n, xn in enumerate(x):
        y[n] = b[0] * xn + z[0]
        for j in range(len(z) - 1):
            z[j] = b[j + 1] * xn + z[j + 1] - a[j + 1] * y[n]
        z[-1] = b[-1] * xn - a[-1] * y[n]
    np.testing.assert_allclose(y, signal.lfilter(b, a, x), rtol=1e-9, atol=1e-12)


def test_eval_is_deterministic_and_uses_the_pinned_llk():
    from oracle.ffi_oracle import mvn_chol_logpts
    gprob = S.make_geometry_problem(n_stations=2, seed=5)
    Q = S.draw_chains(gprob, 3, seed=6)
    p0 = S.split_point(gprob, Q[0])
    s0 = O.geometry_synthetics(gprob, p0)
    assert s0.shape == (6, 40) and np.all(np.isfinite(s0)) and np.abs(s0).max() > 0
    S.attach_geometry_data(gprob, s0)
    wm = gprob["wavemaps"][0]
    lp, synths = O.geometry_seismic_eval(gprob, S.split_point(gprob, Q[1]), return_synth=True)
    np.testing.assert_array_equal(lp, O.geometry_seismic_eval(gprob, S.split_point(gprob, Q[1])))
    ref = mvn_chol_logpts(wm["data"] - synths[0], wm["U"], wm["slog_pdet"], wm["nsamples"], S.split_point(gprob, Q[1])["hypers"][wm["hyper_idx"]])
    np.testing.assert_array_equal(lp, ref)
    # a source at the same place, twice the moment (magnitude + 0.2007) doubles the synthetics
    q2 = Q[0].copy()
    q2[gprob["offsets"]["magnitude"]] += math.log10(2.0) / 1.5
    s2 = O.geometry_synthetics(gprob, S.split_point(gprob, q2))
    np.testing.assert_allclose(s2, 2.0 * s0, rtol=2e-6, atol=1e-6 * np.abs(s0).max())


def test_host_mirrors():
    from beat_b200.geometry import ArrivalTaper
    t = ArrivalTaper()
    assert (t.a, t.b, t.c, t.d) == (-15.0, -10.0, 50.0, 55.0)              # heart.py:271-274 defaults
    assert t.nsamples(2.0) == 120 and t.nsamples(2.0, ("a", "d")) == 140 and t.fadein == 5.0 and t.fadeout == 5.0
    t.check_sample_rate_consistency(0.5)
    with pytest.raises(ValueError, match="inconsistent with sampling rate"):
        ArrivalTaper(-15.0, -10.0, 50.1, 55.0).check_sample_rate_consistency(0.5)
    with pytest.raises(ValueError, match="a < b < c < d"):
        ArrivalTaper(0.0, -1.0, 2.0, 3.0)


GOLDEN_CASES = {
    "stepwise_ml": (dict(n_stations=3, seed=201), ("b", "c")),
    "bandpass_nn": (dict(n_stations=2, seed=202, interpolation="nearest_neighbor",
                         filterer=[dict(kind="bandpass", order=3, lower_corner=0.02, upper_corner=0.5)]), ("b", "c")),
    "bandstop_ad": (dict(n_stations=2, seed=203, channels=("Z", "E"),
                         filterer=[dict(kind="stepwise", order=2, lower_corner=0.05, upper_corner=0.6),
                                   dict(kind="bandstop", order=2, lower_corner=0.12, upper_corner=0.25)]), ("a", "d")),
    "station_corr": (dict(n_stations=3, seed=204, station_corrections=True), ("b", "c")),
    "two_sources": (dict(n_stations=2, seed=205, n_sources=2), ("b", "c")),
}


def load_geometry_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "geometry_golden.npz"))


@pytest.mark.parametrize("name", sorted(GOLDEN_CASES))
def test_oracle_reproduces_reference_driven_golden(name):
    """tests/golden/geometry_golden.npz was produced by the reference's OWN heart.seis_synthetics / post_process_trace /
    Filter.apply / update_target_times (tests/golden/make_geometry_golden.py; pyrocko's Trace and engine replaced by
    stand-ins): the oracle's composition of window, filter, taper and chop must reproduce it exactly."""
    g = load_geometry_golden()
    kw, chop = GOLDEN_CASES[name]
    gprob = S.make_geometry_problem(**kw)
    wm = gprob["wavemaps"][0]
    Q, ref = g[name + "_Q"], g[name + "_synths"]
    for i, q in enumerate(Q):
        point = S.split_point(gprob, q)
        if chop == ("b", "c"):
            mine = O.geometry_synthetics(gprob, point)
        else:
            src = O.point_to_source(gprob, point)
            mine = np.vstack([O.post_process(wm, t, *O.seismogram(gprob, wm, t, src), chop_bounds=chop) for t in range(wm["nt"])])
        np.testing.assert_allclose(mine, ref[i], rtol=1e-12, atol=1e-18)
        shifts = point["time_shifts"][wm["station_idx"]] if wm.get("station_idx") is not None else 0.0
        np.testing.assert_allclose(g[name + "_tmins"][i], wm["arrival_times"] + shifts + dict(zip("abcd", wm["taper"]))[chop[0]])


def test_seis_synthesizer_op_host_logic_without_a_gpu():
    """The Op mirror's perform(): named inputs in any order, scalar or batched, optional per-target time_shift --
    checked against a fake context (the CUDA call itself is covered by the GPU tests)."""
    from beat_b200.geometry import ArrivalTaper, SeisSynthesizer
    from beat_b200.lib import GEOM_VARS

    class FakeCtx:
        def geom_synthetics_batch(self, wid, Q, nt, ns):
            self.Q = Q.copy()
            return np.broadcast_to(Q[:, :1, None], (Q.shape[0], nt, ns)).copy()

    for sc in (False, True):
        op = SeisSynthesizer.__new__(SeisSynthesizer)
        op.nt, op.ns, op.station_corrections = 3, 5, sc
        op._n_par = len(GEOM_VARS) + (3 if sc else 0)
        op.arrival_times = np.array([10.0, 20.0, 30.0])
        op.arrival_taper, op.chop_bounds = ArrivalTaper(-2.0, -1.0, 1.5, 2.5), ("b", "c")
        op._ctx, op._wid = FakeCtx(), 0
        vals = {v: float(i + 1) for i, v in enumerate(GEOM_VARS)}
        inputs = dict(reversed(list(vals.items())))                    # reversed order on purpose
        if sc:
            inputs["time_shift"] = np.array([0.5, -0.5, 0.25])
        synths, tmins = op(inputs)
        assert synths.shape == (3, 5) and tmins.shape == (3,)
        np.testing.assert_array_equal(op._ctx.Q[0, :9], np.arange(1.0, 10.0))
        shift = inputs.get("time_shift", 0.0)
        np.testing.assert_allclose(tmins, op.arrival_times + shift - 1.0)
        if sc:
            np.testing.assert_array_equal(op._ctx.Q[0, 9:], inputs["time_shift"])
        batch = {v: np.full(4, vals[v]) for v in GEOM_VARS}
        if sc:
            batch["time_shift"] = np.tile(inputs["time_shift"], (4, 1))
        sb, tb = op(batch)
        assert sb.shape == (4, 3, 5) and tb.shape == (4, 3)
        with pytest.raises(KeyError):
            op({k: v for k, v in inputs.items() if k != "depth"})
    assert op.infer_shape() == [(3, 5), (3,)]


OP_GOLDEN_CASES = {"op_two_sources": dict(n_stations=2, seed=206, n_sources=2),
                   "op_station_corr": dict(n_stations=3, seed=207, station_corrections=True)}


@pytest.mark.parametrize("name", sorted(OP_GOLDEN_CASES))
def test_oracle_matches_the_reference_op_perform(name):
    """Golden synthetics produced by the reference's OWN pytensorf.SeisSynthesizer.perform (beat/pytensorf.py:241-302:
    adjust_point_units, split_point, update_source, event-time offset, station corrections) -> heart.seis_synthetics,
    run with an absolute event time of 1e6 s (make_geometry_golden.py).  The oracle works relative to the event origin;
    the two agree to float32 rounding of the STF bin weights."""
    g = load_geometry_golden()
    gprob = S.make_geometry_problem(**OP_GOLDEN_CASES[name])
    for q, ref in zip(g[name + "_Q"], g[name + "_synths"]):
        mine = O.geometry_synthetics(gprob, S.split_point(gprob, q))
        np.testing.assert_allclose(mine, ref, rtol=2e-6, atol=2e-6 * np.abs(ref).max())
