"""GPU parity tests of the geometry-mode (point-source) seismic path -- BASELINE config 2, SURVEY rows a12 / f5 --
through the C-ABI (beatgpu_geom_*), against oracle/geom_oracle.py on the same seeded inputs.

Tolerances: synthetics ``atol = rtol = 5e-6`` relative to the largest synthetic amplitude (the reference's own
tolerance for stacked synthetics, test/test_ffi_gfstacking.py:49-58; the GF sum accumulates in float32 like pyrocko's
store, so summation order shows at the 1e-7 level); log-likelihoods ``rtol 1e-5`` (north-star tolerance).
"""
import numpy as np
import pytest

from beat_b200 import synthetic as S

pytestmark = pytest.mark.gpu


def _oracle():
    from oracle import geom_oracle
    return geom_oracle


def _engine(gprob, **kw):
    from beat_b200.geometry import BatchedGeometryLogLike
    return BatchedGeometryLogLike.from_problem(gprob, device=0, **kw)


def _points(gprob, Q):
    return [S.split_point(gprob, q) for q in Q]


def _with_data(gprob, noise="exponential"):
    O = _oracle()
    q0 = S.draw_chains(gprob, 1, seed=1)[0]
    S.attach_geometry_data(gprob, O.geometry_synthetics(gprob, S.split_point(gprob, q0)), noise=noise)
    return gprob


def _assert_logpts_close(gprob, Q, logpts, ref):
    """rtol 1e-5 on the log-likelihood, where 'relative' refers to the quadratic form the path computes: a logpt is
    -(const + quad)/2 with |const| in the hundreds, and can pass through zero for some chain, so a plain relative test on
    logpt itself would ask for more digits than the north-star tolerance does."""
    wm = gprob["wavemaps"][0]
    h = Q[:, gprob["offsets"]["hypers"]:gprob["offsets"]["hypers"] + gprob["n_hypers"]][:, wm["hyper_idx"]]
    const = wm["slog_pdet"][None, :] + wm["nsamples"][None, :] * (2.0 * h + np.log(2.0 * np.pi))
    np.testing.assert_allclose(-2.0 * logpts - const, -2.0 * ref - const, rtol=1e-5)
    np.testing.assert_allclose(logpts, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())


def _assert_synth_close(got, ref):
    scale = np.abs(ref).max()
    np.testing.assert_allclose(got, ref, rtol=5e-6, atol=5e-6 * scale)


@pytest.mark.parametrize("interpolation", ["multilinear", "nearest_neighbor"])
def test_synthetics_match_oracle(interpolation):
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=4, interpolation=interpolation, seed=21)
    Q = S.draw_chains(gprob, 24, seed=3)
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    ev.close()
    ref = np.array([O.geometry_synthetics(gprob, p) for p in _points(gprob, Q)])
    assert got.shape == ref.shape == (24, 12, 40)
    _assert_synth_close(got, ref)


@pytest.mark.parametrize("noise", ["exponential", "variance", "dense"])
def test_loglike_matches_oracle(noise):
    """exponential -> bidiagonal weights (fused streaming misfit), variance -> diagonal, dense -> residuals + DMMA GEMM."""
    O = _oracle()
    gprob = _with_data(S.make_geometry_problem(n_stations=3, seed=31), noise=noise)
    Q = S.draw_chains(gprob, 40, seed=4)
    Q[0] = S.draw_chains(gprob, 1, seed=1)[0]                  # the point the data were made from (smallest residuals)
    ev = _engine(gprob)
    logpts, like = ev(Q)
    ev.close()
    ref = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Q)])
    _assert_logpts_close(gprob, Q, logpts, ref)
    np.testing.assert_allclose(like, ref.sum(axis=1), rtol=1e-5)


def test_hp_specific_and_single_channel():
    O = _oracle()
    gprob = _with_data(S.make_geometry_problem(n_stations=5, channels=("Z",), hp_specific=True, seed=41))
    Q = S.draw_chains(gprob, 16, seed=5)
    ev = _engine(gprob)
    logpts, like = ev(Q)
    one = ev.logp_forw_func(Q[3])
    ev.close()
    ref = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Q)])
    np.testing.assert_allclose(logpts, ref, rtol=1e-5)
    np.testing.assert_allclose(one[0], ref[3], rtol=1e-5)
    np.testing.assert_allclose(one[1], ref[3].sum(), rtol=1e-5)


def test_stf_edge_cases():
    """duration 0 (one STF point), durations that put tmin/tmax exactly between grid points, long STF."""
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=2, duration_bounds=(0.0, 20.0), seed=51)
    Q = S.draw_chains(gprob, 12, seed=6)
    od, ot = gprob["offsets"]["duration"], gprob["offsets"]["time"]
    Q[0, od] = 0.0
    Q[1, od], Q[1, ot] = 0.25, 0.25                             # tmin_stf/deltat = 0.5: rint ties to even
    Q[2, od], Q[2, ot] = 0.5, -0.75
    Q[3, od] = 20.0
    Q[4, od], Q[4, ot] = 1e-9, 1.0
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    ev.close()
    ref = np.array([O.geometry_synthetics(gprob, p) for p in _points(gprob, Q)])
    _assert_synth_close(got, ref)


@pytest.mark.parametrize("stf_type,n_sources,sample_peak", [("Boxcar", 1, True), ("Triangular", 1, True), ("Triangular", 1, False),
                                                          ("Boxcar", 2, True), ("Triangular", 2, True)])
def test_boxcar_and_triangular_stf(stf_type, n_sources, sample_peak):
    """The other two entries of the reference's stf_catalog (beat/sources.py:723-729; SeismicGeometryConfig.stf_type):
    synthetics and log-likelihoods against the oracle, `peak_ratio` of the triangle sampled per chain (and per source) or fixed,
    durations from a spike to 20 s, on and off the grid."""
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=2, duration_bounds=(0.0, 20.0), seed=53, stf_type=stf_type, n_sources=n_sources,
                                    sample_peak_ratio=sample_peak)
    if stf_type == "Triangular" and not sample_peak:
        gprob["peak_ratio"] = 0.3
    Q = S.draw_chains(gprob, 16, seed=8)
    od, ot = gprob["offsets"]["duration"], gprob["offsets"]["time"]
    Q[0, od] = 0.0
    Q[1, od], Q[1, ot] = 0.25, 0.25
    Q[2, od], Q[2, ot] = 4.0, 0.0                                # on the grid: whole-sample centroid shift of the boxcar
    Q[3, od] = 20.0
    if "peak_ratio" in gprob["offsets"]:
        Q[4, gprob["offsets"]["peak_ratio"]], Q[5, gprob["offsets"]["peak_ratio"]] = 0.0, 1.0
    _with_data(gprob)
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    logpts, like = ev(Q)
    ev.close()
    ref = np.array([O.geometry_synthetics(gprob, p) for p in _points(gprob, Q)])
    _assert_synth_close(got, ref)
    ref_lp = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Q)])
    _assert_logpts_close(gprob, Q, logpts, ref_lp)
    # the STF type matters: the half-sinusoid engine gives different synthetics for the same chains
    base = dict(gprob, stf_type="HalfSinusoid")
    ev = _engine(base)
    other = ev.get_synthetics(Q)
    ev.close()
    assert np.abs(other[3] - got[3]).max() > 1e-3 * np.abs(got[3]).max()


def test_stf_argument_errors():
    from beat_b200.lib import Context
    gprob = S.make_geometry_problem(n_stations=2, seed=54)
    ev = _engine(gprob)
    with pytest.raises(ValueError):
        ev.ctx.geom_set_stf("Resonator")
    with pytest.raises(ValueError):
        ev.ctx.geom_set_stf("Triangular", -1.0, -1, 1.5)           # fixed peak_ratio outside [0, 1]
    with pytest.raises(ValueError):
        ev.ctx.geom_set_stf("Triangular", -1.0, gprob["n_params"], 0.5)   # column outside q
    with pytest.raises(ValueError):
        ev.ctx.geom_set_stf("Boxcar", 2.0)                        # anchor outside [-1, 1]
    ev.close()
    c = Context(0)
    with pytest.raises(Exception):
        c.geom_set_stf("Boxcar")                                 # before geom_set_source
    c.close()


@pytest.mark.parametrize("filterer", [
    [],
    [dict(kind="bandpass", order=4, lower_corner=0.02, upper_corner=0.5)],
    [dict(kind="stepwise", order=2, lower_corner=0.05, upper_corner=0.3)],
    [dict(kind="bandstop", order=2, lower_corner=0.12, upper_corner=0.25)],
    [dict(kind="stepwise", order=3, lower_corner=0.02, upper_corner=0.6), dict(kind="bandstop", order=2, lower_corner=0.12, upper_corner=0.25)],
])
def test_filters(filterer):
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=2, filterer=filterer, seed=61)
    Q = S.draw_chains(gprob, 8, seed=7)
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    ev.close()
    ref = np.array([O.geometry_synthetics(gprob, p) for p in _points(gprob, Q)])
    _assert_synth_close(got, ref)


def test_chop_a_d_applies_the_taper_flanks():
    """chop bounds (a, d) keep the raised-cosine flanks (taper_filter_traces' plotting mode, heart.py:4242-4318)."""
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=2, ns=40, seed=71)
    wm = gprob["wavemaps"][0]
    wm["chop_bounds"] = ("a", "d")
    wm["ns"] = 50                                               # (17.5 + 7.5) s at 2 Hz
    Q = S.draw_chains(gprob, 6, seed=8)
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    ev.close()
    ref = []
    for p in _points(gprob, Q):
        src = O.point_to_source(gprob, p)
        rows = []
        for t in range(wm["nt"]):
            raw, itmin = O.seismogram(gprob, wm, t, src)
            rows.append(O.post_process(wm, t, raw, itmin, chop_bounds=("a", "d")))
        ref.append(np.vstack(rows))
    ref = np.array(ref)
    assert got.shape == ref.shape
    assert np.all(ref[:, :, 0] == 0.0) and np.any(ref[:, :, 2] != 0.0)
    _assert_synth_close(got, ref)


def test_source_outside_store_raises_and_marks_nan():
    gprob = _with_data(S.make_geometry_problem(n_stations=2, seed=81))
    Q = S.draw_chains(gprob, 8, seed=9)
    Q[5, gprob["offsets"]["depth"]] = 500.0                     # km: far below the store's deepest source
    ev = _engine(gprob)
    with pytest.raises(IndexError):
        ev(Q)
    import torch
    q_dev = torch.from_numpy(Q).cuda()
    logpts, like = ev.eval_device(q_dev)
    torch.cuda.synchronize()
    ev.ctx.index_violations()                                   # reset the counter
    lp = logpts.cpu().numpy()
    assert np.all(np.isnan(lp[5])) and np.all(np.isfinite(np.delete(lp, 5, axis=0)))
    ev.close()


def test_device_entry_matches_host_entry_and_fixed_variables():
    import torch
    O = _oracle()
    gprob = _with_data(S.make_geometry_problem(n_stations=3, seed=91))
    Q = S.draw_chains(gprob, 32, seed=10)
    ev = _engine(gprob)
    lp_host, like_host = ev(Q)
    lp_dev, like_dev = ev.eval_device(torch.from_numpy(Q).cuda())
    torch.cuda.synchronize()
    np.testing.assert_array_equal(lp_dev.cpu().numpy(), lp_host)
    np.testing.assert_array_equal(like_dev.cpu().numpy(), like_host)
    ev.close()
    # same problem with strike and the hyperparameter held fixed (offset -1 -> value from the fixed vector)
    fixed = np.zeros(9 + gprob["n_hypers"])
    fixed[3], fixed[9] = 77.0, 1.25
    keep = [v for v, _ in gprob["var_order"] if v not in ("strike", "hypers")]
    g2 = dict(gprob)
    g2["offsets"] = {v: i for i, v in enumerate(keep)}
    g2["n_params"] = len(keep)
    g2["fixed"] = fixed
    Q2 = np.ascontiguousarray(Q[:, [gprob["offsets"][v] for v in keep]])
    ev2 = _engine(g2)
    lp2, _ = ev2(Q2)
    ev2.close()
    Qf = Q.copy()
    Qf[:, gprob["offsets"]["strike"]] = 77.0
    Qf[:, gprob["offsets"]["hypers"]] = 1.25
    ref = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Qf)])
    np.testing.assert_allclose(lp2, ref, rtol=1e-5)


def test_seis_synthesizer_op_single_point_and_batch():
    """Calling convention of the reference Op (beat/pytensorf.py:215-302): dict of named variables in,
    (synthetics, tmins) out."""
    from beat_b200.geometry import ArrivalTaper, SeisSynthesizer
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=2, seed=101)
    wm = gprob["wavemaps"][0]
    op = SeisSynthesizer(gprob["store"], gprob["event"], dict(lats=wm["lats"], lons=wm["lons"], azimuths=wm["azimuths"], dips=wm["dips"]),
                         ArrivalTaper(*wm["taper"]), wm["arrival_times"], wm["filterer"], pre_stack_cut=True,
                         interpolation=wm["interpolation"])
    assert op.infer_shape() == [(6, 40), (6,)]
    Q = S.draw_chains(gprob, 5, seed=11)
    p = S.split_point(gprob, Q[0])
    synths, tmins = op({k: v for k, v in p.items() if k != "hypers"})
    _assert_synth_close(synths, O.geometry_synthetics(gprob, p))
    np.testing.assert_allclose(tmins, wm["arrival_times"] + wm["taper"][1])
    batch = {v: Q[:, gprob["offsets"][v]] for v, _ in gprob["var_order"] if v != "hypers"}
    sb, tb = op(batch)
    assert sb.shape == (5, 6, 40) and tb.shape == (5, 6)
    _assert_synth_close(sb[0], synths)
    op.close()
    with pytest.raises(NotImplementedError):
        SeisSynthesizer(gprob["store"], gprob["event"], dict(lats=wm["lats"], lons=wm["lons"], azimuths=wm["azimuths"], dips=wm["dips"]),
                        ArrivalTaper(*wm["taper"]), wm["arrival_times"], wm["filterer"], pre_stack_cut=False)


def test_argument_errors():
    from beat_b200.lib import Context
    gprob = S.make_geometry_problem(n_stations=1, seed=111)
    ev = _engine(gprob)
    with pytest.raises(Exception, match="not uploaded"):
        ev(S.draw_chains(gprob, 2))
    with pytest.raises(ValueError):
        ev(np.zeros((2, 3)))
    # a wavemap whose declared sample count does not match the taper
    wm = gprob["wavemaps"][0]
    with pytest.raises(ValueError, match="chops to"):
        ev.ctx.geom_add_wavemap(ev.store_id, 39, "multilinear", wm["lats"], wm["lons"], wm["azimuths"], wm["dips"], wm["arrival_times"],
                                wm["taper"], ("b", "c"), [], wm["hyper_idx"], wm["nsamples"])
    with pytest.raises(ValueError, match="a < b < c < d"):
        ev.ctx.geom_add_wavemap(ev.store_id, 40, "multilinear", wm["lats"], wm["lons"], wm["azimuths"], wm["dips"], wm["arrival_times"],
                                (0.0, -1.0, 2.0, 3.0), ("b", "c"), [], wm["hyper_idx"], wm["nsamples"])
    ev.close()
    # finite-fault and geometry wavemaps do not mix in one context
    ctx = Context(0)
    ctx.add_wavemap(2, 8, "nearest_neighbor", None, np.zeros(2, np.int32), np.full(2, 8, np.int32))
    st = gprob["store"]
    sid = ctx.geom_upload_store(st["traces"], st["itmin"], st["nsamples"], st["z0"], st["dz"], st["x0"], st["dx"], st["deltat"])
    with pytest.raises(ValueError, match="one context per mode"):
        ctx.geom_add_wavemap(sid, 40, "multilinear", wm["lats"], wm["lons"], wm["azimuths"], wm["dips"], wm["arrival_times"],
                             wm["taper"], ("b", "c"), [], wm["hyper_idx"], wm["nsamples"])
    ctx.close()


def test_long_window_c2_shape_spot_check():
    """One station x 3 components with config-2 sized windows (2048 samples + flanks): the 9-accumulator
    instantiation of the delay-and-sum kernel and ~100 KB of shared memory per CTA."""
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=1, ns=2048, taper=(-34.0, -24.0, 1000.0, 1010.0), nrec=2300, lead=60.0,
                                    dist_range=(2000e3, 2100e3), dx=20e3, seed=121,
                                    filterer=[dict(kind="stepwise", order=4, lower_corner=0.005, upper_corner=0.2)])
    gprob = _with_data(gprob)
    Q = S.draw_chains(gprob, 6, seed=12)
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    logpts, _ = ev(Q)
    ev.close()
    ref = np.array([O.geometry_synthetics(gprob, p) for p in _points(gprob, Q)])
    _assert_synth_close(got, ref)
    refl = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Q)])
    np.testing.assert_allclose(logpts, refl, rtol=1e-5)


def test_smc_driver_runs_on_the_geometry_engine():
    """The lock-step SMC driver (beat_b200/sampler.py, row f1) only needs ``eval_device``: the geometry-mode engine
    plugs in unchanged and the population moves from the prior to the data-generating source."""
    import torch
    from beat_b200 import sampler as SM
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=4, seed=131)
    q_true = S.draw_chains(gprob, 1, seed=1)[0]
    q_true[gprob["offsets"]["hypers"]] = 0.0
    S.attach_geometry_data(gprob, O.geometry_synthetics(gprob, S.split_point(gprob, q_true)), rel_sigma=0.1)
    ev = _engine(gprob)
    lower = np.concatenate([gprob["priors"][n][0] for n, _ in gprob["var_order"]])
    upper = np.concatenate([gprob["priors"][n][1] for n, _ in gprob["var_order"]])
    n_chains = 256
    prior_like = ev(S.draw_chains(gprob, n_chains, seed=2))[1]
    true_like = ev(q_true[None, :])[1][0]
    out = SM.smc_sample(ev.eval_device, lower, upper, n_chains=n_chains, n_steps=20, device=torch.device("cuda", 0), seed=4, max_stages=80)
    assert out["betas"][-1] == 1.0
    post = out["likelihoods"]
    assert np.median(post) > np.median(prior_like)
    assert np.median(post) > true_like - 0.6 * (true_like - np.median(prior_like))
    fresh_logpts, fresh = ev(out["population"])
    np.testing.assert_allclose(post, fresh, rtol=1e-12)
    for c in (0, 100, 255):
        ref = O.geometry_seismic_eval(gprob, S.split_point(gprob, out["population"][c])).sum()
        assert abs(post[c] - ref) <= 1e-5 * abs(ref) + 1e-3
    assert (out["population"] >= lower).all() and (out["population"] <= upper).all()
    ev.close()


@pytest.mark.parametrize("name", ["stepwise_ml", "bandpass_nn", "bandstop_ad", "station_corr", "two_sources"])
def test_cuda_matches_reference_driven_golden(name):
    """CUDA synthetics against tests/golden/geometry_golden.npz -- produced by the reference's own
    heart.seis_synthetics control flow (make_geometry_golden.py); committed fixture, no oracle in the loop."""
    from test_geometry_cpu import GOLDEN_CASES, load_geometry_golden
    g = load_geometry_golden()
    kw, chop = GOLDEN_CASES[name]
    gprob = S.make_geometry_problem(**kw)
    wm = gprob["wavemaps"][0]
    wm["chop_bounds"] = chop
    wm["ns"] = int(g[name + "_synths"].shape[2])
    ev = _engine(gprob)
    got = ev.get_synthetics(g[name + "_Q"])
    ev.close()
    _assert_synth_close(got, g[name + "_synths"])


def test_station_corrections():
    """time_shift hierarchical (SeisSynthesizer.perform, beat/pytensorf.py:248-252): the engine window, the chop position
    and the taper follow arrival + shift of the chain; data stay where prepare_data put them."""
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=3, station_corrections=True, corr_bounds=(-2.0, 2.0), seed=141)
    q0 = S.draw_chains(gprob, 1, seed=1)[0]
    q0[gprob["offsets"]["time_shifts"]:] = 0.0
    S.attach_geometry_data(gprob, O.geometry_synthetics(gprob, S.split_point(gprob, q0)))
    Q = S.draw_chains(gprob, 24, seed=15)
    ots = gprob["offsets"]["time_shifts"]
    Q[0] = q0
    Q[1, ots:] = [0.5, -1.0, 1.5]                               # on the sampling grid
    Q[2, ots:] = [0.25, -0.25, 0.1]                             # off the grid: floor-snapped chop, first sample on the taper flank
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    logpts, like = ev(Q)
    ev.close()
    ref = np.array([O.geometry_synthetics(gprob, p) for p in _points(gprob, Q)])
    _assert_synth_close(got, ref)
    refl = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Q)])
    _assert_logpts_close(gprob, Q, logpts, refl)
    # the Op mirror takes one time_shift per target, like the reference Op
    from beat_b200.geometry import ArrivalTaper, SeisSynthesizer
    wm = gprob["wavemaps"][0]
    op = SeisSynthesizer(gprob["store"], gprob["event"], dict(lats=wm["lats"], lons=wm["lons"], azimuths=wm["azimuths"], dips=wm["dips"]),
                         ArrivalTaper(*wm["taper"]), wm["arrival_times"], wm["filterer"], station_corrections=True)
    p = S.split_point(gprob, Q[2])
    inputs = {k: v for k, v in p.items() if k not in ("hypers", "time_shifts")}
    inputs["time_shift"] = p["time_shifts"][wm["station_idx"]]
    synths, tmins = op(inputs)
    _assert_synth_close(synths, ref[2])
    np.testing.assert_allclose(tmins, wm["arrival_times"] + inputs["time_shift"] + wm["taper"][1])
    op.close()


def test_two_sources_stack_and_loglike():
    """Two DC sources per chain (pymc vectors of shape (2,)): synthetics of the sources are stacked (heart.py:3719-3724).
    The kernels stack the RAW traces and filter once -- equal to the reference's filter-then-stack because demeaning and
    the IIR filters are linear."""
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=3, n_sources=2, seed=151)
    assert gprob["n_params"] == 19
    q0 = S.draw_chains(gprob, 1, seed=1)[0]
    S.attach_geometry_data(gprob, O.geometry_synthetics(gprob, S.split_point(gprob, q0)))
    Q = S.draw_chains(gprob, 20, seed=16)
    Q[0] = q0
    ev = _engine(gprob)
    got = ev.get_synthetics(Q)
    logpts, _ = ev(Q)
    ev.close()
    ref = np.array([O.geometry_synthetics(gprob, p) for p in _points(gprob, Q)])
    _assert_synth_close(got, ref)
    refl = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Q)])
    _assert_logpts_close(gprob, Q, logpts, refl)
    # one source switched off (magnitude -> tiny moment) leaves the other source's synthetics
    Q1 = Q[:4].copy()
    Q1[:, gprob["offsets"]["magnitude"] + 1] = -20.0
    g1 = S.make_geometry_problem(n_stations=3, n_sources=1, seed=151)
    keep = [gprob["offsets"][v] for v, _ in g1["var_order"] if v != "hypers"] + [gprob["offsets"]["hypers"]]
    ev = _engine(gprob); two = ev.get_synthetics(Q1); ev.close()
    ev = _engine(g1); one = ev.get_synthetics(np.ascontiguousarray(Q1[:, keep])); ev.close()
    _assert_synth_close(two, one)


def test_two_wavemaps_p_and_s_windows():
    """Two wavemaps in one problem (the reference loops over self.wavemaps, beat/models/seismic.py:779-829, and
    concatenates their logpts, :837): a P window and a longer S window with its own filter, hyperparameter and sample
    count share the store and the chain parameters."""
    import copy
    O = _oracle()
    gprob = S.make_geometry_problem(n_stations=2, seed=161)
    wm_p = gprob["wavemaps"][0]
    wm_s = copy.deepcopy(wm_p)
    vp, vs = gprob["store"]["vp"], gprob["store"]["vs"]
    wm_s["arrival_times"] = np.rint(wm_p["arrival_times"] * vp / vs / 0.5) * 0.5          # S arrives later; snapped to the grid
    wm_s["taper"] = (-6.0, -4.0, 26.0, 28.0)
    wm_s["ns"] = 60
    wm_s["nsamples"] = np.full(wm_s["nt"], 60, dtype=np.int32)
    wm_s["filterer"] = [dict(kind="bandpass", order=2, lower_corner=0.03, upper_corner=0.3)]
    wm_s["interpolation"] = "nearest_neighbor"
    gprob["wavemaps"].append(wm_s)
    # a second hyperparameter for the S wavemap
    gprob["n_hypers"] = 2
    gprob["var_order"] = [(v, n) if v != "hypers" else (v, 2) for v, n in gprob["var_order"]]
    gprob["n_params"] += 1
    gprob["priors"]["hypers"] = (np.zeros(2), np.full(2, 4.0))
    wm_s["hyper_idx"] = np.ones(wm_s["nt"], dtype=np.int32)
    q0 = S.draw_chains(gprob, 1, seed=1)[0]
    for iw in range(2):
        S.attach_geometry_data(gprob, O.geometry_synthetics(gprob, S.split_point(gprob, q0), iw), iw=iw, seed=5 + iw)
    Q = S.draw_chains(gprob, 16, seed=17)
    Q[0] = q0
    ev = _engine(gprob)
    assert ev.n_out == 12
    logpts, like = ev(Q)
    syn_s = ev.get_synthetics(Q, wmap_index=1)
    ev.close()
    ref = np.array([O.geometry_seismic_eval(gprob, p) for p in _points(gprob, Q)])
    assert ref.shape == (16, 12)
    for iw, sl in enumerate((slice(0, 6), slice(6, 12))):
        wm = gprob["wavemaps"][iw]
        h = Q[:, gprob["offsets"]["hypers"] + iw][:, None]
        const = wm["slog_pdet"][None, :] + wm["nsamples"][None, :] * (2.0 * h + np.log(2.0 * np.pi))
        np.testing.assert_allclose(-2.0 * logpts[:, sl] - const, -2.0 * ref[:, sl] - const, rtol=1e-5)
    np.testing.assert_allclose(like, ref.sum(axis=1), rtol=1e-5)
    _assert_synth_close(syn_s, np.array([O.geometry_synthetics(gprob, p, 1) for p in _points(gprob, Q)]))


@pytest.mark.parametrize("name", ["op_two_sources", "op_station_corr"])
def test_cuda_matches_reference_op_perform_golden(name):
    """CUDA synthetics against golden vectors from the reference's own SeisSynthesizer.perform (absolute event time,
    unit conversion, source update, station corrections; tests/golden/make_geometry_golden.py)."""
    from test_geometry_cpu import OP_GOLDEN_CASES, load_geometry_golden
    g = load_geometry_golden()
    gprob = S.make_geometry_problem(**OP_GOLDEN_CASES[name])
    ev = _engine(gprob)
    got = ev.get_synthetics(g[name + "_Q"])
    ev.close()
    _assert_synth_close(got, g[name + "_synths"])
