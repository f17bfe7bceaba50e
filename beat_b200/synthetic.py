"""
Synthetic FFI problems of the shapes named in BASELINE.json / SURVEY.md section 8(d): GF libraries, data,
covariances, priors and chain parameter matrices with fixed seeds.  There is no network and the reference
ships no FFI project or GF store, so every benchmark and parity input is generated here.

The returned ``prob`` dict is plain numpy and is consumed both by the GPU engine
(``beat_b200.engine.BatchedFFILogLike.from_problem``) and, in tests, by the CPU oracle.  Nothing in this module
evaluates the forward model under test: "observed" data are a noisy on-grid superposition of library rows.

Library recipe: ``G[t,p,d,s,k] = A[t,p] * w((k - k0[t,p] - s*st_step/dt) / (sigma0 + sigma1*d))`` with a
Ricker wavelet ``w(x) = (1 - 2x^2) exp(-x^2)``: the start-time axis is a genuine shift and the duration axis a
genuine broadening, so multilinear interpolation is meaningful.
"""
from __future__ import annotations

import numpy as np

from .covariance import Covariance, exponential_data_covariance, log_determinant, smoothing_operator_nearest_neighbor

SLIP_PRIORS = {"uparr": (-0.05, 6.0), "uperp": (-0.3, 4.0), "utens": (0.0, 1.0)}   # beat/defaults.py:134-139


def ricker(x, xp=np):
    return (1.0 - 2.0 * x * x) * xp.exp(-(x * x))


def library_block(A, k0, ndur, nst, ns, st_step, dt, sigma0=2.0, sigma1=0.75, xp=np, dtype=None):
    """Library values for the targets/patches of ``A``/``k0`` (both [nt_blk, np]) -> [nt_blk, np, ndur, nst, ns]."""
    if xp is np:
        d = np.arange(ndur, dtype=np.float64)[None, None, :, None, None]
        s = np.arange(nst, dtype=np.float64)[None, None, None, :, None]
        k = np.arange(ns, dtype=np.float64)[None, None, None, None, :]
        x = (k - k0[:, :, None, None, None] - s * (st_step / dt)) / (sigma0 + sigma1 * d)
        out = A[:, :, None, None, None] * ricker(x)
        return out if dtype is None else out.astype(dtype)
    import torch
    dev = A.device
    d = torch.arange(ndur, dtype=torch.float64, device=dev)[None, None, :, None, None]
    s = torch.arange(nst, dtype=torch.float64, device=dev)[None, None, None, :, None]
    k = torch.arange(ns, dtype=torch.float64, device=dev)[None, None, None, None, :]
    x = (k - k0[:, :, None, None, None] - s * (st_step / dt)) / (sigma0 + sigma1 * d)
    out = A[:, :, None, None, None] * ricker(x, torch)
    return out if dtype is None else out.to(dtype)


def _noise_cov(structure, ns, dt, sigma, rng, tzero=2.0):
    if structure == "variance":
        return np.eye(ns) * sigma ** 2
    if structure == "exponential":                       # beat/covariance.py:24-51, scaled by the trace variance
        return exponential_data_covariance(ns, dt, tzero) * sigma ** 2
    if structure == "dense":                             # random SPD, recipe of test/test_covariance.py:72-74
        a = rng.random((ns, ns))
        c = a.T.dot(a) + np.eye(ns) * 0.3
        return c * (sigma ** 2 / np.mean(np.diag(c)))
    raise ValueError("unknown noise structure %s" % structure)


def make_problem(nt=8, subfaults=((4, 6, 2.0),), ns=32, ndur=5, nst=None, dt=0.5, dur_min=0.5, dur_step=0.25,
                 st_min=-5.0, st_step=0.5, slip_vars=("uparr", "uperp"), interpolation="multilinear",
                 noise="exponential", seed=1234, station_corrections=False, hp_specific=False,
                 vel_bounds=(2.2, 4.5), time_bounds=(-5.0, 5.0), corr_bounds=(-1.0, 1.0),
                 geodetic=None, laplacian=False, n_wavemaps=1, build_library=True):
    """Build a synthetic FFI problem.  subfaults: sequence of (n_patch_dip, n_patch_strike, patch_size_km).

    geodetic: None or dict(nobs=[n_1, n_2, ...]) -> that many static datasets with dense non-Toeplitz covariance.
    build_library=False leaves ``wm['G']`` empty (the bench fills the library on the device)."""
    rng = np.random.default_rng(seed)
    subfaults = [tuple(sf) for sf in subfaults]
    nsf = len(subfaults)
    npatch = sum(nd * nstr for nd, nstr, _ in subfaults)
    dur_max = dur_min + (ndur - 1) * dur_step
    if station_corrections:
        st_min = st_min - corr_bounds[1]          # starttimes - correction must stay on the library axis
    if nst is None:
        tmax = max(np.hypot(nd, nstr) * h for nd, nstr, h in subfaults) / vel_bounds[0]
        hi = tmax + time_bounds[1] - (corr_bounds[0] if station_corrections else 0.0)
        nst = int(np.ceil((hi - st_min) / st_step)) + 3
    n_stations = nt if station_corrections else 0

    # ---- variables, order and offsets in q (the reference's order comes from pymc's value_vars)
    n_hypers = (nt * n_wavemaps if hp_specific else n_wavemaps) + (len(geodetic["nobs"]) if geodetic else 0) + (1 if laplacian else 0)
    var_order = [(v, npatch) for v in slip_vars] + [("durations", npatch), ("velocities", npatch),
                                                   ("nucleation_strike", nsf), ("nucleation_dip", nsf), ("time", nsf),
                                                   ("hypers", n_hypers)]
    if station_corrections:
        var_order.append(("time_shifts", n_stations))
    offsets, o = {}, 0
    for name, size in var_order:
        offsets[name] = o
        o += size
    n_params = o

    priors = {}
    for v in slip_vars:
        priors[v] = (np.full(npatch, SLIP_PRIORS[v][0]), np.full(npatch, SLIP_PRIORS[v][1]))
    priors["durations"] = (np.full(npatch, dur_min + 1e-6), np.full(npatch, dur_max))
    priors["velocities"] = (np.full(npatch, vel_bounds[0]), np.full(npatch, vel_bounds[1]))
    # upper bound n*h - 0.01: rint((n*h - h/2)/h) would tie-to-even onto index n
    priors["nucleation_strike"] = (np.zeros(nsf), np.array([nstr * h - 0.01 for _, nstr, h in subfaults]))
    priors["nucleation_dip"] = (np.zeros(nsf), np.array([nd * h - 0.01 for nd, _, h in subfaults]))
    priors["time"] = (np.full(nsf, time_bounds[0]), np.full(nsf, time_bounds[1]))
    priors["hypers"] = (np.zeros(n_hypers), np.full(n_hypers, 4.0))
    if station_corrections:
        priors["time_shifts"] = (np.full(n_stations, corr_bounds[0]), np.full(n_stations, corr_bounds[1]))

    prob = dict(subfaults=subfaults, slip_vars=tuple(slip_vars), npatches=npatch, var_order=var_order, offsets=offsets,
                n_params=n_params, n_hypers=n_hypers, n_time_shifts=n_stations, priors=priors, wavemaps=[], dt=dt,
                seed=seed)

    hyper_cursor = 0
    for iw in range(n_wavemaps):
        A = {v: rng.standard_normal((nt, npatch)) for v in slip_vars}
        k0 = {v: rng.uniform(0.15 * ns, 0.45 * ns, (nt, npatch)) for v in slip_vars}
        wm = dict(nt=nt, ns=ns, ndur=ndur, nst=nst, dur_min=dur_min, dur_step=dur_step, st_min=st_min, st_step=st_step,
                  interpolation=interpolation, A=A, k0=k0, G={})
        if build_library:
            for v in slip_vars:
                wm["G"][v] = library_block(A[v], k0[v], ndur, nst, ns, st_step, dt)
        # "observed" data: on-grid superposition + noise
        u_true = {v: rng.uniform(max(0.0, SLIP_PRIORS[v][0]), SLIP_PRIORS[v][1] * 0.5, npatch) for v in slip_vars}
        di = rng.integers(0, ndur, npatch)
        si = rng.integers(nst // 4, max(nst // 4 + 1, 3 * nst // 4), (nt, npatch))
        clean = np.zeros((nt, ns))
        for v in slip_vars:
            if build_library:
                rows = wm["G"][v][np.arange(nt)[:, None], np.arange(npatch)[None, :], di[None, :], si]   # [nt, np, ns]
            else:
                k = np.arange(ns, dtype=np.float64)[None, None, :]
                x = (k - k0[v][:, :, None] - si[:, :, None] * (st_step / dt)) / (2.0 + 0.75 * di[None, :, None])
                rows = A[v][:, :, None] * ricker(x)
            clean += np.einsum("tpk,p->tk", rows, u_true[v])
        sigma = 0.05 * np.abs(clean).max(axis=1)
        U, lp, data = np.zeros((nt, ns, ns)), np.zeros(nt), np.zeros((nt, ns))
        for t in range(nt):
            Ct = _noise_cov(noise, ns, dt, sigma[t], rng)
            cov = Covariance(data=Ct)
            U[t], lp[t] = cov.chol_inverse, cov.log_pdet
            data[t] = clean[t] + np.linalg.cholesky(Ct).dot(rng.standard_normal(ns))
        wm.update(data=data, U=U, slog_pdet=lp, nsamples=np.full(nt, ns, dtype=np.int32), noise=noise)
        if hp_specific:
            wm["hyper_idx"] = np.arange(hyper_cursor, hyper_cursor + nt, dtype=np.int32)
            hyper_cursor += nt
        else:
            wm["hyper_idx"] = np.full(nt, hyper_cursor, dtype=np.int32)
            hyper_cursor += 1
        wm["station_idx"] = np.arange(nt, dtype=np.int32) if station_corrections else None
        prob["wavemaps"].append(wm)

    if geodetic:
        nobs_list = list(geodetic["nobs"])
        nobs = int(sum(nobs_list))
        G = {v: rng.standard_normal((npatch, nobs)) * 0.01 for v in slip_vars}
        u_true = {v: rng.uniform(0.0, 2.0, npatch) for v in slip_vars}
        clean = sum(G[v].T.dot(u_true[v]) for v in slip_vars)
        slices, Us, lps, data = [], [], [], np.zeros(nobs)
        lo = 0
        for n in nobs_list:
            a = rng.random((n, n))
            Cd = (a.T.dot(a) + np.eye(n) * 0.3) * (0.05 * np.abs(clean).max()) ** 2 / n
            cov = Covariance(data=Cd)
            Us.append(cov.chol_inverse)
            lps.append(cov.log_pdet)
            data[lo:lo + n] = clean[lo:lo + n] + np.linalg.cholesky(Cd).dot(rng.standard_normal(n))
            slices.append((lo, lo + n))
            lo += n
        prob["geodetic"] = dict(G=G, data=data, odw=rng.uniform(0.5, 1.0, nobs), slices=slices, U=Us,
                                slog_pdet=np.array(lps), nsamples=np.array(nobs_list, dtype=np.int32),
                                hyper_idx=np.arange(hyper_cursor, hyper_cursor + len(nobs_list), dtype=np.int32))
        hyper_cursor += len(nobs_list)
    if laplacian:
        if nsf != 1:
            raise ValueError("nearest-neighbour laplacian is only valid for a single flat fault (laplacian.py:214)")
        nd, nstr, h = subfaults[0]
        L = smoothing_operator_nearest_neighbor(nstr, nd, h, h)
        prob["laplacian"] = dict(L=L, sdet=log_determinant(L.T * L), hyper_idx=hyper_cursor)   # laplacian.py:57-60
        hyper_cursor += 1
    assert hyper_cursor == n_hypers
    return prob


def draw_chains(prob, B, seed=4321):
    """B parameter vectors iid uniform from the priors -> Q [B, n_params] float64."""
    rng = np.random.default_rng(seed)
    Q = np.empty((B, prob["n_params"]))
    for name, size in prob["var_order"]:
        lo, hi = prob["priors"][name]
        o = prob["offsets"][name]
        Q[:, o:o + size] = rng.uniform(lo, hi, (B, size))
    return Q


def split_point(prob, q):
    """One row of Q -> dict of named variables (what pymc's bijection hands the reference)."""
    return {name: q[prob["offsets"][name]:prob["offsets"][name] + size] for name, size in prob["var_order"]}


# named configurations (BASELINE.json "configs")
def config_c3(small=False, **kw):
    """FFI seismic headline: 200 patches x 64 stations, fast-sweeping rupture, multilinear, exponential noise."""
    if small:
        args = dict(nt=6, subfaults=((5, 8, 2.0),), ns=40, ndur=5)
    else:
        args = dict(nt=64, subfaults=((10, 20, 2.0),), ns=120, ndur=17, nst=64)
    args.update(kw)
    return make_problem(**args)


# ----------------------------------------------------------------------------------------------------------
# geometry mode (BASELINE config 2): double-couple point source, GF store, stations x components
# ----------------------------------------------------------------------------------------------------------
GEOM_VARS = ("east_shift", "north_shift", "depth", "strike", "dip", "rake", "magnitude", "time", "duration")


def _station_latlon(lat0, lon0, azimuth_deg, distance_m):
    """Destination point on a sphere (placement of synthetic stations only; not part of any parity claim)."""
    d, az = distance_m / 6371000.0, np.deg2rad(azimuth_deg)
    la0, lo0 = np.deg2rad(lat0), np.deg2rad(lon0)
    la = np.arcsin(np.sin(la0) * np.cos(d) + np.cos(la0) * np.sin(d) * np.cos(az))
    lo = lo0 + np.arctan2(np.sin(az) * np.sin(d) * np.cos(la0), np.cos(d) - np.sin(la0) * np.sin(la))
    return np.rad2deg(la), np.rad2deg(lo)


def make_gf_store(nz, nx, z0, dz, x0, dx, deltat, nrec, seed=77, vp=6.0e3, vs=3.5e3, lead=10.0, ragged=True):
    """Synthetic GF store in the layout of a pyrocko type-A store with 10 components: per (source depth, distance)
    node and component a float32 record with its own first-sample index and length.  Records hold a P and an S
    wavelet with move-out plus a small permanent offset after S (so that the repeated end value matters)."""
    rng = np.random.default_rng(seed)
    z = z0 + dz * np.arange(nz)[:, None]
    x = x0 + dx * np.arange(nx)[None, :]
    R = np.sqrt(z ** 2 + x ** 2)
    tp, ts = R / vp, R / vs
    itmin = np.floor((tp - lead) / deltat).astype(np.int32)[:, :, None] + rng.integers(-3, 4, (nz, nx, 10)).astype(np.int32)
    nsamples = np.full((nz, nx, 10), nrec, dtype=np.int32)
    if ragged:
        nsamples -= rng.integers(0, max(1, nrec // 8), (nz, nx, 10)).astype(np.int32)
        short = rng.random((nz, nx, 10)) < 0.1
        nsamples[short] = np.maximum(8, nrec // 3)
    t = (itmin[..., None] + np.arange(nrec)[None, None, None, :]) * deltat
    ap = rng.standard_normal((1, 1, 10, 1)) * 1e-19 * (1.0e5 / R)[:, :, None, None]
    as_ = rng.standard_normal((1, 1, 10, 1)) * 2e-19 * (1.0e5 / R)[:, :, None, None]
    sp, ss = 1.5 + 0.2 * rng.random((1, 1, 10, 1)), 2.5 + 0.3 * rng.random((1, 1, 10, 1))
    xp_, xs_ = (t - tp[:, :, None, None]) / sp, (t - ts[:, :, None, None]) / ss
    traces = ap * ricker(xp_) + as_ * (ricker(xs_) + 0.15 / (1.0 + np.exp(-np.clip(xs_, -40, 40))))
    traces = traces.astype(np.float32)
    traces[np.arange(nrec)[None, None, None, :] >= nsamples[..., None]] = 0.0
    return dict(deltat=deltat, nz=nz, nx=nx, z0=z0, dz=dz, x0=x0, dx=dx, traces=traces, itmin=itmin, nsamples=nsamples,
                vp=vp, vs=vs)


def make_geometry_problem(n_stations=4, channels=("N", "E", "Z"), ns=40, deltat=0.5, taper=(-7.5, -5.0, 15.0, 17.5),
                          interpolation="multilinear", filterer=None, dist_range=(500e3, 700e3), shift_km=10.0,
                          depth_range_km=(2.0, 10.0), dz=2.0e3, dx=4.0e3, nrec=200, time_bounds=(-3.0, 3.0),
                          duration_bounds=(0.0, 6.0), seed=99, hp_specific=False, ragged=True, lead=10.0,
                          station_corrections=False, corr_bounds=(-1.0, 1.0), n_sources=1, stf_type="HalfSinusoid",
                          sample_peak_ratio=True):
    """Synthetic geometry-mode seismic problem (one wavemap, one DC source).  ``data`` / weights are attached later by
    ``attach_geometry_data`` from synthetics of a reference point (tests: the oracle's; bench: the GPU engine's)."""
    rng = np.random.default_rng(seed)
    if filterer is None:
        filterer = [dict(kind="stepwise", order=4, lower_corner=0.01, upper_corner=0.4)]   # heart.Filter defaults, corners for 2 Hz
    a, b, c, d = taper
    if int(np.ceil((c - b) / deltat)) != ns:
        raise ValueError("taper b..c is %g s = %g samples, ns = %d" % (c - b, (c - b) / deltat, ns))
    ev_lat, ev_lon = 37.5, 15.0
    margin = shift_km * 1e3 * 1.5 + 2 * dx
    x0 = np.floor((dist_range[0] - margin) / dx) * dx
    nx = int(np.ceil((dist_range[1] + margin - x0) / dx)) + 1
    z0 = depth_range_km[0] * 1e3 - dz
    nz = int(np.ceil((depth_range_km[1] * 1e3 + dz - z0) / dz)) + 1
    store = make_gf_store(nz, nx, z0, dz, x0, dx, deltat, nrec, seed=seed + 1, ragged=ragged, lead=lead)
    chan = dict(N=(0.0, 0.0), E=(90.0, 0.0), Z=(0.0, -90.0))
    st_az = rng.uniform(0.0, 360.0, n_stations)
    st_dist = rng.uniform(dist_range[0], dist_range[1], n_stations)
    lats, lons, azis, dips, arr, codes = [], [], [], [], [], []
    depth_ref = 0.5 * (depth_range_km[0] + depth_range_km[1]) * 1e3
    for s in range(n_stations):
        la, lo = _station_latlon(ev_lat, ev_lon, st_az[s], st_dist[s])
        # fixed phase arrival of the reference event, snapped to the sampling grid (heart.get_phase_arrival_time, snap=True)
        at = np.rint(np.sqrt(st_dist[s] ** 2 + depth_ref ** 2) / store["vp"] / deltat) * deltat
        for ch in channels:
            lats.append(la); lons.append(lo); azis.append(chan[ch][0]); dips.append(chan[ch][1]); arr.append(at)
            codes.append("ST%02d.%s" % (s, ch))
    nt = len(lats)
    n_hypers = nt if hp_specific else 1
    var_order = [(v, n_sources) for v in GEOM_VARS] + [("hypers", n_hypers)]     # pymc vectors of shape (n_sources,)
    if stf_type == "Triangular" and sample_peak_ratio:      # TriangularSTF.peak_ratio is a sampled variable (defaults.py:238-240)
        var_order.insert(len(GEOM_VARS), ("peak_ratio", n_sources))
    if station_corrections:                   # one hierarchical time shift per station (seismic.py:198-294)
        var_order.append(("time_shifts", n_stations))
    offsets, o = {}, 0
    for name, size in var_order:
        offsets[name] = o
        o += size
    priors = dict(east_shift=(-shift_km, shift_km), north_shift=(-shift_km, shift_km), depth=depth_range_km,
                  strike=(0.0, 360.0), dip=(0.0, 90.0), rake=(-180.0, 180.0), magnitude=(5.5, 6.5), time=time_bounds,
                  duration=duration_bounds, hypers=(np.zeros(n_hypers), np.full(n_hypers, 4.0)))
    if stf_type == "Triangular" and sample_peak_ratio:
        priors["peak_ratio"] = (0.0, 1.0)
    if station_corrections:
        priors["time_shifts"] = (np.full(n_stations, corr_bounds[0]), np.full(n_stations, corr_bounds[1]))
    priors = {k: (np.atleast_1d(np.asarray(v[0], dtype=float)), np.atleast_1d(np.asarray(v[1], dtype=float))) for k, v in priors.items()}
    wm = dict(nt=nt, ns=ns, deltat=deltat, interpolation=interpolation, lats=np.array(lats), lons=np.array(lons),
              azimuths=np.array(azis), dips=np.array(dips), arrival_times=np.array(arr), taper=tuple(taper),
              filterer=filterer, codes=codes, nsamples=np.full(nt, ns, dtype=np.int32),
              hyper_idx=(np.arange(nt, dtype=np.int32) if hp_specific else np.zeros(nt, dtype=np.int32)),
              data=None, U=None, slog_pdet=None,
              station_idx=(np.repeat(np.arange(n_stations, dtype=np.int32), len(channels)) if station_corrections else None))
    return dict(mode="geometry", store=store, event=dict(lat=ev_lat, lon=ev_lon), stf_anchor=-1.0, stf_type=stf_type, var_order=var_order,
                offsets=offsets, n_params=o, n_hypers=n_hypers, n_time_shifts=n_stations if station_corrections else 0,
                n_sources=n_sources,
                priors=priors, wavemaps=[wm], seed=seed)


def attach_geometry_data(gprob, synths, noise="exponential", seed=5, rel_sigma=0.05, iw=0):
    """data = synthetics of a reference point + coloured noise; weights U / log-dets from the noise covariance
    (Covariance.chol_inverse / log_pdet)."""
    rng = np.random.default_rng(seed)
    wm = gprob["wavemaps"][iw]
    nt, ns = wm["nt"], wm["ns"]
    synths = np.asarray(synths, dtype=np.float64).reshape(nt, ns)
    sigma = rel_sigma * max(np.abs(synths).max(), 1e-30)
    Ct = _noise_cov(noise, ns, wm["deltat"], sigma, rng)
    cov = Covariance(data=Ct)
    Lc = np.linalg.cholesky(Ct)
    wm["data"] = synths + (Lc.dot(rng.standard_normal((ns, nt)))).T
    wm["U"] = np.broadcast_to(cov.chol_inverse, (nt, ns, ns))
    wm["slog_pdet"] = np.full(nt, cov.log_pdet)
    wm["noise"] = noise
    return gprob
