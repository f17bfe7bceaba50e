"""
oracle/ffi_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/scipy CPU restatement of the reference's per-chain forward model + log-likelihood
for the finite-fault (FFI) path, one chain at a time, float64 end to end, exactly as the
reference evaluates it inside ``logp_forw_func(q)``.  Every function cites the reference
lines it restates (paths relative to /root/reference).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product (``beat_b200``) never does.

Pinning status
--------------
* fast sweeping: the oracle IS the reference's own compiled C (``oracle/_ref``) where it is
  available, else ``oracle/libfsport.so`` (plain-C restatement, validated against ``_ref``).
* index mapping / ``stack_all`` / ``multivariate_normal_chol`` / ``Covariance`` /
  laplacian prior / geodetic stack: pinned by golden vectors under ``tests/golden/`` that were
  produced by IMPORTING the reference's own Python modules in the dev container
  (``tests/golden/make_golden.py``; pytensor/pyrocko are absent there, so the script installs
  numpy-backed stand-ins for the handful of symbols those modules touch at import/call time --
  the arithmetic that runs is the reference's own source).
* composite level: ``ffi_seismic_eval(return_synth=True)`` reproduces, to 1e-9, synthetics of the reference's own
  ``SeismicDistributerComposite.get_synthetics`` + ``FaultGeometry.point2starttimes`` (numpy fast sweep)
  for one / two subfaults with station corrections (``tests/golden/make_ffi_composite_golden.py``), and, to 1e-10,
  the per-dataset logpts of the reference's own production graph ``SeismicDistributerComposite.get_formula`` run eagerly
  (Sweeper Op -> compiled fast_sweep_ext, pytensor-mode stack_all, multivariate_normal_chol; same script).
"""
from __future__ import annotations

import ctypes
import glob
import importlib.util
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LOG_2PI = np.log(2 * np.pi)  # beat/models/distributions.py:13


# --------------------------------------------------------------------------------------
# fast sweeping
# --------------------------------------------------------------------------------------
_ref_ext = None
_port = None


def load_reference_ext():
    """The reference's own compiled ``fast_sweep_ext`` (oracle/_ref), or None."""
    global _ref_ext
    if _ref_ext is None:
        cands = glob.glob(os.path.join(_HERE, "_ref", "fast_sweep_ext*.so"))
        if not cands:
            return None
        spec = importlib.util.spec_from_file_location("fast_sweep_ext", cands[0])
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        _ref_ext = mod
    return _ref_ext


def load_port():
    """ctypes handle on oracle/libfsport.so (plain-C restatement)."""
    global _port
    if _port is None:
        path = os.path.join(_HERE, "libfsport.so")
        if not os.path.exists(path):
            raise RuntimeError("oracle/libfsport.so missing: run `make -C oracle`")
        lib = ctypes.CDLL(path)
        lib.fsport_sweep.restype = ctypes.c_int
        lib.fsport_sweep.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double] + [ctypes.c_long] * 4
        lib.fsport_sweep_batch.restype = None
        lib.fsport_sweep_batch.argtypes = [
            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_void_p, ctypes.c_void_p,
            ctypes.c_long, ctypes.c_long, ctypes.c_long, ctypes.c_void_p]
        _port = lib
    return _port


def fast_sweep(slowness, patch_size, nuc_dip_idx, nuc_strike_idx, n_patch_dip, n_patch_strike, impl="auto"):
    """Rupture onset times, flat [n_patch_dip * n_patch_strike], index = dip*n_strike + strike.

    Argument order follows ``Sweeper.perform`` (beat/pytensorf.py:475-482), which feeds
    ``(nuc_dip, nuc_strike, n_patch_dip, n_patch_strike)`` into the C routine's
    ``(HypoInStk, HypoInDip, NumInStk, NumInDip)`` (beat/fast_sweeping/fast_sweep_ext.c:120).
    impl: "ref" (compiled reference), "port" (libfsport), "auto" (ref if present else port).
    """
    slowness = np.ascontiguousarray(slowness, dtype=np.float64).ravel()
    if impl in ("auto", "ref"):
        ext = load_reference_ext()
        if ext is not None:
            return ext.fast_sweep(slowness, float(patch_size), int(nuc_dip_idx), int(nuc_strike_idx),
                                  int(n_patch_dip), int(n_patch_strike))
        if impl == "ref":
            raise RuntimeError("oracle/_ref not built")
    lib = load_port()
    out = np.empty_like(slowness)
    lib.fsport_sweep(slowness.ctypes.data, out.ctypes.data, float(patch_size), int(nuc_dip_idx),
                     int(nuc_strike_idx), int(n_patch_dip), int(n_patch_strike))
    return out


def fast_sweep_batch_port(slowness, patch_size, nuc_dip_idx, nuc_strike_idx, n_patch_dip, n_patch_strike):
    """Batched libfsport call: slowness [B, np] -> (T [B, np], iters [B])."""
    lib = load_port()
    slowness = np.ascontiguousarray(slowness, dtype=np.float64)
    B = slowness.shape[0]
    hr = np.ascontiguousarray(nuc_dip_idx, dtype=np.int64)
    hc = np.ascontiguousarray(nuc_strike_idx, dtype=np.int64)
    out = np.empty_like(slowness)
    iters = np.zeros(B, dtype=np.int32)
    lib.fsport_sweep_batch(slowness.ctypes.data, out.ctypes.data, float(patch_size), hr.ctypes.data,
                           hc.ctypes.data, int(n_patch_dip), int(n_patch_strike), B, iters.ctypes.data)
    return out, iters


# --------------------------------------------------------------------------------------
# index mapping
# --------------------------------------------------------------------------------------
def positions2idxs(positions, cell_size, min_pos=0.0, dtype="int16"):
    """beat/utility.py:1542-1558 -- round-half-even of the cell-centre offset, cast to int16."""
    return np.round((np.asarray(positions) - min_pos - (cell_size / 2.0)) / cell_size).astype(dtype)


def fault_locations2idxs(positions_dip, positions_strike, patch_size_dip, patch_size_strike):
    """beat/ffi/fault.py:866-894."""
    return (positions2idxs(positions_dip, patch_size_dip), positions2idxs(positions_strike, patch_size_strike))


def times2idxs(x, x_min, x_step, interpolation):
    """beat/ffi/base.py:486-521 (start times) and :535-568 (durations): identical arithmetic.

    nearest_neighbor -> (int16 idx, None); multilinear -> (int16 ceil idx, factor = ceil - x)."""
    x = np.asarray(x, dtype=np.float64)
    if interpolation == "nearest_neighbor":
        return np.round((x - x_min) / x_step).astype("int16"), None
    elif interpolation == "multilinear":
        d = (x - x_min) / x_step
        c = np.ceil(d).astype("int16")
        return c, c - d
    raise NotImplementedError(interpolation)


# --------------------------------------------------------------------------------------
# GF-library stacking
# --------------------------------------------------------------------------------------
def stack_all(G, durations, starttimes, slips, dur_min, dur_step, st_min, st_step, interpolation="nearest_neighbor"):
    """numpy mode of ``SeismicGFLibrary.stack_all`` (beat/ffi/base.py:607-709).

    G: (ntargets, npatches, ndurations, nstarttimes, nsamples); durations [np]; starttimes [nt, np];
    slips [np]  ->  synthetics [nt, nsamples].
    """
    nt, npatch, _, _, ns = G.shape
    tidx = np.arange(nt)[:, None]            # caller: seismic.py:1298 / :1430
    pidx = np.arange(npatch)                 # ffi/base.py:100-104
    di, rf = times2idxs(durations, dur_min, dur_step, interpolation)
    si, sf = times2idxs(starttimes, st_min, st_step, interpolation)

    if interpolation == "nearest_neighbor":                      # :649-660
        cd = G[tidx, pidx, di, si, :].reshape((nt, npatch, ns)).T
        cslips = np.tile(slips, nt).reshape((nt, npatch))
    else:                                                        # :662-698
        d_cc = G[tidx, pidx, di, si, :].reshape((nt, npatch, ns))
        d_fc = G[tidx, pidx, di, si - 1, :].reshape((nt, npatch, ns))
        d_cf = G[tidx, pidx, di - 1, si, :].reshape((nt, npatch, ns))
        d_ff = G[tidx, pidx, di - 1, si - 1, :].reshape((nt, npatch, ns))
        s_cc = (1 - sf) * (1 - rf) * slips
        s_fc = sf * (1.0 - rf) * slips
        s_cf = (1 - sf) * rf * slips
        s_ff = sf * rf * slips
        cd = np.concatenate([d_cc, d_fc, d_cf, d_ff], axis=1).T
        cslips = np.concatenate([s_cc, s_fc, s_cf, s_ff], axis=1)
    return np.einsum("ijk->ik", cd * cslips.T).T                 # :708-709


def stack_all_loops(G, durations, starttimes, slips, dur_min, dur_step, st_min, st_step, interpolation):
    """Explicit-loop twin of :func:`stack_all` (small cases only) -- independent check of the gather."""
    nt, npatch, _, _, ns = G.shape
    out = np.zeros((nt, ns))
    di, rf = times2idxs(durations, dur_min, dur_step, interpolation)
    si, sf = times2idxs(starttimes, st_min, st_step, interpolation)
    for t in range(nt):
        for p in range(npatch):
            if interpolation == "nearest_neighbor":
                out[t] += slips[p] * G[t, p, di[p], si[t, p]]
            else:
                a, b = sf[t, p], rf[p]
                out[t] += slips[p] * ((1 - a) * (1 - b) * G[t, p, di[p], si[t, p]]
                                      + a * (1 - b) * G[t, p, di[p], si[t, p] - 1]
                                      + (1 - a) * b * G[t, p, di[p] - 1, si[t, p]]
                                      + a * b * G[t, p, di[p] - 1, si[t, p] - 1])
    return out


def geodetic_stack_all(G, slips):
    """``GeodeticGFLibrary.stack_all`` (beat/ffi/base.py:292-305): G [np, nobs] -> G.T @ slips."""
    return G.T.dot(slips)


# --------------------------------------------------------------------------------------
# covariance operands (setup time / per SMC stage)
# --------------------------------------------------------------------------------------
def exponential_data_covariance(n, dt, tzero):
    """beat/covariance.py:24-51."""
    return np.exp(-np.abs(np.arange(n)[:, None] - np.arange(n)[None, :]) * (dt / tzero))


def chol_inverse(C):
    """``Covariance.chol_inverse`` (beat/heart.py:211-237): upper-triangular U with U.T @ U = inv(C)."""
    from scipy import linalg
    try:
        return np.linalg.cholesky(np.linalg.inv(C)).T
    except np.linalg.LinAlgError:
        inverse_chol = np.linalg.inv(linalg.cholesky(C, lower=True).T)
        _, chol_ur = np.linalg.qr(inverse_chol.T)
        return chol_ur


def log_pdet(C):
    """``Covariance.log_pdet`` (beat/heart.py:239-245)."""
    from scipy import linalg
    return np.log(np.diag(linalg.cholesky(C, lower=True))).sum() * 2.0


# --------------------------------------------------------------------------------------
# likelihoods
# --------------------------------------------------------------------------------------
def mvn_chol_logpts(residuals, weights, slog_pdets, nsamples, hps):
    """``multivariate_normal_chol`` (beat/models/distributions.py:72-140), one chain.

    residuals: sequence of n_t vectors; weights[i]: (ns_i, ns_i); slog_pdets[i]; nsamples[i] = M_i;
    hps: scalar (hp_specific False) or [n_t] (hp_specific True)."""
    n_t = len(residuals)
    hps = np.broadcast_to(np.asarray(hps, dtype=np.float64), (n_t,))
    logpts = np.zeros(n_t)
    for i in range(n_t):
        M = np.int16(nsamples[i])
        tmp = weights[i].dot(residuals[i])                          # :128
        norm = M * (2 * hps[i] + LOG_2PI)                           # :129
        logpts[i] = (-0.5) * (slog_pdets[i] + norm + (1 / np.exp(hps[i] * 2)) * tmp.dot(tmp))  # :132-137
    return logpts


def smoothing_operator_nearest_neighbor(n_patch_strike, n_patch_dip, patch_size_strike, patch_size_dip):
    """beat/models/laplacian.py:209-258 (with the neighbour flags of ``_patch_locations`` :172-206)."""
    n = n_patch_dip * n_patch_strike
    L = np.zeros((n, n))
    ddip = 1.0 / patch_size_dip ** 2
    dstr = 1.0 / patch_size_strike ** 2
    for i in range(n):
        r, c = divmod(i, n_patch_strike)
        up, down = r > 0, r < n_patch_dip - 1
        left, right = c > 0, c < n_patch_strike - 1
        L[i, i] = -1 * (up * ddip + down * ddip + left * dstr + right * dstr)
        if up:
            L[i, i - n_patch_strike] = ddip
        if down:
            L[i, i + n_patch_strike] = ddip
        if left:
            L[i, i - 1] = dstr
        if right:
            L[i, i + 1] = dstr
    return L


def laplacian_logpt(L, sdet, u, h_lap):
    """One slip variable's smoothness prior (beat/models/laplacian.py:88-96, :128-136).

    sdet = log_determinant(L.T * L) is a static operand computed by the reference at setup."""
    Ls = L.dot(u)
    exponent = Ls.T.dot(Ls)
    npatch = L.shape[0]
    return (-0.5) * (-sdet + (npatch * (LOG_2PI + 2 * h_lap)) + (1.0 / np.exp(h_lap * 2) * exponent))


# --------------------------------------------------------------------------------------
# the full FFI evaluation for one chain
# --------------------------------------------------------------------------------------
def ffi_seismic_eval(prob, point, impl="auto", return_synth=False):
    """One chain's FFI seismic forward model + per-dataset log-likelihoods.

    Restates ``SeismicDistributerComposite.get_formula`` (beat/models/seismic.py:1253-1349; numpy twin
    ``get_synthetics`` :1392-1459).  ``prob`` is a dict (see beat_b200.synthetic.make_problem):
      subfaults: list of (n_patch_dip, n_patch_strike, patch_size)
      wavemaps: list of dicts {G: {var: array5d}, data [nt, ns], U [nt, ns, ns], slog_pdet [nt],
                nsamples [nt], dur_min, dur_step, st_min, st_step, interpolation,
                station_idx [nt] or None, hyper_idx [nt] (index into point['hypers'])}
      slip_vars: e.g. ("uparr", "uperp")
    ``point``: dict of this chain's variables (uparr, uperp, durations, velocities,
      nucleation_strike [nsf], nucleation_dip [nsf], time [nsf], hypers [n_h], time_shifts [n_st] opt).
    """
    npatch_total = sum(nd * ns_ for nd, ns_, _ in prob["subfaults"])
    starttimes0 = np.zeros(npatch_total)
    off = 0
    for isf, (nd, nstr, h) in enumerate(prob["subfaults"]):
        n = nd * nstr
        dipidx, strikeidx = fault_locations2idxs(point["nucleation_dip"][isf], point["nucleation_strike"][isf], h, h)
        vel = point["velocities"][off:off + n]                                   # fault.py:610-612
        t = fast_sweep(1.0 / vel, h, dipidx, strikeidx, nd, nstr, impl=impl)     # seismic.py:1263-1267
        starttimes0[off:off + n] = t + point["time"][isf]                        # seismic.py:1269-1272
        off += n

    logpts_all, synths = [], []
    for wm in prob["wavemaps"]:
        nt = wm["data"].shape[0]
        starttimes = np.tile(starttimes0, nt).reshape(nt, npatch_total)          # seismic.py:1294-1296
        if wm.get("station_idx") is not None and "time_shifts" in point:
            corr = np.asarray(point["time_shifts"])[wm["station_idx"]]
            starttimes = starttimes - np.repeat(corr, npatch_total).reshape(nt, npatch_total)  # :1283-1291
        synth = np.zeros_like(wm["data"], dtype=np.float64)
        for var in prob["slip_vars"]:                                            # seismic.py:1317-1330
            synth += stack_all(wm["G"][var], point["durations"], starttimes, point[var],
                               wm["dur_min"], wm["dur_step"], wm["st_min"], wm["st_step"], wm["interpolation"])
        residuals = wm["data"] - synth                                           # seismic.py:1332
        hps = np.asarray(point["hypers"])[wm["hyper_idx"]]
        logpts_all.append(mvn_chol_logpts(residuals, wm["U"], wm["slog_pdet"], wm["nsamples"], hps))
        synths.append(synth)
    logpts = np.concatenate(logpts_all)                                          # seismic.py:1348
    if return_synth:
        return logpts, synths, starttimes0
    return logpts


def ffi_geodetic_eval(geo, point):
    """``GeodeticDistributerComposite.get_formula`` (beat/models/geodetic.py:1065-1084), one chain.

    geo: {G: {var: [np, nobs]}, data [nobs], odw [nobs], slices: [(lo, hi)...] per dataset,
          U: [list of (n_i, n_i)], slog_pdet, nsamples, hyper_idx}"""
    mu = np.zeros_like(geo["data"], dtype=np.float64)
    for var, G in geo["G"].items():
        mu += geodetic_stack_all(G, point[var])
    r = (geo["data"] - mu) * geo["odw"]
    residuals = [r[lo:hi] for lo, hi in geo["slices"]]
    hps = np.asarray(point["hypers"])[geo["hyper_idx"]]
    return mvn_chol_logpts(residuals, geo["U"], geo["slog_pdet"], geo["nsamples"], hps)


def ffi_laplacian_eval(lap, point, slip_vars):
    """``LaplacianDistributerComposite.get_formula`` (beat/models/laplacian.py:98-139): summed over slip vars."""
    h = np.asarray(point["hypers"])[lap["hyper_idx"]]
    return sum(laplacian_logpt(lap["L"], lap["sdet"], point[v], h) for v in slip_vars)
