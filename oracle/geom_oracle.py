"""
oracle/geom_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

numpy/scipy CPU restatement of the reference's GEOMETRY-mode seismic forward model + log-likelihood
(BASELINE config 2: double-couple point source, N stations x 3 components),
one chain at a time, exactly in the order the reference evaluates it:

    SeisSynthesizer.perform            beat/pytensorf.py:241-302
      -> heart.seis_synthetics         beat/heart.py:3564-3762   (taperers :3636-3649, pre_stack_cut :3651-3653,
                                                                  engine.process :3657, post_process :3679-3690,
                                                                  stack over sources :3719-3724)
         -> post_process_trace         beat/heart.py:3466-3525   (filter, extend, taper, chop)
            Filter.apply               beat/heart.py:366-392
            ArrivalTaper               beat/heart.py:266-336
            DynamicTarget.update_target_times  beat/heart.py:457-477
    residuals / multivariate_normal_chol      beat/models/seismic.py:819-828, beat/models/distributions.py:72-140

PINNED (BEAT side): tests/golden/geometry_golden.npz holds synthetics produced by the reference's OWN
heart.seis_synthetics / get_phase_taperer / update_target_times / post_process_trace / Filter.apply, imported from
/root/reference (tests/golden/make_geometry_golden.py; pyrocko's Trace and engine replaced by stand-ins): the window the
reference puts on each target, the order and arguments of highpass / lowpass / extend / taper / chop, the stacking and
the returned tmins are the reference's, and this module reproduces them bit for bit.  Two more cases run the
reference Op itself (pytensorf.SeisSynthesizer.perform: adjust_point_units, split_point, update_source, event-time
offset, station corrections) with an absolute event time and agree to float32 rounding.

PARITY UNPINNED for the part below the engine.process() call: the arithmetic of the seismogram synthesis, the
filters and the taper lives in the third-party package **pyrocko** (>= 2023.10.11, reference pyproject.toml:35),
which is neither vendored under /root/reference nor installed in this image, and the reference ships no golden
vectors for this path (its tests need private GF stores, test/test_composites.py:69-91).  The functions marked
[pyrocko] restate pyrocko's published algorithm (module and function named in each docstring) from its public
documentation/source as the builder knows it; they cannot be checked against a pyrocko binary here.  What IS pinned:
the two independent formulations of the double-couple moment tensor agree (Euler-rotation form used by
pyrocko.moment_tensor vs Aki & Richards' closed form), the IIR filters are scipy.signal.butter/lfilter (the very
functions pyrocko calls), the likelihood is the golden-vector-pinned mvn_chol_logpts of ffi_oracle, and the
BEAT-side control flow follows the cited reference lines.

Time convention: all times are relative to the reference event's origin time (the reference adds
``events[i].time`` in epoch seconds, beat/pytensorf.py:266, and pyrocko works in epoch seconds throughout; the
window/sample indices are identical as long as the event time is a multiple of the store's sampling interval).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
"""
from __future__ import annotations

import math

import numpy as np
from scipy import signal

from .ffi_oracle import mvn_chol_logpts

d2r = math.pi / 180.0
r2d = 180.0 / math.pi
EARTHRADIUS = 6371.0 * 1000.0                 # [pyrocko] orthodrome.earthradius
EARTHRADIUS_EQUATOR = 6378.14 * 1000.0        # [pyrocko] orthodrome.earthradius_equator
EARTH_OBLATENESS = 1.0 / 298.257223563        # [pyrocko] orthodrome.earth_oblateness
KM = 1000.0                                   # beat/utility.py:72

# GF component scheme 'elastic10' ([pyrocko] gf.meta): records per (source_depth, distance) node
NCOMP = 10
G_NE = (0, 1, 2, 8, 3, 4)                     # components entering the north / east seismogram
G_D = (5, 6, 7, 9)                            # components entering the down seismogram


# --------------------------------------------------------------------------------------
# source: moment tensor, source time function
# --------------------------------------------------------------------------------------
def magnitude_to_moment(magnitude):
    """[pyrocko] moment_tensor.magnitude_to_moment: M0 [Nm] = 10**(1.5*(Mw+10.7)) * 1e-7."""
    return 10.0 ** (1.5 * (magnitude + 10.7)) * 1.0e-7


def euler_to_matrix(alpha, beta, gamma):
    """[pyrocko] moment_tensor.euler_to_matrix (z-x-z Euler angles -> rotation matrix)."""
    ca, cb, cg = math.cos(alpha), math.cos(beta), math.cos(gamma)
    sa, sb, sg = math.sin(alpha), math.sin(beta), math.sin(gamma)
    return np.array([[cb * cg - ca * sb * sg, sb * cg + ca * cb * sg, sa * sg],
                     [-cb * sg - ca * sb * cg, -sb * sg + ca * cb * cg, sa * cg],
                     [sa * sb, -sa * cb, ca]])


def dc_m6(strike, dip, rake, moment=1.0):
    """[pyrocko] moment_tensor.MomentTensor(strike, dip, rake).m6() * moment in north-east-down:
    m = R^T m_unrot R with R = euler_to_matrix(dip, strike, -rake), m_unrot = [[0,0,-1],[0,0,0],[-1,0,0]];
    m6 = (mnn, mee, mdd, mne, mnd, med).  This is what DCSource.discretize_basesource builds."""
    R = euler_to_matrix(dip * d2r, strike * d2r, -rake * d2r)
    m_unrot = np.array([[0.0, 0.0, -1.0], [0.0, 0.0, 0.0], [-1.0, 0.0, 0.0]])
    m = R.T.dot(m_unrot).dot(R) * moment
    return np.array([m[0, 0], m[1, 1], m[2, 2], m[0, 1], m[0, 2], m[1, 2]])


def dc_m6_aki_richards(strike, dip, rake, moment=1.0):
    """Aki & Richards (1980) eq. 4.91 closed form (x=north, y=east, z=down) -- independent check of dc_m6, and
    the form the CUDA kernel evaluates."""
    phi, delta, lam = strike * d2r, dip * d2r, rake * d2r
    sd, cd, s2d, c2d = math.sin(delta), math.cos(delta), math.sin(2 * delta), math.cos(2 * delta)
    sl, cl = math.sin(lam), math.cos(lam)
    sp, cp, s2p, c2p = math.sin(phi), math.cos(phi), math.sin(2 * phi), math.cos(2 * phi)
    mnn = -(sd * cl * s2p + s2d * sl * sp * sp)
    mee = sd * cl * s2p - s2d * sl * cp * cp
    mdd = s2d * sl
    mne = sd * cl * c2p + 0.5 * s2d * sl * s2p
    mnd = -(cd * cl * cp + c2d * sl * sp)
    med = -(cd * cl * sp - c2d * sl * cp)
    return np.array([mnn, mee, mdd, mne, mnd, med]) * moment


def py_round(x):
    """Python 3 round(): half to even (what pyrocko's STF discretisation uses on floats)."""
    return float(np.rint(x))


def halfsinusoid_discretize_t(duration, anchor, deltat, tref):
    """[pyrocko] gf.seismosizer.HalfSinusoidSTF.discretize_t (exponent 1): the STF is integrated over sampling
    intervals centred on the store's time grid; returns (times on the grid, amplitudes summing to 1).
    BEAT initialises its sources with ``HalfSinusoidSTF(anchor=-1)`` (beat/config.py:2060) and samples ``duration``."""
    tmin_stf = tref - duration * (anchor + 1.0) * 0.5
    tmax_stf = tref + duration * (1.0 - anchor) * 0.5
    tmin = py_round(tmin_stf / deltat) * deltat
    tmax = py_round(tmax_stf / deltat) * deltat
    nt = int(py_round((tmax - tmin) / deltat)) + 1
    if nt > 1:
        t_edges = np.maximum(tmin_stf, np.minimum(tmax_stf, np.linspace(tmin - 0.5 * deltat, tmax + 0.5 * deltat, nt + 1)))
        fint = -np.cos((t_edges - tmin_stf) * (math.pi / duration))
        amplitudes = fint[1:] - fint[:-1]
        amplitudes /= np.sum(amplitudes)
    else:
        amplitudes = np.ones(1)
    times = np.linspace(tmin, tmax, nt)
    return times, amplitudes


def plf_integrate_piecewise(x_edges, x, y):
    """[pyrocko] util.plf_integrate_piecewise: integrals over the bins [x_edges[k], x_edges[k+1]] of the piecewise-linear
    function through (x, y), continued by its end values.  Restated as the exact integral: bin by bin, the trapezoids
    between the breakpoints falling inside the bin (a jump -- two breakpoints with the same x -- contributes nothing)."""
    x_edges, x, y = (np.asarray(v, dtype=np.float64) for v in (x_edges, x, y))
    out = np.zeros(x_edges.size - 1)
    for k in range(out.size):
        lo, hi = x_edges[k], x_edges[k + 1]
        inner = [(xi, i) for i, xi in enumerate(x) if lo < xi < hi]
        px = [lo] + [xi for xi, _ in inner] + [hi]
        # value at the bin edges: np.interp semantics (right-continuous at a jump is immaterial for the integral
        # unless the edge sits exactly on it, where the one-sided limits towards the bin interior are what counts)
        def left_limit(t):
            i = np.searchsorted(x, t, side="left")            # first breakpoint >= t
            if i == 0:
                return y[0]
            if i == x.size:
                return y[-1]
            return y[i - 1] + (y[i] - y[i - 1]) * (t - x[i - 1]) / (x[i] - x[i - 1]) if x[i] > x[i - 1] else y[i - 1]

        def right_limit(t):
            i = np.searchsorted(x, t, side="right")           # first breakpoint > t
            if i == 0:
                return y[0]
            if i == x.size:
                return y[-1]
            return y[i - 1] + (y[i] - y[i - 1]) * (t - x[i - 1]) / (x[i] - x[i - 1]) if x[i] > x[i - 1] else y[i]
        tot = 0.0
        for j in range(len(px) - 1):
            tot += 0.5 * (right_limit(px[j]) + left_limit(px[j + 1])) * (px[j + 1] - px[j])
        out[k] = tot
    return out


def sshift(times, amplitudes, tshift, deltat):
    """[pyrocko] gf.seismosizer.sshift: move a discretised STF by a sub-sample amount (linear split between the two
    neighbouring grid points: one point more); a shift that is a whole number of samples leaves the arrays untouched."""
    t0 = math.floor(tshift / deltat) * deltat
    t1 = math.ceil(tshift / deltat) * deltat
    if t0 == t1:
        return times, amplitudes
    amplitudes2 = np.zeros(amplitudes.size + 1)
    amplitudes2[:-1] += (t1 - tshift) / deltat * amplitudes
    amplitudes2[1:] += (tshift - t0) / deltat * amplitudes
    times2 = np.arange(times.size + 1, dtype=np.float64) * deltat + times[0] + t0
    return times2, amplitudes2


def boxcar_discretize_t(duration, anchor, deltat, tref):
    """[pyrocko] gf.seismosizer.BoxcarSTF.discretize_t: bin integrals of the boxcar on the store's time grid, then a
    sub-sample shift so that the discrete centroid equals ``centroid_time = tref - duration * anchor / 2``."""
    tmin_stf = tref - duration * (anchor + 1.0) * 0.5
    tmax_stf = tref + duration * (1.0 - anchor) * 0.5
    tmin = py_round(tmin_stf / deltat) * deltat
    tmax = py_round(tmax_stf / deltat) * deltat
    nt = int(py_round((tmax - tmin) / deltat)) + 1
    times = np.linspace(tmin, tmax, nt)
    amplitudes = np.ones_like(times)
    if times.size > 1:
        t_edges = np.linspace(tmin - 0.5 * deltat, tmax + 0.5 * deltat, nt + 1)
        t = tmin_stf + duration * np.array([0.0, 0.0, 1.0, 1.0])
        f = np.array([0.0, 1.0, 1.0, 0.0])
        amplitudes = plf_integrate_piecewise(t_edges, t, f)
        amplitudes /= np.sum(amplitudes)
    tshift = np.sum(amplitudes * times) - (tref - 0.5 * duration * anchor)
    return sshift(times, amplitudes, -tshift, deltat)


def triangular_centroid_ratio(peak_ratio):
    """[pyrocko] TriangularSTF.centroid_ratio."""
    ra = peak_ratio
    rb = 1.0 - ra
    return ra + (rb ** 2 / 3.0 - ra ** 2 / 3.0) / (ra + rb)


def triangular_discretize_t(duration, peak_ratio, anchor, deltat, tref):
    """[pyrocko] gf.seismosizer.TriangularSTF.discretize_t (+ tminmax_stf): a triangle of base ``duration`` whose apex sits
    at ``peak_ratio`` of the base, anchored through its centroid; bin integrals on the store's time grid."""
    ca = triangular_centroid_ratio(peak_ratio)
    cb = 1.0 - ca
    if anchor <= 0.0:
        tmin_stf = tref - ca * duration * (anchor + 1.0)
        tmax_stf = tmin_stf + duration
    else:
        tmax_stf = tref + cb * duration * (1.0 - anchor)
        tmin_stf = tmax_stf - duration
    tmin = py_round(tmin_stf / deltat) * deltat
    tmax = py_round(tmax_stf / deltat) * deltat
    nt = int(py_round((tmax - tmin) / deltat)) + 1
    if nt > 1:
        t_edges = np.linspace(tmin - 0.5 * deltat, tmax + 0.5 * deltat, nt + 1)
        t = tmin_stf + duration * np.array([0.0, peak_ratio, 1.0])
        f = np.array([0.0, 1.0, 0.0])
        amplitudes = plf_integrate_piecewise(t_edges, t, f)
        amplitudes /= np.sum(amplitudes)
    else:
        amplitudes = np.ones(1)
    times = np.linspace(tmin, tmax, nt)
    return times, amplitudes


def stf_discretize_t(gprob, src, deltat):
    """The source time function BEAT attaches to a geometry-mode source: ``stf_catalog[stf_type](anchor=-1)``
    (beat/config.py:2058-2060; catalogue beat/sources.py:723-729 = pyrocko's Boxcar / Triangular / HalfSinusoid); the
    sampled ``duration`` (and ``peak_ratio`` for the triangle) reach it through utility.update_source (:773-797)."""
    kind = gprob.get("stf_type", "HalfSinusoid")
    anchor = gprob["stf_anchor"]
    if kind == "HalfSinusoid":
        return halfsinusoid_discretize_t(src["duration"], anchor, deltat, src["time"])
    if kind == "Boxcar":
        return boxcar_discretize_t(src["duration"], anchor, deltat, src["time"])
    if kind == "Triangular":
        return triangular_discretize_t(src["duration"], src.get("peak_ratio", gprob.get("peak_ratio", 0.5)), anchor, deltat, src["time"])
    raise ValueError("unknown stf_type %r" % (kind,))


# --------------------------------------------------------------------------------------
# geometry: source -> receiver distance, azimuth, back-azimuth
# --------------------------------------------------------------------------------------
def ne_to_latlon(lat0, lon0, north_m, east_m):
    """[pyrocko] orthodrome.ne_to_latlon -> azidist_to_latlon_rad (spherical earth, arcsin form)."""
    a = math.sqrt(north_m ** 2 + east_m ** 2) / EARTHRADIUS
    gamma = math.atan2(east_m, north_m)
    b = math.pi / 2.0 - lat0 * d2r
    alphasign = -1.0 if gamma < 0.0 else 1.0
    gamma = abs(gamma)
    c = math.acos(min(1.0, max(-1.0, math.cos(a) * math.cos(b) + math.sin(a) * math.sin(b) * math.cos(gamma))))
    sc = math.sin(c)
    alpha = math.asin(min(1.0, max(-1.0, (math.sin(a) * math.sin(gamma) / sc) if sc != 0.0 else 0.0)))
    if math.cos(a) - math.cos(b) * math.cos(c) < 0.0:
        alpha = (math.pi - alpha) if alpha > 0.0 else (-math.pi - alpha)
    lat = r2d * (math.pi / 2.0 - c)
    lon = lon0 + r2d * alpha * alphasign
    lon = (lon + 180.0) % 360.0 - 180.0        # orthodrome.wrap(lon, -180, 180)
    return lat, lon


def cosdelta(alat, alon, blat, blon):
    """[pyrocko] orthodrome.cosdelta."""
    return min(1.0, math.sin(alat * d2r) * math.sin(blat * d2r)
               + math.cos(alat * d2r) * math.cos(blat * d2r) * math.cos(d2r * (blon - alon)))


def azimuth(alat, alon, blat, blon):
    """[pyrocko] orthodrome.azimuth (degrees, from a towards b)."""
    return r2d * math.atan2(math.cos(alat * d2r) * math.cos(blat * d2r) * math.sin(d2r * (blon - alon)),
                            math.sin(d2r * blat) - math.sin(d2r * alat) * cosdelta(alat, alon, blat, blon))


def distance_accurate50m(alat, alon, blat, blon):
    """[pyrocko] orthodrome.distance_accurate50m (Meeus' ellipsoidal formula), metres."""
    f = (alat + blat) * d2r / 2.0
    g = (alat - blat) * d2r / 2.0
    h = (alon - blon) * d2r / 2.0
    s = math.sin(g) ** 2 * math.cos(h) ** 2 + math.cos(f) ** 2 * math.sin(h) ** 2
    c = math.cos(g) ** 2 * math.cos(h) ** 2 + math.sin(f) ** 2 * math.sin(h) ** 2
    w = math.atan(math.sqrt(s / c))
    if w == 0.0:
        return 0.0
    r = math.sqrt(s * c) / w
    d = 2.0 * w * EARTHRADIUS_EQUATOR
    h1 = (3.0 * r - 1.0) / (2.0 * c)
    h2 = (3.0 * r + 1.0) / (2.0 * s)
    return d * (1.0 + EARTH_OBLATENESS * h1 * math.sin(f) ** 2 * math.cos(g) ** 2
                - EARTH_OBLATENESS * h2 * math.cos(f) ** 2 * math.sin(g) ** 2)


def source_receiver_geometry(ev_lat, ev_lon, north_m, east_m, rlat, rlon):
    """[pyrocko] gf.meta.DiscretizedSource.distances_to / azibazis_to for one point source whose position is
    (north_shift, east_shift) about the event origin, and a receiver at (rlat, rlon) without shifts."""
    if ev_lat == rlat and ev_lon == rlon:                        # same_origin: cartesian branch
        dist = math.sqrt(north_m ** 2 + east_m ** 2)
        azi = r2d * math.atan2(0.0 - east_m, 0.0 - north_m)
        return dist, azi, azi + 180.0
    slat, slon = ne_to_latlon(ev_lat, ev_lon, north_m, east_m)
    dist = distance_accurate50m(slat, slon, rlat, rlon)
    return dist, azimuth(slat, slon, rlat, rlon), azimuth(rlat, rlon, slat, slon)


# --------------------------------------------------------------------------------------
# GF store (ConfigTypeA: source depth x distance grid, one receiver depth)
# --------------------------------------------------------------------------------------
class OutOfBounds(Exception):
    """[pyrocko] gf.meta.OutOfBounds -- the reference turns the resulting failure into a ValueError
    (beat/heart.py:3659-3662); the CUDA path reports BEATGPU_E_INDEX and NaN logpts."""


def store_nodes(store, depth, distance, interpolation):
    """[pyrocko] gf.meta.ConfigTypeA index functions: nearest_neighbor -> round() of the fractional index;
    multilinear -> the 2x2 vicinity with weights (1-frac) for the floor node and frac for the ceil node
    (a coordinate that sits on a node contributes one node only).  Returns [(iz, ix, weight), ...]."""
    xa = (depth - store["z0"]) / store["dz"]
    xb = (distance - store["x0"]) / store["dx"]
    nz, nx = store["nz"], store["nx"]
    eps = 1e-9
    if not (-eps <= xa <= nz - 1 + eps and -eps <= xb <= nx - 1 + eps):
        raise OutOfBounds("depth %g m, distance %g m" % (depth, distance))
    xa = min(max(xa, 0.0), nz - 1.0)
    xb = min(max(xb, 0.0), nx - 1.0)
    if interpolation == "nearest_neighbor":
        return [(int(np.rint(xa)), int(np.rint(xb)), 1.0)]
    out = []
    for ia, wa in ((math.floor(xa), 1.0 - (xa - math.floor(xa))), (math.ceil(xa), xa - math.floor(xa))):
        if wa == 0.0 or (ia == math.floor(xa) and False):
            continue
        for ib, wb in ((math.floor(xb), 1.0 - (xb - math.floor(xb))), (math.ceil(xb), xb - math.floor(xb))):
            if wb == 0.0:
                continue
            out.append((int(ia), int(ib), wa * wb))
    return out


def component_weights(m6, azi, bazi, tazi, tdip):
    """[pyrocko] gf.meta.DiscretizedMTSource.make_weights (scheme 'elastic10') folded with the sensor orientation
    of gf.Target (north = ca*cd, east = sa*cd, down = sd of the target's azimuth/dip): weight of each of the 10
    GF components in the seismogram of one target (factors below 1e-15 in magnitude are dropped, as pyrocko's
    ``nonzero`` does, so a vertical sensor reads the down seismogram only)."""
    sa, ca = math.sin(azi * d2r), math.cos(azi * d2r)
    sa2, ca2 = math.sin(2.0 * azi * d2r), math.cos(2.0 * azi * d2r)
    sb, cb = math.sin(bazi * d2r - math.pi), math.cos(bazi * d2r - math.pi)
    f0 = m6[0] * ca ** 2 + m6[1] * sa ** 2 + m6[3] * sa2
    f1 = m6[4] * ca + m6[5] * sa
    f2 = m6[2]
    f3 = 0.5 * (m6[1] - m6[0]) * sa2 + m6[3] * ca2
    f4 = m6[5] * ca - m6[4] * sa
    f5 = m6[0] * sa ** 2 + m6[1] * ca ** 2 - m6[3] * sa2
    w_n = (cb * f0, cb * f1, cb * f2, cb * f5, -sb * f3, -sb * f4)
    w_e = (sb * f0, sb * f1, sb * f2, sb * f5, cb * f3, cb * f4)
    w_d = (f0, f1, f2, f5)
    fn = math.cos(tazi * d2r) * math.cos(tdip * d2r)
    fe = math.sin(tazi * d2r) * math.cos(tdip * d2r)
    fd = math.sin(tdip * d2r)
    # [pyrocko] seismosizer.VectorRule.apply_: a base seismogram enters only if nonzero(factor, eps=1e-15)
    fn, fe, fd = (f if abs(f) > 1e-15 else 0.0 for f in (fn, fe, fd))
    W = np.zeros(NCOMP)
    for g, wn, we in zip(G_NE, w_n, w_e):
        W[g] = fn * wn + fe * we
    for g, wd in zip(G_D, w_d):
        W[g] = fd * wd
    return W


def store_sum(store, elements, itmin_out, nsamples):
    """[pyrocko] gf.store.Store.sum (reference implementation): out[i] += w * trace[(itmin_out + i) - idelay - itmin_rec]
    for every element (iz, ix, g, delay, weight); a delay that is not a multiple of deltat is split linearly
    between floor and ceil; samples before / after a stored trace repeat its first / last value.
    Accumulates in float32 like the GF store's dtype."""
    deltat = store["deltat"]
    out = np.zeros(nsamples, dtype=np.float32)
    i_abs = itmin_out + np.arange(nsamples)
    for iz, ix, g, delay, weight in elements:
        n_rec = int(store["nsamples"][iz, ix, g])
        it_rec = int(store["itmin"][iz, ix, g])
        data = store["traces"][iz, ix, g, :n_rec]
        x = delay / deltat
        fl, ce = math.floor(x), math.ceil(x)
        parts = [(fl, 1.0)] if fl == ce else [(fl, ce - x), (ce, x - fl)]
        for idelay, frac in parts:
            j = np.clip(i_abs - int(idelay) - it_rec, 0, n_rec - 1)
            out += np.float32(weight * frac) * data[j]
    return out


def target_window(wm, t):
    """Window the engine computes for target t (pre_stack_cut, beat/heart.py:3651-3653 -> update_target_times
    :457-477: taper a/d widened by twice the fade-in), as [pyrocko] seismosizer does: itmin = floor(tmin/deltat),
    nsamples = ceil(tmax/deltat) - itmin + 1."""
    a, b, c, d = (wm["arrival_times"][t] + x for x in wm["taper"])
    tol = 2.0 * (b - a)
    tmin, tmax = a - tol, d + tol
    itmin = int(math.floor(tmin / wm["deltat"]))
    itmax = int(math.ceil(tmax / wm["deltat"]))
    return itmin, itmax - itmin + 1


def seismogram(gprob, wm, t, src):
    """[pyrocko] LocalEngine.process for one (DC point source, target): discretised source (STF points on the
    time grid, each carrying m6 * amplitude) x interpolation nodes x GF components -> delay-and-sum over the
    target's window.  Returns (float32 trace, itmin)."""
    store = gprob["store"]
    dist, azi, bazi = source_receiver_geometry(gprob["event"]["lat"], gprob["event"]["lon"], src["north_shift"],
                                               src["east_shift"], wm["lats"][t], wm["lons"][t])
    m6 = dc_m6(src["strike"], src["dip"], src["rake"], magnitude_to_moment(src["magnitude"]))
    times, amps = stf_discretize_t(gprob, src, store["deltat"])
    W = component_weights(m6, azi, bazi, wm["azimuths"][t], wm["dips"][t])
    nodes = store_nodes(store, src["depth"], dist, wm["interpolation"])
    elements = []
    for tk, ak in zip(times, amps):
        for iz, ix, wn in nodes:
            for g in range(NCOMP):
                if W[g] != 0.0:
                    elements.append((iz, ix, g, tk, ak * wn * W[g]))
    itmin, n = target_window(wm, t)
    return store_sum(store, elements, itmin, n), itmin


# --------------------------------------------------------------------------------------
# post-processing (beat/heart.py:3466-3525)
# --------------------------------------------------------------------------------------
def filter_sections(filterer, deltat):
    """Filter.apply (beat/heart.py:377-392) -> [pyrocko] Trace.highpass/lowpass/bandpass: scipy.signal.butter(order,
    corner*2*deltat, btype) + lfilter, the high-pass (and the single band-pass) after removing the mean.
    BandstopFilter.apply (:406-412) -> bandstop without demeaning.  Returns [(b, a, demean), ...]."""
    out = []
    for f in filterer:
        if f["kind"] == "stepwise":
            out.append(signal.butter(f["order"], [f["lower_corner"] * 2.0 * deltat], btype="high") + (True,))
            out.append(signal.butter(f["order"], [f["upper_corner"] * 2.0 * deltat], btype="low") + (False,))
        elif f["kind"] == "bandpass":
            out.append(signal.butter(f["order"], [c * 2.0 * deltat for c in (f["lower_corner"], f["upper_corner"])],
                                     btype="band") + (True,))
        elif f["kind"] == "bandstop":
            out.append(signal.butter(f["order"], [c * 2.0 * deltat for c in (f["lower_corner"], f["upper_corner"])],
                                     btype="bandstop") + (False,))
        else:
            raise ValueError(f["kind"])
    return out


def cos_taper_inplace(y, x0, dx, a, b, c, d):
    """[pyrocko] trace.CosTaper.__call__ -> apply_costaper: zero before a / after d, raised-cosine flanks, indices
    snapped with ceil."""
    n = y.size

    def hi(x):
        return max(0, min(int(math.ceil((x - x0) / dx)), n))

    y[:hi(a)] = 0.0
    y[hi(a):hi(b)] *= 0.5 - 0.5 * np.cos((dx * np.arange(hi(a), hi(b)) - (a - x0)) / (b - a) * math.pi)
    y[hi(c):hi(d)] *= 0.5 + 0.5 * np.cos((dx * np.arange(hi(c), hi(d)) - (c - x0)) / (d - c) * math.pi)
    y[hi(d):] = 0.0


def chop_indices(tmin_trace, deltat, n, tlo, thi):
    """[pyrocko] Trace.chop(tlo, thi, snap=(floor, floor)) as called at beat/heart.py:3522."""
    ibeg = max(0, int(math.floor((tlo - tmin_trace) / deltat)))
    iend = min(n, int(math.floor((thi - tmin_trace) / deltat)))
    return ibeg, iend


def post_process(wm, t, raw, itmin, chop_bounds=("b", "c")):
    """post_process_trace (beat/heart.py:3466-3525) for a synthetic: no transfer function (target.response None,
    :3496), filters (:3508-3511), no down-sampling, extend/taper/chop (:3516-3523, tolerance factor 0)."""
    deltat = wm["deltat"]
    y = raw.astype(np.float64)
    for b, a, demean in filter_sections(wm["filterer"], deltat):
        if demean:
            y = y - np.mean(y)
        y = signal.lfilter(b, a, y)
    ta, tb, tc, td = (wm["arrival_times"][t] + x for x in wm["taper"])
    bounds = dict(a=ta, b=tb, c=tc, d=td)
    lower, upper = bounds[chop_bounds[0]], bounds[chop_bounds[1]]
    tmin_trace = itmin * deltat
    # trace.extend(lower, upper, fillmethod="zeros") is a no-op: the engine window contains [a, d]
    cos_taper_inplace(y, tmin_trace, deltat, ta, tb, tc, td)
    ibeg, iend = chop_indices(tmin_trace, deltat, y.size, lower, upper)
    return y[ibeg:iend]


# --------------------------------------------------------------------------------------
# the composite evaluation
# --------------------------------------------------------------------------------------
def point_to_sources(gprob, point):
    """utility.adjust_point_units (beat/utility.py:651-675: km -> m for the location variables) + utility.split_point
    (:678-770: one dict per source) + update_source (:773-797; ``duration`` goes to the STF).  ``time`` is relative to
    the event origin."""
    n = int(gprob.get("n_sources", 1))
    out = []
    for s in range(n):
        src = {k: float(np.asarray(v).ravel()[s]) for k, v in point.items() if k not in ("hypers", "time_shifts")}
        for k in ("east_shift", "north_shift", "depth"):
            src[k] *= KM
        out.append(src)
    return out


def point_to_source(gprob, point):
    """Single-source problems: the one source of ``point_to_sources``."""
    return point_to_sources(gprob, point)[0]


def geometry_synthetics(gprob, point, iw=0):
    """heart.seis_synthetics(..., outmode='array') for one wavemap: every source is synthesised and post-processed on
    its own, then the traces are stacked: [nt, ns] float64."""
    wm = gprob["wavemaps"][iw]
    if wm.get("station_idx") is not None:
        # SeisSynthesizer.perform (beat/pytensorf.py:248-252): arrival_times + time_shifts; the hierarchical is gathered per
        # target with wmap.station_correction_idxs (beat/models/seismic.py:781-784).  Windows, taper and chop follow.
        shifts = np.asarray(point["time_shifts"], dtype=np.float64)[wm["station_idx"]]
        wm = dict(wm, arrival_times=np.asarray(wm["arrival_times"]) + shifts)
    outstack = None
    for src in point_to_sources(gprob, point):                  # engine.process iterates sources, then targets (:3676)
        rows = []
        for t in range(wm["nt"]):
            raw, itmin = seismogram(gprob, wm, t, src)
            rows.append(post_process(wm, t, raw, itmin))
        synths = np.vstack(rows)
        outstack = synths if outstack is None else outstack + synths       # heart.py:3719-3724
    return outstack


def geometry_seismic_eval(gprob, point, return_synth=False):
    """One evaluation of the geometry-mode seismic composite (beat/models/seismic.py:737-837): per wavemap
    synthetics -> residuals = data - synths (:819) -> multivariate_normal_chol (:821-827).  Returns logpts of all
    datasets concatenated (:837)."""
    hypers = np.asarray(point["hypers"], dtype=np.float64)
    out, synths = [], []
    for iw, wm in enumerate(gprob["wavemaps"]):
        s = geometry_synthetics(gprob, point, iw)
        synths.append(s)
        res = wm["data"] - s
        out.append(mvn_chol_logpts(res, wm["U"], wm["slog_pdet"], wm["nsamples"], hypers[wm["hyper_idx"]]))
    logpts = np.concatenate(out)
    return (logpts, synths) if return_synth else logpts
