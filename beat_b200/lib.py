"""
ctypes binding of ``libbeatgpu.so`` (C-ABI declared in ``include/beatgpu.h``).

This is the only place Python touches the native library.  There is NO CPU fallback: if the shared
object is missing or cannot be loaded, importing a GPU code path raises ``BeatGpuLibraryError`` with the
build command; if no CUDA device is present, ``Context()`` raises ``BeatGpuError``.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbeatgpu.so")

F32, F64 = 0, 1
NEAREST, MULTILINEAR = 0, 1
INTERPOLATION = {"nearest_neighbor": NEAREST, "multilinear": MULTILINEAR}
MAX_SLIPVARS = 3

E_CUDA, E_ARG, E_INDEX, E_NOTREADY, E_NONFINITE, E_IO = 1, 2, 3, 4, 5, 6


class BeatGpuLibraryError(ImportError):
    pass


class BeatGpuError(RuntimeError):
    """CUDA / call-order failure reported by libbeatgpu."""

    def __init__(self, code, msg):
        super().__init__("libbeatgpu error %d: %s" % (code, msg))
        self.code = code


class GFLibraryError(Exception):
    """Same name as the reference's error type (beat/ffi/base.py:58-59)."""


class Layout(C.Structure):
    _fields_ = [
        ("n_params", C.c_int32),
        ("n_slipvars", C.c_int32),
        ("off_slip", C.c_int32 * MAX_SLIPVARS),
        ("off_durations", C.c_int32),
        ("off_velocities", C.c_int32),
        ("off_nucleation_strike", C.c_int32),
        ("off_nucleation_dip", C.c_int32),
        ("off_time", C.c_int32),
        ("off_hypers", C.c_int32),
        ("n_hypers", C.c_int32),
        ("off_time_shifts", C.c_int32),
        ("n_time_shifts", C.c_int32),
    ]


class GeomLayout(C.Structure):
    """beatgpu_geom_layout (geometry-mode parameter vector -> DC source variables)."""
    _fields_ = [(n, C.c_int32) for n in (
        "n_params", "off_east_shift", "off_north_shift", "off_depth", "off_strike", "off_dip", "off_rake",
        "off_magnitude", "off_time", "off_duration", "off_hypers", "n_hypers", "off_time_shifts", "n_time_shifts", "n_sources")]


GEOM_VARS = ("east_shift", "north_shift", "depth", "strike", "dip", "rake", "magnitude", "time", "duration")
GEOM_MAX_ORDER = 8

# every symbol include/beatgpu.h declares, with its ctypes signature (tests check the .so exports them all)
_P = C.c_void_p
_SIGNATURES = {
    "beatgpu_version": (C.c_int, []),
    "beatgpu_ctx_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "beatgpu_ctx_destroy": (None, [_P]),
    "beatgpu_last_error": (C.c_char_p, [_P]),
    "beatgpu_sync": (C.c_int, [_P]),
    "beatgpu_set_stream": (C.c_int, [_P, _P, C.c_int]),
    "beatgpu_host_register": (C.c_int, [_P, _P, C.c_int64]),
    "beatgpu_host_unregister": (C.c_int, [_P, _P]),
    "beatgpu_device_info": (C.c_int, [_P, C.POINTER(C.c_int), C.c_char_p, C.c_int]),
    "beatgpu_set_fault": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "beatgpu_set_layout": (C.c_int, [_P, C.POINTER(Layout), _P]),
    "beatgpu_add_wavemap": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, C.POINTER(C.c_int)]),
    "beatgpu_upload_gflib": (C.c_int, [_P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, C.c_double, C.c_double,
                                       C.c_double, C.c_double]),
    "beatgpu_alloc_gflib": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, C.c_double, C.c_double, C.c_double,
                                      C.c_double, C.POINTER(_P), C.POINTER(C.c_int64)]),
    "beatgpu_upload_data": (C.c_int, [_P, C.c_int, _P]),
    "beatgpu_update_weights": (C.c_int, [_P, C.c_int, _P, _P, C.c_double]),
    "beatgpu_update_weights_dev": (C.c_int, [_P, C.c_int, _P, _P, C.c_double]),
    "beatgpu_set_geodetic": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, _P, _P, _P]),
    "beatgpu_update_geodetic_weights": (C.c_int, [_P, _P, _P]),
    "beatgpu_set_laplacian": (C.c_int, [_P, _P, C.c_double, C.c_int]),
    "beatgpu_n_outputs": (C.c_int, [_P, C.POINTER(C.c_int)]),
    "beatgpu_fast_sweep_batch": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, _P, _P, _P]),
    "beatgpu_stack_batch": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P]),
    "beatgpu_misfit_batch": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    "beatgpu_misfit_batch_dev": (C.c_int, [_P, C.c_int, C.c_int, _P, _P, C.c_int, _P]),
    "beatgpu_ffi_loglike_batch": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "beatgpu_ffi_loglike_batch_dev": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "beatgpu_ffi_synthetics_batch": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "beatgpu_get_starttimes": (C.c_int, [_P, C.c_int, _P]),
    "beatgpu_index_violations": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "beatgpu_launch_count": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "beatgpu_last_stack_ms": (C.c_int, [_P, C.POINTER(C.c_float)]),
    "beatgpu_stack_ms_accum": (C.c_int, [_P, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "beatgpu_stack_blocking": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64),
                                         C.POINTER(C.c_int64)]),
    "beatgpu_source_hash": (C.c_char_p, []),
    "beatgpu_geom_timeouts": (C.c_int, [_P, C.POINTER(C.c_int64)]),
    "beatgpu_geom_set_source": (C.c_int, [_P, C.POINTER(GeomLayout), _P, C.c_double, C.c_double, C.c_double]),
    "beatgpu_geom_set_stf": (C.c_int, [_P, C.c_int, C.c_double, C.c_int, C.c_double]),
    "beatgpu_geom_upload_store": (C.c_int, [_P, _P, C.c_double, C.c_double, C.c_double, C.c_double, C.c_double, _P, _P, _P,
                                            C.POINTER(C.c_int)]),
    "beatgpu_geom_add_wavemap": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P, _P, _P, _P, _P, C.c_int, C.c_int,
                                           C.c_int, _P, _P, _P, C.c_int, _P, _P, _P, C.POINTER(C.c_int)]),
    "beatgpu_geom_loglike_batch": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "beatgpu_geom_loglike_batch_dev": (C.c_int, [_P, C.c_int, _P, _P, _P]),
    "beatgpu_geom_synthetics_batch": (C.c_int, [_P, C.c_int, C.c_int, _P, _P]),
    "beatgpu_trace_append": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int64, _P, C.c_int]),
    "beatgpu_probe_gather": (C.c_int, [_P, C.c_int, C.c_int64, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float),
                                       C.POINTER(C.c_double)]),
}

STF_TYPES = {"HalfSinusoid": 0, "Boxcar": 1, "Triangular": 2}          # BEATGPU_STF_*

_lib = None


def load():
    """Load libbeatgpu.so (once).  Raises BeatGpuLibraryError if it is not built -- never falls back."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BeatGpuLibraryError(
                "%s not found.  Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(nvcc -gencode arch=compute_100a,code=sm_100a).  beat_b200 has no CPU fallback." % LIB_PATH)
        try:
            lib = C.CDLL(LIB_PATH)
        except OSError as e:  # pragma: no cover
            raise BeatGpuLibraryError("cannot load %s: %s" % (LIB_PATH, e))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def source_hash():
    """Hash of the sources the loaded libbeatgpu.so was built from (beatgpu_source_hash)."""
    return load().beatgpu_source_hash().decode()


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


def _i32(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.int32)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected shape %s, got %s" % (tuple(shape), a.shape))
    return a


def _check_q(q, n_params, what):
    """The C entries trust [B, n_params]: a parameter matrix of another width must never reach them."""
    if q.ndim != 2 or (n_params is not None and q.shape[1] != n_params):
        raise ValueError("%s: q must be [B, %s], got %s" % (what, n_params, q.shape))
    return q


def trace_append(dir_path, chain_offset, records, n_threads=4):
    """Append the step-major record block ``records`` ([n_steps, n_chains] structured, or [n_steps, n_chains, k]) to the
    chain files of ``dir_path`` (native writev + threads, beatgpu_trace_append; no CUDA context needed).  Raises OSError."""
    lib = load()
    records = np.ascontiguousarray(records)
    n_steps, n_chains = records.shape[0], records.shape[1]
    rec_bytes = records.dtype.itemsize * int(np.prod(records.shape[2:], dtype=np.int64))
    rc = lib.beatgpu_trace_append(os.fsencode(dir_path), int(chain_offset), int(n_chains), int(n_steps), int(rec_bytes),
                                  records.ctypes.data_as(C.c_void_p), int(n_threads))
    if rc:
        msg = lib.beatgpu_last_error(None).decode()
        if rc == E_IO:
            raise OSError(msg)
        if rc == E_ARG:
            raise ValueError(msg)
        raise BeatGpuError(rc, msg)


class Context:
    """One libbeatgpu context = one (process, GPU).  Thin, typed wrappers over the C entry points."""

    def __init__(self, device=0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.beatgpu_ctx_create(int(device), C.byref(h))
        if rc:
            raise BeatGpuError(rc, self._lib.beatgpu_last_error(None).decode())
        self._h = h
        self.device = int(device)
        # sizes the C entries trust (they read nt*ns, nt*ns*ns, canon_len ... elements from raw pointers): every array
        # is checked against them here before its pointer crosses the ABI
        self._np_total = self._nsf = None
        self._canon_len = None
        self._n_hypers = 0
        self._n_slipvars = None
        self._wm = {}               # wavemap id -> (nt, ns)
        self._geo = None            # (nobs, [n_i per dataset])

    def close(self):
        if getattr(self, "_h", None):
            self._lib.beatgpu_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            msg = self._lib.beatgpu_last_error(self._h).decode()
            if rc == E_INDEX:
                raise IndexError(msg)            # what numpy/pytensor raise in the reference
            if rc == E_ARG:
                raise ValueError(msg)
            raise BeatGpuError(rc, msg)

    # ------------------------------------------------------------------ context
    def sync(self):
        self._check(self._lib.beatgpu_sync(self._h))

    def set_stream(self, cuda_stream_ptr, external=True):
        """external=True: enqueue on the given CUDA stream handle (0 = default stream); False: private stream."""
        self._check(self._lib.beatgpu_set_stream(self._h, C.c_void_p(cuda_stream_ptr or 0), 1 if external else 0))

    def pin(self, array):
        """Page-lock a numpy array in place (cudaHostRegister) for full-rate H2D/D2H in the host-pointer entries."""
        if not array.flags["C_CONTIGUOUS"]:
            raise ValueError("only C-contiguous arrays can be pinned")
        self._check(self._lib.beatgpu_host_register(self._h, array.ctypes.data_as(C.c_void_p), array.nbytes))
        return array

    def unpin(self, array):
        self._check(self._lib.beatgpu_host_unregister(self._h, array.ctypes.data_as(C.c_void_p)))

    def device_info(self):
        n = C.c_int()
        buf = C.create_string_buffer(256)
        self._check(self._lib.beatgpu_device_info(self._h, C.byref(n), buf, 256))
        return n.value, buf.value.decode()

    # ------------------------------------------------------------------ operands
    def set_fault(self, n_patch_dip, n_patch_strike, patch_size):
        nd, ns, ps = _i32(n_patch_dip), _i32(n_patch_strike), _f64(patch_size)
        if nd.ndim != 1 or ns.shape != nd.shape or ps.shape != nd.shape:
            raise ValueError("set_fault: n_patch_dip, n_patch_strike and patch_size must be 1-d arrays of one length")
        self._check(self._lib.beatgpu_set_fault(self._h, len(nd), _ptr(nd), _ptr(ns), _ptr(ps)))
        self._nsf = len(nd)
        self._np_total = int((nd.astype(np.int64) * ns).sum())

    def set_layout(self, layout: Layout, fixed=None):
        if self._np_total is None:
            raise BeatGpuError(E_NOTREADY, "set_layout: call set_fault first")
        canon = (layout.n_slipvars + 2) * self._np_total + 3 * self._nsf + layout.n_hypers + layout.n_time_shifts
        fx = None if fixed is None else _f64(fixed, (canon,))      # the C side copies canon_len doubles
        self._check(self._lib.beatgpu_set_layout(self._h, C.byref(layout), _ptr(fx)))
        self._n_params = layout.n_params
        self._canon_len, self._n_hypers, self._n_slipvars = canon, layout.n_hypers, layout.n_slipvars

    def add_wavemap(self, n_targets, n_samples, interpolation, station_idx, hyper_idx, nsamples):
        nt = int(n_targets)
        st = None if station_idx is None else _i32(station_idx, (nt,))
        hi, nsm = _i32(hyper_idx, (nt,)), _i32(nsamples, (nt,))
        wid = C.c_int()
        self._check(self._lib.beatgpu_add_wavemap(self._h, int(n_targets), int(n_samples),
                                                  INTERPOLATION.get(interpolation, interpolation),
                                                  _ptr(st), _ptr(hi), _ptr(nsm), C.byref(wid)))
        self._wm[wid.value] = (nt, int(n_samples))
        return wid.value

    def _wm_shape(self, wmap, what):
        if wmap not in self._wm:
            raise ValueError("%s: unknown wavemap id %r" % (what, wmap))
        return self._wm[wmap]

    def upload_gflib(self, wmap, slipvar, traces, store_dtype, dur_min, dur_step, st_min, st_step):
        if traces.dtype == np.float64:
            src = F64
        elif traces.dtype == np.float32:
            src = F32
        else:
            raise GFLibraryError("GF library dtype %s not supported" % traces.dtype)
        if traces.ndim != 5 or not traces.flags["C_CONTIGUOUS"]:
            raise GFLibraryError("GF library must be a C-contiguous 5-d array (targets, patches, durations, starttimes, samples)")
        nt, ns = self._wm_shape(wmap, "upload_gflib")
        if traces.shape[0] != nt or traces.shape[4] != ns:
            raise GFLibraryError("GF library is %s, wavemap %d has %d targets x %d samples" % (traces.shape, wmap, nt, ns))
        dims = np.asarray(traces.shape, dtype=np.int64)
        self._check(self._lib.beatgpu_upload_gflib(self._h, wmap, slipvar, traces.ctypes.data_as(C.c_void_p), src,
                                                   store_dtype, _ptr(dims), dur_min, dur_step, st_min, st_step))

    def alloc_gflib(self, wmap, slipvar, store_dtype, dims, dur_min, dur_step, st_min, st_step):
        dims = np.asarray(dims, dtype=np.int64)
        p, ld = C.c_void_p(), C.c_int64()
        self._check(self._lib.beatgpu_alloc_gflib(self._h, wmap, slipvar, store_dtype, _ptr(dims), dur_min, dur_step,
                                                  st_min, st_step, C.byref(p), C.byref(ld)))
        return p.value, ld.value

    def upload_data(self, wmap, data):
        d = _f64(data, self._wm_shape(wmap, "upload_data"))
        self._check(self._lib.beatgpu_upload_data(self._h, wmap, _ptr(d)))

    def update_weights_dev(self, wmap, U_dev_ptr, slog_pdet_dev_ptr, band_rtol=-1.0):
        """Weights already on the device (e.g. torch tensors from covariance.weights_from_residuals_device)."""
        self._check(self._lib.beatgpu_update_weights_dev(self._h, wmap, C.c_void_p(U_dev_ptr), C.c_void_p(slog_pdet_dev_ptr),
                                                         band_rtol))

    def update_weights(self, wmap, U, slog_pdet, band_rtol=-1.0):
        nt, ns = self._wm_shape(wmap, "update_weights")
        U, lp = _f64(U, (nt, ns, ns)), _f64(slog_pdet, (nt,))
        self._check(self._lib.beatgpu_update_weights(self._h, wmap, _ptr(U), _ptr(lp), band_rtol))

    def set_geodetic(self, slices, G_list, data, odw, U_list, slog_pdet, nsamples, hyper_idx):
        if self._n_slipvars is None:
            raise BeatGpuError(E_NOTREADY, "set_geodetic: call set_fault and set_layout first")
        lo = _i32([s[0] for s in slices])
        hi = _i32([s[1] for s in slices])
        nds = len(slices)
        data = _f64(data)
        if data.ndim != 1:
            raise ValueError("set_geodetic: data must be the concatenated observation vector [nobs]")
        nobs = data.shape[0]
        odw = _f64(odw, (nobs,))
        if len(G_list) != self._n_slipvars:
            raise ValueError("set_geodetic: %d libraries for %d slip variables" % (len(G_list), self._n_slipvars))
        Gs = [_f64(g, (self._np_total, nobs)) for g in G_list]
        arr = (C.c_void_p * len(Gs))(*[g.ctypes.data for g in Gs])
        if len(U_list) != nds:
            raise ValueError("set_geodetic: %d weight matrices for %d datasets" % (len(U_list), nds))
        sizes = [int(h - l) for l, h in zip(lo, hi)]
        Ucat = np.concatenate([_f64(u, (n, n)).ravel() for u, n in zip(U_list, sizes)])
        lp, nsm, hix = _f64(slog_pdet, (nds,)), _i32(nsamples, (nds,)), _i32(hyper_idx, (nds,))
        self._geo = (nobs, sizes)
        self._check(self._lib.beatgpu_set_geodetic(self._h, len(data), len(slices), _ptr(lo), _ptr(hi),
                                                   C.cast(arr, C.c_void_p), _ptr(data), _ptr(odw), _ptr(Ucat), _ptr(lp),
                                                   _ptr(nsm), _ptr(hix)))

    def update_geodetic_weights(self, U_list, slog_pdet):
        if self._geo is None:
            raise BeatGpuError(E_NOTREADY, "update_geodetic_weights: geodetic composite not set")
        sizes = self._geo[1]
        if len(U_list) != len(sizes):
            raise ValueError("update_geodetic_weights: %d weight matrices for %d datasets" % (len(U_list), len(sizes)))
        Ucat = np.concatenate([_f64(u, (n, n)).ravel() for u, n in zip(U_list, sizes)])
        lp = _f64(slog_pdet, (len(sizes),))
        self._check(self._lib.beatgpu_update_geodetic_weights(self._h, _ptr(Ucat), _ptr(lp)))

    def set_laplacian(self, L, sdet, hyper_idx):
        if self._np_total is None:
            raise BeatGpuError(E_NOTREADY, "set_laplacian: call set_fault and set_layout first")
        L = _f64(L, (self._np_total, self._np_total))
        self._check(self._lib.beatgpu_set_laplacian(self._h, _ptr(L), float(sdet), int(hyper_idx)))

    def n_outputs(self):
        n = C.c_int()
        self._check(self._lib.beatgpu_n_outputs(self._h, C.byref(n)))
        return n.value

    # ------------------------------------------------------------------ hot path
    def fast_sweep_batch(self, subfault, slowness, nuc_dip_idx, nuc_strike_idx, return_iters=False):
        s = _f64(slowness)
        if s.ndim != 2:
            raise ValueError("slowness must be [B, npatches_subfault]")
        B = s.shape[0]
        di, si = _i32(nuc_dip_idx), _i32(nuc_strike_idx)
        if di.shape != (B,) or si.shape != (B,):
            raise ValueError("nucleation indices must be [B]")
        out = np.empty_like(s)
        it = np.zeros(B, dtype=np.int32) if return_iters else None
        self._check(self._lib.beatgpu_fast_sweep_batch(self._h, subfault, B, _ptr(s), _ptr(di), _ptr(si), _ptr(out), _ptr(it)))
        return (out, it) if return_iters else out

    def stack_batch(self, wmap, durations, starttimes, slips, nt, ns):
        d, st, sl = _f64(durations), _f64(starttimes), _f64(slips)
        if (nt, ns) != self._wm_shape(wmap, "stack_batch"):
            raise ValueError("stack_batch: wavemap %d is %s, not (%d, %d)" % (wmap, self._wm[wmap], nt, ns))
        if d.ndim != 2 or sl.ndim != 3:
            raise ValueError("stack_batch: durations [B, np], starttimes [B, nt, np], slips [nvar, B, np]")
        B, npatch = d.shape
        nvar = sl.shape[0]
        if st.shape != (B, nt, npatch) or sl.shape != (nvar, B, npatch) or (self._np_total is not None and npatch != self._np_total):
            raise ValueError("stack_batch: inconsistent shapes")
        out = np.empty((B, nt, ns))
        self._check(self._lib.beatgpu_stack_batch(self._h, wmap, B, nvar, _ptr(d), _ptr(st), _ptr(sl), _ptr(out)))
        return out

    def misfit_batch(self, wmap, residuals, hypers):
        r, h = _f64(residuals), _f64(hypers)
        if r.ndim != 3 or r.shape[1:] != self._wm_shape(wmap, "misfit_batch"):
            raise ValueError("misfit_batch: residuals must be [B, %d, %d], got %s" % (self._wm[wmap] + (r.shape,)))
        B, nt, _ = r.shape
        if h.ndim != 2 or h.shape[0] != B:
            raise ValueError("hypers must be [B, n_hypers]")
        out = np.empty((B, nt))
        self._check(self._lib.beatgpu_misfit_batch(self._h, wmap, B, _ptr(r), _ptr(h), h.shape[1], _ptr(out)))
        return out

    def misfit_batch_dev(self, wmap, B, resid_ptr, hyp_ptr, n_hypers, logpts_ptr):
        self._check(self._lib.beatgpu_misfit_batch_dev(self._h, wmap, int(B), C.c_void_p(resid_ptr), C.c_void_p(hyp_ptr),
                                                       int(n_hypers), C.c_void_p(logpts_ptr)))

    def ffi_loglike_batch(self, q, logpts=None, like=None):
        """Host-pointer entry (copies in and out, synchronises).  q [B, n_params] float64 C-contiguous."""
        q = _check_q(_f64(q), getattr(self, "_n_params", None), "ffi_loglike_batch")
        B = q.shape[0]
        n_out = self.n_outputs()
        if logpts is None:
            logpts = np.empty((B, n_out))
        if like is None:
            like = np.empty(B)
        self._check(self._lib.beatgpu_ffi_loglike_batch(self._h, B, _ptr(q), _ptr(logpts), _ptr(like)))
        return logpts, like

    def ffi_loglike_batch_ptr(self, B, q_ptr, logpts_ptr, like_ptr):
        """Host-pointer entry on raw addresses (e.g. pinned torch tensors)."""
        self._check(self._lib.beatgpu_ffi_loglike_batch(self._h, int(B), C.c_void_p(q_ptr), C.c_void_p(logpts_ptr),
                                                        C.c_void_p(like_ptr)))

    def ffi_loglike_batch_dev(self, B, q_dev_ptr, logpts_dev_ptr, like_dev_ptr):
        """Device-pointer entry: enqueues on the ctx stream, no synchronisation."""
        self._check(self._lib.beatgpu_ffi_loglike_batch_dev(self._h, int(B), C.c_void_p(q_dev_ptr), C.c_void_p(logpts_dev_ptr),
                                                            C.c_void_p(like_dev_ptr or 0)))

    def ffi_synthetics_batch(self, wmap, q, nt, ns):
        q = _check_q(_f64(q), getattr(self, "_n_params", None), "ffi_synthetics_batch")
        out = np.empty((q.shape[0], nt, ns))
        self._check(self._lib.beatgpu_ffi_synthetics_batch(self._h, wmap, q.shape[0], _ptr(q), _ptr(out)))
        return out

    def get_starttimes(self, B, npatches):
        out = np.empty((B, npatches))
        self._check(self._lib.beatgpu_get_starttimes(self._h, B, _ptr(out)))
        return out

    def index_violations(self):
        n = C.c_int64()
        self._check(self._lib.beatgpu_index_violations(self._h, C.byref(n)))
        return n.value

    def launch_count(self):
        n = C.c_int64()
        self._check(self._lib.beatgpu_launch_count(self._h, C.byref(n)))
        return n.value

    def last_stack_ms(self):
        ms = C.c_float()
        self._check(self._lib.beatgpu_last_stack_ms(self._h, C.byref(ms)))
        return ms.value

    def stack_ms_accum(self, reset=True):
        """(sum of the stack + misfit kernel durations [ms], evaluations covered) since the last reset."""
        ms, n = C.c_double(), C.c_int64()
        self._check(self._lib.beatgpu_stack_ms_accum(self._h, 1 if reset else 0, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def stack_blocking(self, wmap, n_slipvars=None):
        """L2 blocking of the stacking pass: dict(chunk_patches, n_chunks, chunk_bytes, l2_bytes)."""
        c, n, cb, l2 = C.c_int(), C.c_int(), C.c_int64(), C.c_int64()
        nv = int(n_slipvars if n_slipvars is not None else (self._n_slipvars or 1))
        self._check(self._lib.beatgpu_stack_blocking(self._h, wmap, nv, C.byref(c), C.byref(n), C.byref(cb), C.byref(l2)))
        return dict(chunk_patches=c.value, n_chunks=n.value, chunk_bytes=cb.value, l2_bytes=l2.value)

    def geom_timeouts(self):
        n = C.c_int64()
        self._check(self._lib.beatgpu_geom_timeouts(self._h, C.byref(n)))
        return n.value

    def probe_gather(self, mode, ws_bytes, row_bytes=480, rows_per_warp=4096, n_launch=3):
        """Diagnostics: measured row-gather bandwidth [GB/s].  mode 0 LDG.128, 1 TMA per row + smem read, 2 TMA per row only,
        3 / 4 batched TMA (16 rows per mbarrier) with / without smem read, 5 rows from local shared memory, 6 / 7 / 8 rows
        from the shared memory of a 2 / 4 / 8-CTA cluster (DSMEM)."""
        ms, nbytes = C.c_float(), C.c_double()
        self._check(self._lib.beatgpu_probe_gather(self._h, mode, int(ws_bytes), row_bytes, rows_per_warp, n_launch,
                                                   C.byref(ms), C.byref(nbytes)))
        return nbytes.value / (ms.value * 1e-3) / 1e9

    # ------------------------------------------------------------------ geometry mode (config 2)
    def geom_set_source(self, layout, fixed, event_lat, event_lon, stf_anchor=-1.0):
        fx = None if fixed is None else _f64(fixed)
        self._check(self._lib.beatgpu_geom_set_source(self._h, C.byref(layout), _ptr(fx), float(event_lat), float(event_lon),
                                                      float(stf_anchor)))
        self._geom_n_params = layout.n_params

    def geom_set_stf(self, stf_type, stf_anchor=-1.0, off_peak_ratio=-1, fixed_peak_ratio=0.5):
        """stf_type: a key of the reference's stf_catalog (beat/sources.py:723-729): Boxcar, Triangular, HalfSinusoid."""
        if stf_type not in STF_TYPES:
            raise ValueError("stf_type %r not in %s (beat/config.py:1359-1365)" % (stf_type, sorted(STF_TYPES)))
        self._check(self._lib.beatgpu_geom_set_stf(self._h, STF_TYPES[stf_type], float(stf_anchor), int(off_peak_ratio),
                                                   float(fixed_peak_ratio)))

    def geom_upload_store(self, traces, itmin, nsamples, z0, dz, x0, dx, deltat):
        traces = np.ascontiguousarray(traces, dtype=np.float32)
        if traces.ndim != 4:
            raise ValueError("GF store traces must be (n_depths, n_distances, n_components, n_samples)")
        itmin = np.ascontiguousarray(itmin, dtype=np.int32)
        nsamples = np.ascontiguousarray(nsamples, dtype=np.int32)
        if itmin.shape != traces.shape[:3] or nsamples.shape != traces.shape[:3]:
            raise ValueError("itmin / nsamples must have one entry per record")
        dims = np.asarray(traces.shape, dtype=np.int64)
        sid = C.c_int()
        self._check(self._lib.beatgpu_geom_upload_store(self._h, _ptr(dims), z0, dz, x0, dx, deltat, _ptr(traces), _ptr(itmin),
                                                        _ptr(nsamples), C.byref(sid)))
        return sid.value

    def geom_add_wavemap(self, store_id, ns, interpolation, lats, lons, azimuths, dips, arrival_times, taper_abcd,
                         chop_bounds, sections, hyper_idx, nsamples, station_idx=None):
        """sections: [(b, a, demean), ...] as scipy.signal.butter returns them (demean only on the first)."""
        lats, lons, az, dp, at = (_f64(x) for x in (lats, lons, azimuths, dips, arrival_times))
        nt = lats.size
        abcd = _f64(taper_abcd, (4,))
        lo, hi = ("abcd".index(c) for c in chop_bounds)
        nsec = len(sections)
        order = np.zeros(max(nsec, 1), dtype=np.int32)
        sb = np.zeros((max(nsec, 1), GEOM_MAX_ORDER + 1))
        sa = np.zeros((max(nsec, 1), GEOM_MAX_ORDER + 1))
        for i, (b, a, demean) in enumerate(sections):
            b, a = np.atleast_1d(b), np.atleast_1d(a)
            if demean and i > 0:
                raise NotImplementedError("demeaning is supported before the first filter section only")
            if max(b.size, a.size) - 1 > GEOM_MAX_ORDER:
                raise NotImplementedError("IIR sections up to order %d" % GEOM_MAX_ORDER)
            order[i] = max(b.size, a.size) - 1
            sb[i, :b.size] = b
            sa[i, :a.size] = a
        demean_first = int(bool(nsec and sections[0][2]))
        hidx, nsm = _i32(hyper_idx), _i32(nsamples)
        sidx = None if station_idx is None else _i32(station_idx)
        wid = C.c_int()
        self._check(self._lib.beatgpu_geom_add_wavemap(
            self._h, store_id, nt, ns, INTERPOLATION[interpolation], _ptr(lats), _ptr(lons), _ptr(az), _ptr(dp), _ptr(at),
            _ptr(abcd), lo, hi, nsec, _ptr(order), _ptr(sb), _ptr(sa), demean_first, _ptr(sidx), _ptr(hidx), _ptr(nsm),
            C.byref(wid)))
        self._wm[wid.value] = (int(nt), int(ns))
        return wid.value

    def geom_loglike_batch(self, Q):
        Q = _check_q(_f64(Q), getattr(self, "_geom_n_params", None), "geom_loglike_batch")
        B = Q.shape[0]
        n_out = self.n_outputs()
        logpts, like = np.empty((B, n_out)), np.empty(B)
        self._check(self._lib.beatgpu_geom_loglike_batch(self._h, B, _ptr(Q), _ptr(logpts), _ptr(like)))
        return logpts, like

    def geom_loglike_batch_ptr(self, B, q_ptr, logpts_ptr, like_ptr):
        self._check(self._lib.beatgpu_geom_loglike_batch(self._h, int(B), C.c_void_p(q_ptr), C.c_void_p(logpts_ptr),
                                                         C.c_void_p(like_ptr)))

    def geom_loglike_batch_dev(self, B, q_ptr, logpts_ptr, like_ptr):
        self._check(self._lib.beatgpu_geom_loglike_batch_dev(self._h, int(B), C.c_void_p(q_ptr), C.c_void_p(logpts_ptr),
                                                             C.c_void_p(like_ptr or 0)))

    def geom_synthetics_batch(self, wmap, Q, nt, ns):
        Q = _check_q(_f64(Q), getattr(self, "_geom_n_params", None), "geom_synthetics_batch")
        out = np.empty((Q.shape[0], nt, ns))
        self._check(self._lib.beatgpu_geom_synthetics_batch(self._h, wmap, Q.shape[0], _ptr(Q), _ptr(out)))
        return out
