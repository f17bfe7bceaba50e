"""
Host-side mirrors of the reference's operator interface for the hot path, backed by libbeatgpu.

Same names, argument meaning and error behaviour as the reference so the parity tests read like the
reference's own (test/test_fastsweep.py, test/test_ffi.py, test/test_models.py):

  * ``Sweeper``                   <- beat/pytensorf.py:410-503  (a pytensor ``Op`` when pytensor is importable)
  * ``SeismicGFLibrary.stack_all``<- beat/ffi/base.py:607-709
  * ``GeodeticGFLibrary.stack_all`` is part of the fused path only (beat/ffi/base.py:292-305)
  * ``multivariate_normal_chol``  <- beat/models/distributions.py:72-140
  * ``FFILogLike``                <- the sub-graph of SeismicDistributerComposite.get_formula
                                     (beat/models/seismic.py:1253-1349) as ONE Op

Every entry accepts a leading chain axis (the reference has none: it evaluates one chain per call,
beat/sampler/metropolis.py:349); without it the call is the reference's single-chain call.
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .lib import Context, GFLibraryError

try:  # the reference's Op base class, when available (absent in the build container)
    import pytensor.tensor as _tt
    from pytensor.graph import Apply as _Apply
    _OpBase = _tt.Op
    HAVE_PYTENSOR = True
except Exception:  # pragma: no cover - exercised on boxes without pytensor
    _tt = None
    _Apply = None
    HAVE_PYTENSOR = False

    class _OpBase(object):
        """Minimal stand-in with the Op calling convention: ``perform(node, inputs, output_storage)``."""

        def __call__(self, *inputs):
            out = [[None] for _ in range(getattr(self, "n_outputs", 1))]
            self.perform(None, [np.asarray(i) for i in inputs], out)
            return out[0][0] if len(out) == 1 else [o[0] for o in out]


class Sweeper(_OpBase):
    """GPU implementation of the fast-sweeping Op (reference: beat/pytensorf.py:410-503).

    Parameters are the reference's: patch_size [km], n_patch_dip, n_patch_strike, implementation.
    ``implementation`` must be "cuda" (or the reference's "c", which this Op replaces); anything else raises
    NotImplementedError exactly like the reference (pytensorf.py:494-498)."""

    __props__ = ("patch_size", "n_patch_dip", "n_patch_strike", "implementation")

    def __init__(self, patch_size, n_patch_dip, n_patch_strike, implementation="cuda", device=0):
        self.patch_size = np.float64(patch_size)
        self.n_patch_dip = int(n_patch_dip)
        self.n_patch_strike = int(n_patch_strike)
        self.implementation = implementation
        self._device = device
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            ctx = Context(self._device)
            ctx.set_fault([self.n_patch_dip], [self.n_patch_strike], [self.patch_size])
            self._ctx = ctx
        return self._ctx

    def make_node(self, *inputs):
        """As the reference (pytensorf.py:432-441): inputs wrapped with ``as_tensor_variable``, output type taken from a
        zero array of ``infer_shape``; a slowness MATRIX [B, np] (the batch axis this Op adds) declares a matrix output."""
        inlist = [_tt.as_tensor_variable(i) for i in inputs]
        shape = self.infer_shape()[0]
        if getattr(inlist[0], "ndim", 1) == 2:
            shape = (1,) + tuple(shape)
        outv = _tt.as_tensor_variable(np.zeros(shape))
        return _Apply(self, inlist, [outv.type()])

    def perform(self, node, inputs, output):
        slownesses, nuc_dip, nuc_strike = inputs
        if self.implementation not in ("cuda", "c"):
            raise NotImplementedError("Fast sweeping for implementation %s not implemented!" % self.implementation)
        s = np.asarray(slownesses, dtype=np.float64)
        if s.dtype != np.float64:
            raise AttributeError("array of unexpected type")
        single = s.ndim == 1
        s2 = s.reshape(1, -1) if single else s
        if s2.shape[1] != self.n_patch_dip * self.n_patch_strike:
            raise AttributeError("array is of unexpected size")     # fast_sweep_ext.c:36-39
        nd = np.atleast_1d(np.asarray(nuc_dip)).astype(np.int32)
        ns = np.atleast_1d(np.asarray(nuc_strike)).astype(np.int32)
        out = self._context().fast_sweep_batch(0, s2, nd, ns)
        output[0][0] = out[0] if single else out

    def infer_shape(self, fgraph=None, node=None, input_shapes=None):
        n = self.n_patch_dip * self.n_patch_strike                          # pytensorf.py:502-503
        if input_shapes and len(input_shapes[0]) == 2:
            return [(input_shapes[0][0], n)]
        return [(n,)]


def _times2idxs(x, x_min, x_step, interpolation):
    x = np.asarray(x, dtype=np.float64)
    if interpolation == "nearest_neighbor":
        return np.round((x - x_min) / x_step).astype("int16"), None
    elif interpolation == "multilinear":
        d = (x - x_min) / x_step
        c = np.ceil(d).astype("int16")
        return c, c - d
    raise NotImplementedError("Interpolation scheme %s not implemented!" % interpolation)


class SeismicGFLibrary(object):
    """Device-resident seismic GF library with the reference's stacking interface (beat/ffi/base.py:322-802).

    ``traces``: dict slip-variable -> ndarray (ntargets, npatches, ndurations, nstarttimes, nsamples) or a single
    ndarray for one component."""

    def __init__(self, traces, duration_min, duration_sampling, starttime_min, starttime_sampling,
                 interpolation_default="nearest_neighbor", store_dtype="float64", device=0):
        if not isinstance(traces, dict):
            traces = {"uparr": traces}
        self.components = list(traces.keys())
        first = traces[self.components[0]]
        if first.ndim != 5:
            raise GFLibraryError("Seismic Greens Function Library is not set up!")
        self.dimensions = tuple(first.shape)
        self.duration_min, self.duration_sampling = float(duration_min), float(duration_sampling)
        self.starttime_min, self.starttime_sampling = float(starttime_min), float(starttime_sampling)
        self._traces = traces
        self._store = _lib.F32 if store_dtype in ("float32", _lib.F32) else _lib.F64
        self._device = device
        self._ctxs = {}

    ntargets = property(lambda self: self.dimensions[0])
    npatches = property(lambda self: self.dimensions[1])
    ndurations = property(lambda self: self.dimensions[2])
    nstarttimes = property(lambda self: self.dimensions[3])
    nsamples = property(lambda self: self.dimensions[4])

    # index helpers (host side, for inspection / library filling; the kernels do the same arithmetic per tap)
    def starttimes2idxs(self, starttimes, interpolation="nearest_neighbor"):
        """beat/ffi/base.py:486-521."""
        return _times2idxs(starttimes, self.starttime_min, self.starttime_sampling, interpolation)

    def durations2idxs(self, durations, interpolation="nearest_neighbor"):
        """beat/ffi/base.py:535-568."""
        return _times2idxs(durations, self.duration_min, self.duration_sampling, interpolation)

    def idxs2durations(self, idxs):
        return idxs * self.duration_sampling + self.duration_min          # base.py:523-527

    def idxs2starttimes(self, idxs):
        return idxs * self.starttime_sampling + self.starttime_min        # base.py:529-533

    def _context(self, interpolation):
        if interpolation not in _lib.INTERPOLATION:
            raise NotImplementedError("Interpolation scheme %s not implemented!" % interpolation)   # base.py:700-703
        if interpolation not in self._ctxs:
            ctx = Context(self._device)
            ctx.set_fault([1], [self.npatches], [1.0])
            nt, ns = self.ntargets, self.nsamples
            wid = ctx.add_wavemap(nt, ns, interpolation, None, np.zeros(nt, np.int32), np.full(nt, ns, np.int32))
            for iv, comp in enumerate(self.components):
                ctx.upload_gflib(wid, iv, np.ascontiguousarray(self._traces[comp]), self._store, self.duration_min,
                                 self.duration_sampling, self.starttime_min, self.starttime_sampling)
            self._ctxs[interpolation] = (ctx, wid)
        return self._ctxs[interpolation]

    def stack_all(self, durations, starttimes, slips, targetidxs=None, patchidxs=None, interpolation="nearest_neighbor",
                  component=None):
        """Stack all patches for all targets (reference signature, base.py:607-615).

        durations [np] or [B, np]; starttimes [nt, np] or [B, nt, np]; slips [np] or [B, np] (one component, like the
        reference) or dict component -> slips to sum several components in one pass.  Returns [nt, ns] or [B, nt, ns]."""
        if targetidxs is None:
            raise ValueError("Target indexes have to be defined!")            # base.py:630-631
        if patchidxs is not None and len(patchidxs) != self.npatches:
            raise NotImplementedError("patch subsets are not supported by the GPU library")
        ctx, wid = self._context(interpolation)
        d = np.asarray(durations, dtype=np.float64)
        single = d.ndim == 1
        d = np.atleast_2d(d)
        B = d.shape[0]
        st = np.asarray(starttimes, dtype=np.float64).reshape(B, self.ntargets, self.npatches)
        if isinstance(slips, dict):
            comps = list(slips.keys())
            sl = np.stack([np.asarray(slips[c], dtype=np.float64).reshape(B, self.npatches) for c in comps])
        else:
            comps = [component or self.components[0]]
            sl = np.asarray(slips, dtype=np.float64).reshape(1, B, self.npatches)
        order = [self.components.index(c) for c in comps]
        if order != list(range(len(order))):
            # the kernel sums components 0..n-1 of the context; pad leading unused components with zero slip
            full = np.zeros((max(order) + 1, B, self.npatches))
            for i, o in enumerate(order):
                full[o] = sl[i]
            sl = full
        out = ctx.stack_batch(wid, d, st, sl, self.ntargets, self.nsamples)
        return out[0] if single else out


def multivariate_normal_chol(datasets, weights, hyperparams, residuals, hp_specific=False, device=0):
    """Batched GPU ``multivariate_normal_chol`` (reference: beat/models/distributions.py:72-140).

    datasets: objects with ``.samples``, ``.typ`` and ``.covariance.slog_pdet`` / ``.covariance.log_pdet`` as in the
    reference; weights: list of (ns, ns) arrays (``chol_inverse``); hyperparams: dict name -> scalar / [n_t] /
    [B] / [B, n_t]; residuals [n_t, ns] or [B, n_t, ns].  All datasets must have the same number of samples
    (one wavemap).  Returns logpts [n_t] or [B, n_t]."""
    n_t = len(datasets)
    r = np.asarray(residuals, dtype=np.float64)
    single = r.ndim == 2
    if single:
        r = r[None]
    B, _, ns = r.shape
    U = np.stack([np.asarray(w, dtype=np.float64) for w in weights])
    lp = np.array([float(np.asarray(getattr(ds.covariance, "slog_pdet", None) if getattr(ds.covariance, "slog_pdet", None)
                                    is not None else ds.covariance.log_pdet)) for ds in datasets])
    # hypers block: one column per dataset
    H = np.zeros((B, n_t))
    counts = {}
    for i, ds in enumerate(datasets):
        name = "_".join(("h", ds.typ))                                  # distributions.py:24-25
        hp = np.asarray(hyperparams[name], dtype=np.float64)
        if hp_specific:                                                # one entry per dataset of that type (:123-126)
            k = counts.get(name, 0)
            counts[name] = k + 1
            H[:, i] = hp[..., k]
        elif hp.ndim == 0 or hp.size == 1:                             # one value for all chains and datasets
            H[:, i] = hp.reshape(())
        elif hp.ndim == 1 and hp.shape[0] == B and B != n_t:           # [B]: per chain
            H[:, i] = hp
        elif hp.ndim == 1 and hp.shape[0] == n_t and B != n_t:         # [n_t]: per dataset
            H[:, i] = hp[i]
        elif hp.ndim == 2 and hp.shape == (B, n_t):                    # [B, n_t]
            H[:, i] = hp[:, i]
        elif hp.ndim == 2 and hp.shape == (B, 1):
            H[:, i] = hp[:, 0]
        else:
            raise ValueError("hyperparameter %s: shape %s is ambiguous or does not match B=%d chains / n_t=%d datasets "
                             "(pass a scalar, [B, 1] or [B, n_t])" % (name, hp.shape, B, n_t))
    ctx = Context(device)
    try:
        wid = ctx.add_wavemap(n_t, ns, "nearest_neighbor", None, np.arange(n_t, dtype=np.int32),
                              np.array([ds.samples for ds in datasets], dtype=np.int32))
        ctx.update_weights(wid, U, lp)
        out = ctx.misfit_batch(wid, r, H)
    finally:
        ctx.close()
    return out[0] if single else out


class FFILogLike(_OpBase):
    """ONE Op for the whole FFI seismic(+geodetic+laplacian) likelihood sub-graph; wraps ``BatchedFFILogLike``.

    ``perform`` takes the flat parameter vector q (or a [B, n_params] matrix) and returns ``logpts`` ([n_out] /
    [B, n_out]) and ``like`` -- the Deterministics the reference names ``seis_like`` / ``geo_like`` /
    ``laplacian_like`` and ``like`` (beat/models/seismic.py:1348, problems.py:246-247)."""

    __props__ = ("name",)
    n_outputs = 2

    def __init__(self, evaluator, name="ffi_loglike"):
        self.evaluator = evaluator
        self.name = name

    def make_node(self, q):
        q = _tt.as_tensor_variable(q)
        if getattr(q, "ndim", 1) == 2:                                       # [B, n_params]: logpts [B, n_out], like [B]
            return _Apply(self, [q], [_tt.dmatrix(), _tt.dvector()])
        return _Apply(self, [q], [_tt.dvector(), _tt.dscalar()])

    def perform(self, node, inputs, output):
        (q,) = inputs
        q = np.asarray(q, dtype=np.float64)
        logpts, like = self.evaluator(q)
        if q.ndim == 1:
            output[0][0], output[1][0] = logpts[0], np.asarray(like[0])
        else:
            output[0][0], output[1][0] = logpts, like

    def infer_shape(self, fgraph=None, node=None, input_shapes=None):
        if input_shapes and len(input_shapes[0]) == 2:
            return [(input_shapes[0][0], self.evaluator.n_out), (input_shapes[0][0],)]
        return [(self.evaluator.n_out,), ()]
