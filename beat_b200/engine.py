"""
Batched forward-model + log-likelihood evaluator: the drop-in for the reference's compiled
``logp_forw_func(q)`` (beat/sampler/base.py:598-615, called at beat/sampler/metropolis.py:349), evaluated for all
chains of an SMC / PT population in lock-step on one B200.

``BatchedFFILogLike`` owns one libbeatgpu context (one process per GPU), keeps every static operand resident in
HBM (GF libraries, data, weights, geometry) and exposes

  * ``__call__(Q)``           host numpy ``[B, n_params]`` -> ``(logpts [B, n_out], like [B])`` (copies in/out)
  * ``eval_pinned(...)``      same through caller-provided pinned buffers (the e2e benchmark path)
  * ``eval_device(q_dev)``    torch CUDA tensors in / out, nothing leaves the device (the sampler path)
  * ``update_weights(...)``   between SMC stages (beat/models/seismic.py:1509-1534)
  * ``logp_forw_func(q)``     single-chain call with the reference's return convention (list of arrays).

PyTorch is used only as plumbing for device memory / streams / torch.distributed; all arithmetic is in the
hand-written CUDA kernels behind the C-ABI.  There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import lib as _lib
from .lib import F32, F64, Context, Layout


def _dtype_code(store_dtype):
    if store_dtype in (F32, "float32", "f32", np.float32):
        return F32
    if store_dtype in (F64, "float64", "f64", np.float64):
        return F64
    raise ValueError("store_dtype must be float32 or float64")


class BatchedFFILogLike:
    def __init__(self, device=0):
        self.ctx = Context(device)
        self.device = device
        self.wmap_ids = []
        self._wm_shapes = []
        self.n_params = 0
        self.npatches = 0

    # ------------------------------------------------------------------ construction
    @classmethod
    def from_problem(cls, prob, device=0, store_dtype="float32", upload_libraries=True):
        """Upload a problem dict (see beat_b200.synthetic.make_problem for the schema)."""
        self = cls(device)
        ctx = self.ctx
        sfs = prob["subfaults"]
        ctx.set_fault([s[0] for s in sfs], [s[1] for s in sfs], [s[2] for s in sfs])
        off = prob["offsets"]
        L = Layout()
        L.n_params = prob["n_params"]
        L.n_slipvars = len(prob["slip_vars"])
        for i in range(_lib.MAX_SLIPVARS):
            L.off_slip[i] = off.get(prob["slip_vars"][i], -1) if i < L.n_slipvars else -1
        L.off_durations = off.get("durations", -1)
        L.off_velocities = off.get("velocities", -1)
        L.off_nucleation_strike = off.get("nucleation_strike", -1)
        L.off_nucleation_dip = off.get("nucleation_dip", -1)
        L.off_time = off.get("time", -1)
        L.off_hypers = off.get("hypers", -1)
        L.n_hypers = prob["n_hypers"]
        L.off_time_shifts = off.get("time_shifts", -1)
        L.n_time_shifts = prob.get("n_time_shifts", 0)
        ctx.set_layout(L, prob.get("fixed"))
        self.n_params = prob["n_params"]
        self.npatches = prob["npatches"]
        self.slip_vars = tuple(prob["slip_vars"])
        code = _dtype_code(store_dtype)
        self.store_dtype = code
        for wm in prob["wavemaps"]:
            wid = ctx.add_wavemap(wm["nt"], wm["ns"], wm["interpolation"], wm.get("station_idx"), wm["hyper_idx"], wm["nsamples"])
            self.wmap_ids.append(wid)
            self._wm_shapes.append((wm["nt"], wm["ns"]))
            if upload_libraries:
                for iv, v in enumerate(self.slip_vars):
                    ctx.upload_gflib(wid, iv, np.ascontiguousarray(wm["G"][v]), code, wm["dur_min"], wm["dur_step"],
                                     wm["st_min"], wm["st_step"])
            ctx.upload_data(wid, wm["data"])
            ctx.update_weights(wid, wm["U"], wm["slog_pdet"])
        if prob.get("geodetic"):
            g = prob["geodetic"]
            ctx.set_geodetic(g["slices"], [g["G"][v] for v in self.slip_vars], g["data"], g["odw"], g["U"],
                             g["slog_pdet"], g["nsamples"], g["hyper_idx"])
        if prob.get("laplacian"):
            lp = prob["laplacian"]
            ctx.set_laplacian(lp["L"], lp["sdet"], lp["hyper_idx"])
        self.n_out = ctx.n_outputs()
        esz = 4 if code == F32 else 8
        self._bytes_per_eval = sum(wm["nt"] * prob["npatches"] * wm["ns"] * (4 if wm["interpolation"] == "multilinear" else 1)
                                   * len(self.slip_vars) * esz for wm in prob["wavemaps"])
        return self

    def alloc_library(self, wmap_index, slipvar_index, dims, dur_min, dur_step, st_min, st_step):
        """Allocate a library in HBM and return (device_ptr, row_stride) for device-side filling (bench)."""
        return self.ctx.alloc_gflib(self.wmap_ids[wmap_index], slipvar_index, self.store_dtype, dims, dur_min, dur_step,
                                    st_min, st_step)

    def update_weights_device(self, wmap_index, U_dev, slog_pdet_dev):
        """Between SMC stages with the new weights still on the device: contiguous CUDA float64 torch tensors
        [nt, ns, ns] and [nt] (``covariance.weights_from_residuals_device``); no host round trip."""
        import torch
        for x in (U_dev, slog_pdet_dev):
            if x.dtype != torch.float64 or not x.is_cuda or not x.is_contiguous():
                raise ValueError("weights must be contiguous CUDA float64 tensors")
        torch.cuda.current_stream(U_dev.device).synchronize()          # the tensors were produced on torch's stream
        self.ctx.update_weights_dev(self.wmap_ids[wmap_index], U_dev.data_ptr(), slog_pdet_dev.data_ptr())

    def update_weights(self, wmap_index, U, slog_pdet):
        self.ctx.update_weights(self.wmap_ids[wmap_index], U, slog_pdet)

    # ------------------------------------------------------------------ evaluation
    def __call__(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        if Q.ndim == 1:
            Q = Q[None, :]
        if Q.shape[1] != self.n_params:
            raise ValueError("q has %d parameters, model expects %d" % (Q.shape[1], self.n_params))
        return self.ctx.ffi_loglike_batch(Q)

    def logp_forw_func(self, q):
        """Single chain, reference return convention: list ``[logpts..., like]`` of ndarrays."""
        logpts, like = self(np.asarray(q, dtype=np.float64)[None, :])
        return [logpts[0], like[0]]

    def eval_pinned(self, B, q_pinned_ptr, logpts_pinned_ptr, like_pinned_ptr):
        """End-to-end call on caller-owned (pinned) host buffers: H2D + kernels + D2H + sync."""
        self.ctx.ffi_loglike_batch_ptr(B, q_pinned_ptr, logpts_pinned_ptr, like_pinned_ptr)

    def eval_device(self, q_dev, logpts_out=None, like_out=None):
        """torch CUDA float64 tensor [B, n_params] -> (logpts [B, n_out], like [B]) torch tensors on the same device.

        Enqueued on torch's current stream; no host synchronisation."""
        import torch
        if q_dev.dtype != torch.float64 or not q_dev.is_cuda or not q_dev.is_contiguous():
            raise ValueError("q_dev must be a contiguous CUDA float64 tensor")
        B = q_dev.shape[0]
        if logpts_out is None:
            logpts_out = torch.empty((B, self.n_out), dtype=torch.float64, device=q_dev.device)
        if like_out is None:
            like_out = torch.empty((B,), dtype=torch.float64, device=q_dev.device)
        stream = torch.cuda.current_stream(q_dev.device).cuda_stream
        if getattr(self, "_bound_stream", -1) != stream:
            self.ctx.set_stream(stream, external=True)      # order our kernels with torch's work on its current stream
            self._bound_stream = stream
        self.ctx.ffi_loglike_batch_dev(B, q_dev.data_ptr(), logpts_out.data_ptr(), like_out.data_ptr())
        return logpts_out, like_out

    def drain_diagnostics(self):
        """Counters of the device-pointer path since the last call (read + reset; one host sync): library taps that fell
        outside the library (the chains concerned carry NaN logpts and are rejected by the sampler)."""
        return {"index_violations": self.ctx.index_violations()}

    def get_synthetics(self, Q, wmap_index=0):
        """Forward model only (reference: SeismicDistributerComposite.get_synthetics, outmode="array",
        beat/models/seismic.py:1351-1507): Q [B, n_params] or [n_params] -> synthetics [B, nt, ns] / [nt, ns]."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        single = Q.ndim == 1
        if single:
            Q = Q[None, :]
        wm = self._wm_shapes[wmap_index]
        out = self.ctx.ffi_synthetics_batch(self.wmap_ids[wmap_index], Q, wm[0], wm[1])
        return out[0] if single else out

    def stats(self, B=None):
        """Counters for logging (the reference only has wall-clock debug lines, beat/models/seismic.py:1229,1345-1346):
        kernels launched so far, duration of the last stacking pass and -- given the batch size -- the achieved
        algorithmic GB/s of that pass (SURVEY 8d bytes: nt*np*ns*K*nvar*sizeof(gf) per evaluation)."""
        out = {"launches": self.ctx.launch_count()}
        try:
            ms = self.ctx.last_stack_ms()
        except Exception:
            return out
        out["last_stack_ms"] = ms
        if B:
            out["evals_per_s_stack"] = B / (ms / 1e3)
            if getattr(self, "_bytes_per_eval", None):
                out["algorithmic_GBps"] = self._bytes_per_eval * B / (ms / 1e3) / 1e9
        return out

    def starttimes(self, B):
        return self.ctx.get_starttimes(B, self.npatches)

    def close(self):
        self.ctx.close()
