"""
Geometry-mode (point-source) seismic forward model + log-likelihood on the GPU: host-side mirror of the reference's
interface for BASELINE config 2, backed by libbeatgpu's ``beatgpu_geom_*`` entries.

  * ``ArrivalTaper``            <- beat/heart.py:266-336   (a, b, c, d; ``nsamples``, ``duration``, ``fadein``)
  * ``Filter`` / ``BandstopFilter`` <- beat/heart.py:342-412 (``apply`` becomes ``sections(deltat)``: the IIR cascade
                                   pyrocko's Trace.highpass/lowpass/bandpass/bandstop would run, designed with the same
                                   scipy.signal.butter call -- set-up time only)
  * ``SeisSynthesizer``         <- beat/pytensorf.py:129-311 (Op calling convention: dict of named source variables in,
                                   ``(synthetics [nt, ns], tmins [nt])`` out; a leading chain axis batches it)
  * ``BatchedGeometryLogLike``  <- the compiled ``logp_forw_func(q)`` of SeismicGeometryComposite.get_formula
                                   (beat/models/seismic.py:737-837) for all chains of a population

There is no CPU fallback: the synthesis runs in the CUDA kernels of csrc/geom.cuh or not at all.
"""
from __future__ import annotations

import numpy as np

from .lib import GEOM_VARS, Context, GeomLayout
from .ops import HAVE_PYTENSOR, _Apply, _OpBase, _tt


class ArrivalTaper(object):
    """Cosine arrival taper, times [s] w.r.t. the phase arrival (beat/heart.py:266-336)."""

    def __init__(self, a=-15.0, b=-10.0, c=50.0, d=55.0):
        self.a, self.b, self.c, self.d = float(a), float(b), float(c), float(d)
        if not self.a < self.b < self.c < self.d:
            raise ValueError("Taper values violate: a < b < c < d")          # heart.py:326-327

    def duration(self, chop_bounds=("b", "c")):
        return getattr(self, chop_bounds[1]) - getattr(self, chop_bounds[0])

    def nsamples(self, sample_rate, chop_bounds=("b", "c")):
        return int(np.ceil(sample_rate * self.duration(chop_bounds)))     # heart.py:294-304

    @property
    def fadein(self):
        return self.b - self.a

    @property
    def fadeout(self):
        return self.d - self.c

    def check_sample_rate_consistency(self, deltat):
        for cb in (("b", "c"), ("a", "d")):                                # heart.py:277-292
            ratio = self.duration(cb) / deltat
            if abs(ratio - round(ratio)) > 1e-9:
                raise ValueError("Taper duration %g of %s is inconsistent with sampling rate of %g! Please adjust Taper values!"
                                 % (self.duration(cb), ", ".join(cb), deltat))

    def abcd(self):
        return (self.a, self.b, self.c, self.d)


class Filter(object):
    """Time-domain band-pass (beat/heart.py:366-392): stepwise = high-pass (after removing the mean) then low-pass."""

    def __init__(self, lower_corner=0.001, upper_corner=0.1, order=4, stepwise=True):
        self.lower_corner, self.upper_corner, self.order, self.stepwise = lower_corner, upper_corner, int(order), bool(stepwise)

    def sections(self, deltat):
        from scipy import signal
        if self.stepwise:
            return [signal.butter(self.order, [self.lower_corner * 2.0 * deltat], btype="high") + (True,),
                    signal.butter(self.order, [self.upper_corner * 2.0 * deltat], btype="low") + (False,)]
        return [signal.butter(self.order, [c * 2.0 * deltat for c in (self.lower_corner, self.upper_corner)], btype="band") + (True,)]


class BandstopFilter(object):
    """beat/heart.py:395-412 (no demeaning)."""

    def __init__(self, lower_corner=0.12, upper_corner=0.25, order=4):
        self.lower_corner, self.upper_corner, self.order = lower_corner, upper_corner, int(order)

    def sections(self, deltat):
        from scipy import signal
        return [signal.butter(self.order, [c * 2.0 * deltat for c in (self.lower_corner, self.upper_corner)], btype="bandstop") + (False,)]


def filterer_from_dicts(filterer):
    """Problem-dict filter entries (kind, order, lower_corner, upper_corner) -> Filter objects."""
    out = []
    for f in filterer:
        if isinstance(f, (Filter, BandstopFilter)):
            out.append(f)
        elif f["kind"] in ("stepwise", "bandpass"):
            out.append(Filter(f["lower_corner"], f["upper_corner"], f["order"], stepwise=f["kind"] == "stepwise"))
        elif f["kind"] == "bandstop":
            out.append(BandstopFilter(f["lower_corner"], f["upper_corner"], f["order"]))
        else:
            raise ValueError("unknown filter kind %r" % (f["kind"],))
    return out


def _sections(filterer, deltat):
    secs = []
    for f in filterer_from_dicts(filterer):
        secs.extend(f.sections(deltat))
    return secs


def _layout_from_offsets(offsets, n_params, n_hypers, n_time_shifts=0, n_sources=1):
    L = GeomLayout()
    L.n_params = n_params
    for v in GEOM_VARS:
        setattr(L, "off_" + v, offsets.get(v, -1))
    L.off_hypers = offsets.get("hypers", -1)
    L.n_hypers = n_hypers
    L.off_time_shifts = offsets.get("time_shifts", -1)
    L.n_time_shifts = n_time_shifts
    L.n_sources = n_sources
    return L


class BatchedGeometryLogLike:
    """Batched ``logp_forw_func`` of a geometry-mode seismic problem (see module docstring)."""

    def __init__(self, device=0):
        self.ctx = Context(device)
        self.device = device
        self.wmap_ids = []
        self._wm_shapes = []
        self.n_params = 0

    @classmethod
    def from_problem(cls, gprob, device=0, upload_data=True):
        """Upload a problem dict (schema: beat_b200.synthetic.make_geometry_problem)."""
        self = cls(device)
        ctx = self.ctx
        st = gprob["store"]
        fixed = gprob.get("fixed")
        ctx.geom_set_source(_layout_from_offsets(gprob["offsets"], gprob["n_params"], gprob["n_hypers"],
                                                 gprob.get("n_time_shifts", 0), gprob.get("n_sources", 1)), fixed,
                            gprob["event"]["lat"], gprob["event"]["lon"], gprob.get("stf_anchor", -1.0))
        if gprob.get("stf_type", "HalfSinusoid") != "HalfSinusoid":
            ctx.geom_set_stf(gprob["stf_type"], gprob.get("stf_anchor", -1.0), gprob["offsets"].get("peak_ratio", -1),
                             gprob.get("peak_ratio", 0.5))
        self.store_id = ctx.geom_upload_store(st["traces"], st["itmin"], st["nsamples"], st["z0"], st["dz"], st["x0"], st["dx"],
                                              st["deltat"])
        self.n_params = gprob["n_params"]
        self._bytes_per_eval = 0
        for wm in gprob["wavemaps"]:
            taper = wm["taper"] if not isinstance(wm["taper"], ArrivalTaper) else wm["taper"].abcd()
            ArrivalTaper(*taper).check_sample_rate_consistency(st["deltat"])
            wid = ctx.geom_add_wavemap(self.store_id, wm["ns"], wm["interpolation"], wm["lats"], wm["lons"], wm["azimuths"],
                                       wm["dips"], wm["arrival_times"], taper, wm.get("chop_bounds", ("b", "c")),
                                       _sections(wm["filterer"], st["deltat"]), wm["hyper_idx"], wm["nsamples"],
                                       station_idx=wm.get("station_idx"))
            self.wmap_ids.append(wid)
            self._wm_shapes.append((wm["nt"], wm["ns"]))
            if upload_data and wm.get("data") is not None:
                ctx.upload_data(wid, wm["data"])
                ctx.update_weights(wid, np.ascontiguousarray(wm["U"]), wm["slog_pdet"])
            # algorithmic bytes per evaluation: receivers x nodes x 10 components x window x 4 B
            a, b, c, d = taper
            for at in {(la, lo, at) for la, lo, at in zip(wm["lats"], wm["lons"], wm["arrival_times"])}:
                n_raw = int(np.ceil((at[2] + d + 2 * (b - a)) / st["deltat"]) - np.floor((at[2] + a - 2 * (b - a)) / st["deltat"])) + 1
                self._bytes_per_eval += gprob.get("n_sources", 1) * (4 if wm["interpolation"] == "multilinear" else 1) * 10 * n_raw * 4
        self.n_out = ctx.n_outputs()
        return self

    def upload_data(self, wmap_index, data, U, slog_pdet):
        wid = self.wmap_ids[wmap_index]
        self.ctx.upload_data(wid, data)
        self.ctx.update_weights(wid, np.ascontiguousarray(U), slog_pdet)

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
        self.ctx.update_weights(self.wmap_ids[wmap_index], np.ascontiguousarray(U), slog_pdet)

    def __call__(self, Q):
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        if Q.ndim == 1:
            Q = Q[None, :]
        if Q.shape[1] != self.n_params:
            raise ValueError("q has %d parameters, model expects %d" % (Q.shape[1], self.n_params))
        return self.ctx.geom_loglike_batch(Q)

    def logp_forw_func(self, q):
        logpts, like = self(np.asarray(q, dtype=np.float64)[None, :])
        return [logpts[0], like[0]]

    def eval_pinned(self, B, q_pinned_ptr, logpts_pinned_ptr, like_pinned_ptr):
        self.ctx.geom_loglike_batch_ptr(B, q_pinned_ptr, logpts_pinned_ptr, like_pinned_ptr)

    def eval_device(self, q_dev, logpts_out=None, like_out=None):
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
            self.ctx.set_stream(stream, external=True)
            self._bound_stream = stream
        self.ctx.geom_loglike_batch_dev(B, q_dev.data_ptr(), logpts_out.data_ptr(), like_out.data_ptr())
        return logpts_out, like_out

    def drain_diagnostics(self):
        """Counters of the device-pointer path since the last call (read + reset; one host sync): proposals whose source
        left the GF store (their logpts are NaN, i.e. rejected) and GF-store bulk copies that timed out (must be 0)."""
        return {"index_violations": self.ctx.index_violations(), "geom_timeouts": self.ctx.geom_timeouts()}

    def get_synthetics(self, Q, wmap_index=0):
        """heart.seis_synthetics(..., outmode="array") for every chain: [B, nt, ns] (or [nt, ns] for one point)."""
        Q = np.ascontiguousarray(Q, dtype=np.float64)
        single = Q.ndim == 1
        if single:
            Q = Q[None, :]
        nt, ns = self._wm_shapes[wmap_index]
        out = self.ctx.geom_synthetics_batch(self.wmap_ids[wmap_index], Q, nt, ns)
        return out[0] if single else out

    def close(self):
        self.ctx.close()


class SeisSynthesizer(_OpBase):
    """Mirror of the reference Op (beat/pytensorf.py:129-311; a ``pytensor.tensor.Op`` subclass when pytensor is
    importable, with the reference's dict-taking ``make_node``) for ONE double-couple source: called with a dict of named
    source variables (values scalar / [1] as the reference, or [B] / [B, 1] for a batch of chains) it returns
    ``(synthetics, tmins)`` -- [nt, ns] and [nt] (with a leading chain axis for a batch).  ``tmins`` are the taper's
    lower chop bound per target (beat/heart.py:3729), constant because the arrival times are fixed."""

    __props__ = ("store", "event", "targets", "arrival_taper", "arrival_times", "filterer", "pre_stack_cut",
                 "station_corrections")

    def __init__(self, store, event, targets, arrival_taper, arrival_times, filterer, pre_stack_cut=True,
                 station_corrections=False, interpolation="multilinear", stf_anchor=-1.0, chop_bounds=("b", "c"), device=0,
                 stf_type="HalfSinusoid"):
        if not pre_stack_cut:
            raise NotImplementedError("only pre_stack_cut=True (the reference's default) is implemented")
        self.store, self.event, self.targets = store, event, targets
        self.arrival_taper = arrival_taper if isinstance(arrival_taper, ArrivalTaper) else ArrivalTaper(*arrival_taper)
        self.arrival_times = np.asarray(arrival_times, dtype=np.float64)
        self.filterer = filterer_from_dicts(filterer)
        self.pre_stack_cut = pre_stack_cut
        self.chop_bounds = tuple(chop_bounds)
        self.nt = len(targets["lats"])
        self.ns = self.arrival_taper.nsamples(1.0 / store["deltat"], self.chop_bounds)
        self.varnames = list(GEOM_VARS)
        self.station_corrections = bool(station_corrections)
        offsets = {v: i for i, v in enumerate(GEOM_VARS)}
        n_par, n_ts = len(GEOM_VARS), 0
        if self.station_corrections:            # the Op receives one time_shift per target (pytensorf.py:248-252)
            offsets["time_shifts"], n_ts = n_par, self.nt
            n_par += n_ts
        self.stf_type = stf_type
        if stf_type == "Triangular":            # the triangle's `peak_ratio` is one more (optional) named input; 0.5 when absent
            offsets["peak_ratio"] = n_par
            n_par += 1
        self._off_peak = offsets.get("peak_ratio", -1)
        self._n_par = n_par
        self._ctx = Context(device)
        self._ctx.geom_set_source(_layout_from_offsets(offsets, n_par, 0, n_ts), None, event["lat"], event["lon"], stf_anchor)
        if stf_type != "HalfSinusoid":
            self._ctx.geom_set_stf(stf_type, stf_anchor, self._off_peak)
        sid = self._ctx.geom_upload_store(store["traces"], store["itmin"], store["nsamples"], store["z0"], store["dz"],
                                          store["x0"], store["dx"], store["deltat"])
        self._wid = self._ctx.geom_add_wavemap(sid, self.ns, interpolation, targets["lats"], targets["lons"], targets["azimuths"],
                                               targets["dips"], self.arrival_times, self.arrival_taper.abcd(), self.chop_bounds,
                                               _sections(self.filterer, store["deltat"]), np.zeros(self.nt, np.int32),
                                               np.full(self.nt, self.ns, np.int32),
                                               station_idx=np.arange(self.nt, dtype=np.int32) if self.station_corrections else None)

    def infer_shape(self, fgraph=None, node=None, input_shapes=None):
        return [(self.nt, self.ns), (self.nt,)]                           # pytensorf.py:303-311

    def perform(self, node, inputs, output):
        point = {v: np.asarray(i, dtype=np.float64) for v, i in zip(self.varnames, inputs)}
        shifts = point.pop("time_shift", None)
        peak = point.pop("peak_ratio", None)
        point = {v: point[v] for v in GEOM_VARS}                     # input order does not matter (pytensorf.py:133)
        B = max(p.size for p in point.values())
        batched = any(p.ndim >= 1 and p.size > 1 for p in point.values())
        Q = np.zeros((B, self._n_par))
        off_peak = getattr(self, "_off_peak", -1)
        if off_peak >= 0:
            Q[:, off_peak] = 0.5 if peak is None else peak.reshape(-1)
        for i, v in enumerate(GEOM_VARS):
            Q[:, i] = point[v].reshape(-1)
        arrival = np.broadcast_to(self.arrival_times, (B, self.nt))
        if self.station_corrections:
            if shifts is None:
                raise KeyError("time_shift")
            Q[:, len(GEOM_VARS):] = shifts.reshape(-1, self.nt)
            arrival = arrival + Q[:, len(GEOM_VARS):]
        synths = self._ctx.geom_synthetics_batch(self._wid, Q, self.nt, self.ns)
        tmins = arrival + getattr(self.arrival_taper, self.chop_bounds[0])       # heart.py:3729
        output[0][0] = synths if batched else synths[0]
        output[1][0] = tmins.copy() if batched else tmins[0].copy()

    def make_node(self, inputs):
        """``inputs``: dict of named tensors, exactly like the reference (pytensorf.py:215-239)."""
        self.varnames = list(inputs.keys())
        inlist = [_tt.as_tensor_variable(i) for i in inputs.values()]
        outm_shape, outv_shape = self.infer_shape()
        outm = _tt.as_tensor_variable(np.zeros(outm_shape))
        outv = _tt.as_tensor_variable(np.zeros(outv_shape))
        return _Apply(self, inlist, [outm.type(), outv.type()])

    if not HAVE_PYTENSOR:
        def __call__(self, inputs):
            """Eager call without pytensor: ``inputs`` is the dict of named variables ``make_node`` would take."""
            missing = [v for v in GEOM_VARS if v not in inputs]
            if missing:
                raise KeyError("source variables missing: %s" % ", ".join(missing))
            if self.station_corrections and "time_shift" not in inputs:
                raise KeyError("time_shift")
            self.varnames = list(inputs.keys())
            out = [[None], [None]]
            self.perform(None, [inputs[v] for v in self.varnames], out)
            return out[0][0], out[1][0]

    def close(self):
        self._ctx.close()
