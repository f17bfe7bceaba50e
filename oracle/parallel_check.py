"""
oracle/parallel_check.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Fans the CPU checkers out over the host cores so that the `-m gpu` parity tests can compare the CUDA path with the
oracle at BASELINE.json's full sizes ("llk rtol 1e-5 on all B chains", section 7 step 2: "start-time indices
identical for >= 1e6 random chains") in seconds instead of hours.  Workers are SPAWNED, not forked: the parent test
process holds a CUDA context (H6) and the children must never inherit it; they only run numpy, the reference's own
compiled fast_sweep_ext (oracle/_ref) and oracle/libfsport.so.

Only tests/ may import this module.
"""
from __future__ import annotations

import multiprocessing as mp
import os
from concurrent.futures import ProcessPoolExecutor

import numpy as np


def n_workers(limit=32):
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    return max(1, min(n, limit))


def _init():
    for v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[v] = "1"
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass


def pool(workers=None):
    return ProcessPoolExecutor(max_workers=workers or n_workers(), mp_context=mp.get_context("spawn"), initializer=_init)


# ------------------------------------------------------------------------------------------------ fast sweeping
def sweep_chunk(task):
    """(slowness [n, np], patch_size, nuc_dip_idx [n], nuc_strike_idx [n], n_dip, n_strike) ->
    dict(ref = start times of the REFERENCE's compiled fast_sweep_ext [n, np] or None when oracle/_ref is absent,
         port = start times of the C restatement, iters = its outer iteration counts)."""
    from oracle import ffi_oracle as O
    slow, h, hr, hc, nd, ns = task
    ext = O.load_reference_ext()
    ref = None
    if ext is not None:
        ref = np.empty_like(slow)
        for i in range(slow.shape[0]):
            # argument order of Sweeper.perform (beat/pytensorf.py:475-482)
            ref[i] = ext.fast_sweep(np.ascontiguousarray(slow[i]), float(h), int(hr[i]), int(hc[i]), int(nd), int(ns))
    port, iters = O.fast_sweep_batch_port(slow, h, hr, hc, nd, ns)
    return dict(ref=ref, port=port, iters=iters)


# ------------------------------------------------------------------------------------------------ full-size llk
def _slice_target(wm, slip_vars, t):
    """The operands of target t of a wavemap (no libraries), as a one-target wavemap."""
    sidx = wm.get("station_idx")
    out = {k: wm[k] for k in ("ns", "ndur", "nst", "dur_min", "dur_step", "st_min", "st_step", "interpolation")}
    out.update(nt=1, A={v: wm["A"][v][t:t + 1] for v in slip_vars}, k0={v: wm["k0"][v][t:t + 1] for v in slip_vars},
               data=wm["data"][t:t + 1], U=wm["U"][t:t + 1], slog_pdet=wm["slog_pdet"][t:t + 1], nsamples=wm["nsamples"][t:t + 1],
               hyper_idx=wm["hyper_idx"][t:t + 1], station_idx=None if sidx is None else np.asarray(sidx)[t:t + 1])
    return out


def target_logpts(task):
    """One target of a full-size FFI problem, all chains: the library block of that target is regenerated here from the
    synthetic recipe (float64, beat_b200.synthetic.library_block -- the same recipe the device fill uses), then the
    oracle evaluates every chain against it.  task = (meta, wm1, Q): meta = the problem dict without wavemaps, wm1 = the
    one-target wavemap of `_slice_target`.  Returns logpts [B] of that target."""
    from beat_b200 import synthetic
    from oracle import ffi_oracle as O
    meta, wm1, Q = task
    wm1 = dict(wm1)
    wm1["G"] = {v: synthetic.library_block(wm1["A"][v], wm1["k0"][v], wm1["ndur"], wm1["nst"], wm1["ns"], wm1["st_step"], meta["dt"])
                for v in meta["slip_vars"]}
    sub = dict(meta, wavemaps=[wm1])
    out = np.empty(Q.shape[0])
    for c in range(Q.shape[0]):
        out[c] = O.ffi_seismic_eval(sub, synthetic.split_point(meta, Q[c]), impl="port")[0]
    return out


def full_size_logpts(prob, Q, executor, data=None, U=None, slog_pdet=None):
    """Oracle logpts [B, nt] of a one-wavemap problem built with build_library=False (libraries regenerated per target
    inside the workers).  data / U / slog_pdet optionally replace the wavemap's (near-MAP populations)."""
    wm = dict(prob["wavemaps"][0])
    if data is not None:
        wm["data"] = data
    if U is not None:
        wm["U"], wm["slog_pdet"] = U, slog_pdet
    meta = {k: v for k, v in prob.items() if k not in ("wavemaps", "geodetic", "laplacian")}
    tasks = [(meta, _slice_target(wm, prob["slip_vars"], t), Q) for t in range(wm["nt"])]
    return np.stack(list(executor.map(target_logpts, tasks)), axis=1)
