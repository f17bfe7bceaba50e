#!/usr/bin/env python
"""
bench.py -- forward+loglike evals/s on the FFI seismic 200-patch configuration (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A "step" is one lock-step evaluation of the hot path (fast sweep -> GF stack -> residual -> misfit -> like)
for every chain held by a rank: C3 = 1 subfault 10x20 patches (2 km), 64 targets x 120 samples, library
17 durations x 64 start times, slip components uparr+uperp, multilinear interpolation, exponential (Toeplitz)
data covariance; 4000 chains per GPU (weak scaling: chains are independent, sharded across ranks, one NCCL
all-gather of the per-chain llk per step for N > 1).  Synthetic GF libraries are generated directly in HBM.

`value`  : chains/s with q already resident in HBM (device-pointer entry), CUDA events, max over ranks.
`e2e`    : the same through the host-pointer C-ABI entry: pinned q -> H2D -> kernels -> D2H of logpts+like.
`roofline`: GF-stacking kernel, algorithmic bytes / its CUDA-event duration (event pairs recorded by the library
inside the timed loop) vs measured HBM copy peak; `traffic` only from an ncu capture of this very kernel source + blocking.
`strict_f64`: the same workload with the library stored and every operation done in f64 (the reference's precision).
`strong` (N > 1): the configuration as BASELINE.json names it -- n_chains = 4000 partitioned over the ranks.
`sampler_step`: the evaluation inside the lock-step Metropolis step (CUDA graph), also with every record written to trace files.
`cpu_baseline` / `--impl reference`: the oracle's restatement of the reference CPU path (reference's own compiled
fast_sweep_ext when oracle/_ref exists, numpy stack_all mode, numpy llk), one chain at a time, fanned out over the
host cores with a fork pool the way the reference's paripool does, on a host library with all 17 duration nodes when the
box has the memory -- a reported baseline, not the target.
`--config c4 | c5 | c2 | c2llk | c3big`, `--interpolation nearest_neighbor`, `--store f64`, `--noise dense`: the other
configurations of BASELINE.json / SURVEY 8d for the record (never reported under the C3 metric name).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

_T0 = time.perf_counter()


def log(msg):
    print("[bench %7.1fs] %s" % (time.perf_counter() - _T0, msg), file=sys.stderr, flush=True)


def usable_cores():
    """Host threads this process may use, bounded by memory (each worker peaks at ~0.6 GB of numpy temporaries)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    try:
        import psutil
        n = min(n, max(1, int(psutil.virtual_memory().available / 2**30 / 1.5)))
    except Exception:
        pass
    return max(1, min(n, int(os.environ.get("BENCH_MAX_CORES", "256"))))


def _worker_init():
    """One BLAS/OpenMP thread per worker: the pool already uses every core (128 workers x 128 BLAS threads spin-wait
    each other to a standstill otherwise)."""
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass


def pool_map(fn, items, cores, timeout):
    """Fork pool that raises instead of hanging when a worker dies (reference: beat/parallel.py:186-282)."""
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("fork"), initializer=_worker_init) as ex:
        list(ex.map(fn, range(cores), timeout=timeout))       # warm the workers
        t0 = time.perf_counter()
        list(ex.map(fn, items, chunksize=max(1, len(items) // (4 * cores)), timeout=timeout))
        return time.perf_counter() - t0


METRIC = "forward+loglike evals/sec (FFI seismic 200-patch)"
UNIT = "evals/s"


NOISE = "exponential"


def metric_name():
    if NOISE != "exponential":
        return "forward+loglike evals/sec (%s, %s covariance; not the BASELINE metric)" % (CONFIG.upper(), NOISE)
    return METRIC if CONFIG == "c3" else "forward+loglike evals/sec (%s, not the BASELINE metric)" % CONFIG.upper()


CONFIG = "c3"


def c3_args(quick):
    """Problem shapes.  c3 is the BASELINE.json metric (headline); c4 / c5 are the other FFI configs of BASELINE.json,
    selectable with --config for the record (profiles/README.md), never reported under the c3 metric name."""
    if quick:
        return dict(nt=8, subfaults=((5, 8, 2.0),), ns=40, ndur=5, nst=40)
    if CONFIG == "c4":      # joint geodetic + seismic FFI, dense non-Toeplitz geodetic covariance, laplacian prior
        return dict(nt=64, subfaults=((10, 20, 2.0),), ns=120, ndur=17, nst=64, geodetic=dict(nobs=[500]), laplacian=True)
    if CONFIG == "c3big":   # C3 with a 4x larger per-patch library block (128 start times x 240 samples): exercises the L2-derived chunk
        return dict(nt=16, subfaults=((10, 20, 2.0),), ns=240, ndur=17, nst=128)
    if CONFIG == "c5":      # multi-fault FFI, 2 segments x 150 patches
        return dict(nt=64, subfaults=((10, 15, 2.0), (10, 15, 2.0)), ns=120, ndur=17, nst=64)
    return dict(nt=64, subfaults=((10, 20, 2.0),), ns=120, ndur=17, nst=64)


def workload_config(args, n_gpus, chains):
    a = c3_args(args.quick)
    nd, nstr, h = a["subfaults"][0]
    npatch = sum(x[0] * x[1] for x in a["subfaults"])
    extra = {"c3": "", "c3big": " (per-patch library block 4x that of C3)",
             "c4": " + geodetic static (500 obs, dense non-Toeplitz C) + laplacian prior",
             "c5": " (2 subfaults x %d patches)" % (nd * nstr)}[CONFIG if not args.quick else "c3"]
    return {
        "workload": "%s FFI seismic: %d patches (%dx%d, %.1f km) x %d targets x %d samples, library %d durations x %d "
                    "starttimes, uparr+uperp, %s, exponential covariance%s" % (CONFIG.upper(), npatch, nd, nstr, h, a["nt"], a["ns"],
                                                                            a["ndur"], a["nst"], args.interpolation, extra),
        "chains_per_gpu": chains, "global_chains": chains * n_gpus, "parallelism": "chains sharded x%d" % n_gpus,
        "gf_storage": args.store, "interpolation": args.interpolation,
        "accumulate": ("f32 FMA per patch pair, f64 running sum" if args.store == "f32" else "f64 FMA") + "; sweep, residual, misfit in f64",
        "scaling_protocol": "weak: chains_per_gpu fixed, one all-gather of llk per step; the `strong` block of the line holds n_chains = %d partitioned over the ranks" % chains,
        "l2": "inputs>L2 (GF library %.1f GB per GPU; q rotates between steps)" % (lib_bytes(a, args.store) / 1e9),
    }


def lib_bytes(a, store):
    npatch = sum(x[0] * x[1] for x in a["subfaults"])
    return 2 * a["nt"] * npatch * a["ndur"] * a["nst"] * a["ns"] * (4 if store == "f32" else 8)


def algorithmic_bytes_per_eval(a, store, interpolation):
    npatch = sum(x[0] * x[1] for x in a["subfaults"])
    K = 4 if interpolation == "multilinear" else 1
    return a["nt"] * npatch * a["ns"] * K * 2 * (4 if store == "f32" else 8)


# --------------------------------------------------------------------------------------------- CPU baseline
_CPU = {}


def _cpu_eval(i):
    from oracle import ffi_oracle
    from beat_b200 import synthetic
    q = _CPU["Q"][i]
    return ffi_oracle.ffi_seismic_eval(_CPU["prob"], synthetic.split_point(_CPU["prob"], q), impl=_CPU["impl"]).sum()


def host_memory_available():
    """Bytes this process may still allocate: the smaller of the machine's available memory and the cgroup's head-room."""
    avail = None
    try:
        import psutil
        avail = int(psutil.virtual_memory().available)
    except Exception:
        pass
    try:
        lim = open("/sys/fs/cgroup/memory.max").read().strip()
        if lim != "max":
            room = int(lim) - int(open("/sys/fs/cgroup/memory.current").read().strip())
            avail = room if avail is None else min(avail, room)
    except Exception:
        pass
    return avail


def _fill_target(job):
    """Fork-pool task: library values of one (slip variable, target) written into the inherited shared mapping."""
    from beat_b200 import synthetic
    v, t = job
    wm, a = _CPU["fill_wm"], _CPU["fill_args"]
    wm["G"][v][t] = synthetic.library_block(wm["A"][v][t:t + 1], wm["k0"][v][t:t + 1], a["ndur"], a["nst"], a["ns"], wm["st_step"],
                                            _CPU["fill_dt"])[0]
    return 0


def build_cpu_problem(args):
    """The host-side twin of the GPU workload.  Same shapes and the SAME library axes (17 duration nodes, 26.7 GB of f64
    at C3, as the reference holds it: mmap'd f64, ffi/base.py:171-176) when the box has the memory for it -- the library
    lives in one anonymous shared mapping that the fork pool fills in parallel and the evaluation workers inherit.  On a
    box without that head-room (or with BENCH_CPU_NDUR set) the host library keeps fewer duration nodes; the bytes gathered
    per evaluation are the same either way (nt*np*K*nvar rows of ns samples) and the `sample` text says which it was."""
    import mmap
    from beat_b200 import synthetic
    _worker_init()                          # one BLAS thread in this process, too: a threaded BLAS call after a fork pool can dead-lock
    a = c3_args(args.quick)
    ndur_full = a["ndur"]
    need = lib_bytes(a, "f64")
    avail = host_memory_available()
    cores = usable_cores()
    if os.environ.get("BENCH_CPU_NDUR"):
        ndur_host = max(2, min(ndur_full, int(os.environ["BENCH_CPU_NDUR"])))
    else:                                   # library + ~1.5 GB of numpy temporaries per worker + slack
        ndur_host = ndur_full if (avail is not None and avail > need + (cores * 2 + 24) * 2**30) else 2
    if args.quick or ndur_host < ndur_full:
        a = dict(a, ndur=ndur_host)
        prob = synthetic.make_problem(interpolation=args.interpolation, seed=1234, **a)
        prob["host_library"] = ("all %d duration nodes" % ndur_full) if ndur_host == ndur_full else "%d of the %d duration nodes (%s)" % (
            ndur_host, ndur_full, "BENCH_CPU_NDUR" if os.environ.get("BENCH_CPU_NDUR") else
            "bounded RAM: %s GB available, %.0f GB library" % ("?" if avail is None else "%.0f" % (avail / 2**30), need / 2**30))
        return prob
    t0 = time.perf_counter()
    prob = synthetic.make_problem(interpolation=args.interpolation, seed=1234, build_library=False, **a)
    npatch = prob["npatches"]
    for wm in prob["wavemaps"]:
        shape = (a["nt"], npatch, a["ndur"], a["nst"], a["ns"])
        for v in prob["slip_vars"]:
            mm = mmap.mmap(-1, int(np.prod(shape)) * 8)                 # MAP_SHARED | MAP_ANONYMOUS: inherited across fork
            wm["G"][v] = np.frombuffer(mm, dtype=np.float64).reshape(shape)
        _CPU.update(fill_wm=wm, fill_args=a, fill_dt=prob["dt"])
        jobs = [(v, t) for v in prob["slip_vars"] for t in range(a["nt"])]
        import multiprocessing as mp
        from concurrent.futures import ProcessPoolExecutor
        with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("fork"), initializer=_worker_init) as ex:
            list(ex.map(_fill_target, jobs, timeout=600))
    prob["host_library"] = "all %d duration nodes, %.1f GB f64 in shared host memory" % (a["ndur"], need / 1e9)
    log("cpu baseline: %.1f GB host library filled by %d workers in %.1f s" % (need / 1e9, cores, time.perf_counter() - t0))
    return prob


def cpu_baseline(args, n_evals_per_core=4, cores=None):
    """Times the reference-style CPU path; returns (dict for the JSON line, evals/s)."""
    from beat_b200 import synthetic
    from oracle import ffi_oracle
    subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    cores = cores or usable_cores()
    _worker_init()
    log("cpu baseline: building host problem (%d cores)" % cores)
    prob = build_cpu_problem(args)
    n = max(cores, n_evals_per_core * cores)
    Q = synthetic.draw_chains(prob, n, seed=999)
    have_ref = ffi_oracle.load_reference_ext() is not None
    _CPU.update(prob=prob, Q=Q, impl="ref" if have_ref else "port")
    log("cpu baseline: 1-core timing")
    # single core
    t0 = time.perf_counter()
    n1 = max(2, min(8, n))
    for i in range(n1):
        _cpu_eval(i)
    one = n1 / (time.perf_counter() - t0)
    log("cpu baseline: 1 core %.2f evals/s" % one)
    # chains fanned out over a fork pool (reference: beat/parallel.py:186-282).  The path is bound by memory traffic and
    # page faults of numpy temporaries, so more workers are not always faster: report the best of {all, 1/2, 1/4} cores
    allc, best_cores = 0.0, cores
    for nc in sorted({cores, max(1, cores // 2), max(1, cores // 4)}, reverse=True):
        nn = max(nc, n_evals_per_core * nc)
        log("cpu baseline: fanning %d chains over %d cores" % (nn, nc))
        rate = nn / pool_map(_cpu_eval, list(range(nn)), nc, timeout=240)
        log("cpu baseline: %d cores %.2f evals/s" % (nc, rate))
        if rate > allc:
            allc, best_cores, n = rate, nc, nn
    cores = best_cores
    info = {"value": allc, "unit": UNIT, "cores": cores, "kind": "port",
            "value_1core": one,
            "sample": "%d chains of the same C3 shapes (f64 host library, %s), one chain per call: "
                      "%s fast_sweep + numpy stack_all (%s) + numpy mvn-chol llk, fork pool over %d cores"
                      % (n, prob["host_library"], "reference's compiled" if have_ref else "C-restated", args.interpolation, cores)}
    _CPU.clear()                         # drops the shared host library (26.7 GB at C3) before the GPU arm starts
    del prob
    import gc
    gc.collect()
    return info, allc


def run_reference_arm(args):
    """`--impl reference`: the reference's CPU path (see module docstring) on this box's host cores, K timed steps of a
    bounded sample each.  The worker count is calibrated first (all / half / quarter of the cores): the path is bound by
    memory traffic and page faults of its numpy temporaries and does not scale to all hardware threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    for v in ("OPENBLAS_NUM_THREADS", "OMP_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ.setdefault(v, "1")
    import multiprocessing as mp
    from concurrent.futures import ProcessPoolExecutor
    from beat_b200 import synthetic
    from oracle import ffi_oracle
    subprocess.call(["make", "-s", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    all_cores = usable_cores()
    prob = build_cpu_problem(args)
    cands = sorted({all_cores, max(1, all_cores // 2), max(1, all_cores // 4)}, reverse=True)
    n_max = all_cores * 2
    Q = synthetic.draw_chains(prob, n_max * (args.steps + args.warmup + len(cands)), seed=999)
    have_ref = ffi_oracle.load_reference_ext() is not None
    _CPU.update(prob=prob, Q=Q, impl="ref" if have_ref else "port")
    best = (0.0, all_cores)
    k = 0
    for nc in cands:                                   # calibration: one step per candidate worker count
        with ProcessPoolExecutor(max_workers=nc, mp_context=mp.get_context("fork"), initializer=_worker_init) as pool:
            list(pool.map(_cpu_eval, range(nc), timeout=300))
            t0 = time.perf_counter()
            list(pool.map(_cpu_eval, range(k, k + 2 * nc), timeout=300))
            rate = 2 * nc / (time.perf_counter() - t0)
        log("reference arm: %d workers -> %.2f evals/s" % (nc, rate))
        k += 2 * nc
        if rate > best[0]:
            best = (rate, nc)
    cores = best[1]
    per_step = cores * 2
    with ProcessPoolExecutor(max_workers=cores, mp_context=mp.get_context("fork"), initializer=_worker_init) as pool:
        for _ in range(args.warmup):
            list(pool.map(_cpu_eval, range(k, k + per_step), timeout=300))
            k += per_step
        t0 = time.perf_counter()
        for _ in range(args.steps):
            list(pool.map(_cpu_eval, range(k, k + per_step), timeout=300))
            k += per_step
        dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        # the GPU arm's own config (same workload); what the CPU actually ran per step is in cpu_baseline.sample
        "config": workload_config(args, args.gpus, args.chains),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "%d chains per step (bounded sample of the %d-chain workload), same C3 shapes, f64 host "
                                   "library with %s; %s fast_sweep + numpy stack_all + numpy llk, one "
                                   "chain per call, fork pool over %d of %d cores (best of %s)"
                                   % (per_step, args.chains, prob["host_library"], "reference's compiled" if have_ref else "C-restated", cores,
                                      all_cores, cands)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  NVML (pynvml) is polled every few milliseconds
    from a thread that is already running when the region starts -- a freshly spawned `nvidia-smi -lms` often delivers its
    first line only after a 120 ms region has ended; it remains the fall-back when pynvml is missing."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu_index):
        self.rows, self.proc, self.nvml, self._stop = [], None, None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.handle = None
            try:        # CUDA_VISIBLE_DEVICES may renumber devices: resolve the NVML handle through the device's UUID
                import torch
                uuid = str(torch.cuda.get_device_properties(gpu_index).uuid)
                uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
                try:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid)
                except TypeError:
                    self.handle = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
            except Exception:
                self.handle = None
            if self.handle is None:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            try:
                self.power_limit = pynvml.nvmlDeviceGetEnforcedPowerLimit(self.handle) / 1000.0
            except Exception:
                self.power_limit = None
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        while not self._stop:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                try:
                    pw = nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0
                except Exception:
                    pw = None
                self.rows.append((time.perf_counter(), sm, mask, pw))
            except Exception:
                pass
            time.sleep(0.004)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if self.nvml:
            self._stop = True
            self.thread.join(timeout=1.0)
            inside = [r for r in self.rows if t0 <= r[0] <= t1]
            reasons = sorted({name for r in inside for bit, name in self.NVML_REASONS if r[2] & bit})
            pw = [r[3] for r in inside if r[3] is not None]
            return {"sm_mhz": float(np.median([r[1] for r in inside])) if inside else None, "sm_max_mhz": self.max_sm,
                    "reasons": reasons, "samples": len(inside), "source": "nvml",
                    # NVML's power figure is a ~1 s moving average, so it lags a 0.1 s region; reported for context
                    "power_w": float(np.median(pw)) if pw else None, "power_limit_w": self.power_limit}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if ts < t0 or ts > t1 + 0.1:
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[0]))
                mx = float(p[1])
            except Exception:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi"}


# --------------------------------------------------------------------------------------------- GPU arm
def bind_to_gpu_numa(torch, device):
    """N > 1: pin this rank to the CPU cores NVML names as local to its GPU BEFORE the pinned host buffers are allocated,
    so that they are first touched on the GPU's NUMA node (eight ranks copying q from one node share its memory
    controller and the inter-socket link).  Returns a short description for the JSON line; never fatal."""
    try:
        import pynvml
        pynvml.nvmlInit()
        uuid = str(torch.cuda.get_device_properties(device).uuid)
        uuid = uuid if uuid.startswith("GPU-") else "GPU-" + uuid
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid)
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus or len(cpus) == len(os.sched_getaffinity(0)):
            return "not bound (NVML affinity = the whole process mask)"
        os.sched_setaffinity(0, cpus)
        return "bound to %d GPU-local cores (NVML)" % len(cpus)
    except Exception as e:
        return "not bound (%s)" % type(e).__name__


def run_gpu_arm(args):
    # stdout carries exactly ONE line (the JSON): native libraries (NCCL prints its version banner with printf) write
    # to file descriptor 1 directly, so fd 1 points at stderr until the line is ready
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    try:
        line = _run_gpu_arm(args)
    finally:
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        os.close(saved_stdout)
    if line is not None:
        print(json.dumps(line), flush=True)


DTYPE_LABEL = {"f32": "f32 library storage, f32 FMA per patch pair, f64 running sum; sweep, residual, misfit in f64",
               "f64": "f64 (library storage and every operation)"}


def touched_library_bytes(prob, a, store, interpolation, Q, t0):
    """Distinct library bytes the chains `Q` (rupture onset times `t0` [B, np] from the device) gather in one step:
    the HBM traffic an L2-perfect gather needs.  Same cells for every target (no station corrections in this workload)."""
    wm = prob["wavemaps"][0]
    off = prob["offsets"]
    npatch = prob["npatches"]
    dur = Q[:, off["durations"]:off["durations"] + npatch]
    x = (dur - wm["dur_min"]) / wm["dur_step"]
    y = (t0 - wm["st_min"]) / wm["st_step"]
    nst = wm["nst"]
    total_rows = 0
    for p in range(npatch):
        if interpolation == "multilinear":
            dc, sc = np.ceil(x[:, p]).astype(np.int64), np.ceil(y[:, p]).astype(np.int64)
            dfl = np.where(dc - x[:, p] == 0.0, dc, dc - 1)
            sfl = np.where(sc - y[:, p] == 0.0, sc, sc - 1)
            rows = np.concatenate([dc * nst + sc, dc * nst + sfl, dfl * nst + sc, dfl * nst + sfl])
        else:
            rows = np.rint(x[:, p]).astype(np.int64) * nst + np.rint(y[:, p]).astype(np.int64)
        total_rows += np.unique(rows).size
    row_bytes = a["ns"] * (4 if store == "f32" else 8)
    return int(total_rows) * a["nt"] * len(prob["slip_vars"]) * row_bytes


def _run_gpu_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        args.gpus = world
    n_gpus = max(1, world)

    # CPU baseline first (fork pool before CUDA is initialised in this process), rank 0 at N=1 only
    cpu_info = None
    if rank == 0 and n_gpus == 1 and not args.no_cpu_baseline:
        try:
            cpu_info, _ = cpu_baseline(args)
        except Exception as e:                     # a reported baseline must never take the GPU measurement down
            cpu_info = {"value": None, "unit": UNIT, "cores": usable_cores(), "kind": "port", "sample": "failed: %r" % (e,)}
            log("cpu baseline failed: %r" % (e,))

    import torch
    import torch.distributed as dist
    from beat_b200 import lib as beatlib
    from beat_b200 import synthetic
    from beat_b200.devlib import fill_library_on_device
    from beat_b200.engine import BatchedFFILogLike

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the GPU arm has no CPU fallback (use --impl reference for the CPU path)")
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    if n_gpus > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG", "WARN")        # keep stdout to the one JSON line (no "NCCL version" banner)
        dist.init_process_group("nccl", device_id=device)

    host_affinity = bind_to_gpu_numa(torch, device) if n_gpus > 1 else "single rank: not bound"
    a = c3_args(args.quick)
    # c4 / c5 are named with a GLOBAL population (BASELINE.json: n_chains = 2000 on 4 GPUs; PT with 512 chains on 8 GPUs),
    # partitioned over the ranks like the reference partitions n_chains over its workers (sampler/smc.py:423-427)
    global_chains = args.global_chains or ({"c4": 2000, "c5": 512}.get(CONFIG) if args.chains == 4000 else None)
    if global_chains:
        if global_chains % n_gpus:
            raise SystemExit("bench.py: n_chains / n_gpus has to be a whole number (%d / %d)" % (global_chains, n_gpus))
        args.chains = global_chains // n_gpus
    B = args.chains
    prob = synthetic.make_problem(interpolation=args.interpolation, seed=1234, build_library=False, noise=args.noise, **a)
    n_rot = 3
    Qs = [synthetic.draw_chains(prob, B, seed=4321 + 17 * rank + 1000 * r) for r in range(n_rot)]
    q_pin = [torch.from_numpy(q).pin_memory() for q in Qs]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def barrier():
        if n_gpus > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=device)
        if n_gpus > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def make_evaluator(store):
        ev = BatchedFFILogLike.from_problem(prob, device=local_rank, store_dtype="float32" if store == "f32" else "float64",
                                            upload_libraries=False)
        log("filling %.1f GB of synthetic GF library (%s) in HBM" % (lib_bytes(a, store) / 1e9, store))
        fill_library_on_device(ev, prob, torch, device, store)
        return ev

    def measure(ev, store, nchains, steps, warmup, gather_every_step, with_clocks=False, with_e2e=True):
        """K timed lock-step evaluations of `nchains` chains per rank: resident (`value`) and through the host-pointer
        entry (`e2e`).  The stack + misfit kernel time is summed from the event pairs the library records INSIDE the
        timed loop.  N > 1: one all-gather of the per-chain llk -- every step (the round-1 protocol) or once at the end
        of the K steps (the SMC stage boundary, timed separately)."""
        n_out = ev.n_out
        q_dev = [torch.from_numpy(q[:nchains]).to(device).contiguous() for q in Qs]
        logpts_dev = torch.empty((nchains, n_out), dtype=torch.float64, device=device)
        like_dev = torch.empty((nchains,), dtype=torch.float64, device=device)
        like_all = torch.empty((n_gpus * nchains,), dtype=torch.float64, device=device) if n_gpus > 1 else None
        logpts_pin = torch.empty((nchains, n_out), dtype=torch.float64).pin_memory()
        like_pin = torch.empty((nchains,), dtype=torch.float64).pin_memory()
        qp = [q_pin[r][:nchains] for r in range(n_rot)]             # leading rows of a pinned C-contiguous matrix: contiguous

        def step_resident(i):
            ev.eval_device(q_dev[i % n_rot], logpts_dev, like_dev)
            if n_gpus > 1 and gather_every_step:
                dist.all_gather_into_tensor(like_all, like_dev)

        def step_e2e(i):
            ev.eval_pinned(nchains, qp[i % n_rot].data_ptr(), logpts_pin.data_ptr(), like_pin.data_ptr())
            if n_gpus > 1 and gather_every_step:
                like_dev.copy_(like_pin, non_blocking=True)
                dist.all_gather_into_tensor(like_all, like_dev)

        step_resident(0)                            # binds the ctx to torch's current stream: torch events bracket our kernels
        torch.cuda.synchronize(device)
        viol = ev.ctx.index_violations()
        if viol:
            raise SystemExit("bench.py: %d library index violations in the synthetic chains" % viol)
        if not torch.isfinite(like_dev).all():
            raise SystemExit("bench.py: non-finite llk in the synthetic chains")
        sampler = ClockSampler(local_rank) if (rank == 0 and with_clocks) else None   # polling before the warm-up
        for i in range(warmup):
            step_resident(i)
        barrier()
        ev.ctx.stack_ms_accum(reset=True)
        launches0 = ev.ctx.launch_count()
        t_wall0 = time.perf_counter()
        e0.record()
        for i in range(steps):
            step_resident(i)
        e1.record()
        barrier()
        t_wall1 = time.perf_counter()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = ev.ctx.launch_count() - launches0
        k_sum, k_n = ev.ctx.stack_ms_accum(reset=True)             # the K event pairs recorded inside the timed loop
        clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
        out = {"ms_per_step": ms_total / steps, "value": n_gpus * nchains * steps / (ms_total / 1e3),
               "kernel_ms": k_sum / max(1, k_n), "kernel_evals": int(k_n),
               "kernel_share_of_step": max_over_ranks(k_sum) / ms_total, "launches": int(launches), "clocks": clocks}
        if n_gpus > 1 and not gather_every_step:
            # the stage-boundary exchange, on its own: llk of every chain to every rank (beat_b200.distributed)
            barrier()
            e0.record()
            dist.all_gather_into_tensor(like_all, like_dev)
            e1.record()
            barrier()
            out["allgather_llk_ms"] = max_over_ranks(e0.elapsed_time(e1))
        if with_e2e:
            for i in range(max(1, warmup // 2)):
                step_e2e(i)
            barrier()
            e0.record()
            for i in range(steps):
                step_e2e(i)
            e1.record()
            barrier()
            ms_e2e = max_over_ranks(e0.elapsed_time(e1))
            assert np.isfinite(like_pin.numpy()).all()
            out["e2e_value"] = n_gpus * nchains * steps / (ms_e2e / 1e3)
            out["e2e_ms_per_step"] = ms_e2e / steps
        out["q_dev"], out["bufs"] = q_dev, (logpts_dev, like_dev)
        return out

    # ---------------- headline: f32 library, `chains` per GPU (weak scaling), all-gather of llk every step for N > 1
    ev = make_evaluator(args.store)
    head = measure(ev, args.store, B, args.steps, args.warmup, gather_every_step=True, with_clocks=True)
    log("value %.0f evals/s (%.2f ms/step, stack+misfit %.2f ms)" % (head["value"], head["ms_per_step"], head["kernel_ms"]))
    blocking = ev.ctx.stack_blocking(ev.wmap_ids[0], len(prob["slip_vars"]))

    # ---------------- strong scaling: the configuration as BASELINE.json names it -- n_chains = 4000 partitioned over the N
    # ranks (reference: sampler/smc.py:423-427, base.py:534-535); the llk all-gather happens once per stage, as in SMC
    strong = None
    if n_gpus > 1 and not args.quick and B % n_gpus == 0 and CONFIG == "c3":
        per = B // n_gpus
        sm = measure(ev, args.store, per, args.steps, args.warmup, gather_every_step=False)
        t0 = ev.starttimes(per)                                          # rupture onset times of the last evaluated batch
        touched = touched_library_bytes(prob, a, args.store, args.interpolation, Qs[(args.steps - 1) % n_rot][:per], t0) if rank == 0 else 0
        peak = load_peaks()["hbm_gbs"]
        strong = {"global_chains": B, "chains_per_gpu": per, "value": sm["value"], "unit": UNIT, "ms_per_step": sm["ms_per_step"],
                  "e2e": sm.get("e2e_value"), "kernel_ms": sm["kernel_ms"], "kernel_share_of_step": sm["kernel_share_of_step"],
                  "allgather_llk_ms_per_stage": sm.get("allgather_llk_ms"),
                  "collective": "none inside the timed steps; one all-gather of llk[%d] per stage, timed separately" % B,
                  "touched_library_bytes_per_gpu": touched,
                  "hbm": {"achieved_GBps": touched / (sm["kernel_ms"] / 1e3) / 1e9 if touched else None, "peak_GBps": peak,
                          "frac": (touched / (sm["kernel_ms"] / 1e3) / 1e9 / peak) if touched else None,
                          "what": "distinct library rows the rank's chains gather in a step (each row crosses HBM once when "
                                  "the gather is L2-perfect) / stack+misfit kernel time"}}
        log("strong: %d chains/GPU -> %.0f evals/s (%.3f ms/step)" % (per, sm["value"], sm["ms_per_step"]))

    # ---------------- the same evaluation inside the lock-step Metropolis driver (propose -> bounds -> eval -> accept),
    # everything resident on the device: what a sampler stage actually achieves per GPU; then with the trace writer on
    from beat_b200.sampler import BatchedMetropolis
    lower = np.concatenate([prob["priors"][n][0] for n, _ in prob["var_order"]])
    upper = np.concatenate([prob["priors"][n][1] for n, _ in prob["var_order"]])

    def sampler_rate(cuda_graph, nchains=None):
        nchains = nchains or B
        m = BatchedMetropolis(ev.eval_device, lower, upper, nchains, device=device, tune=True, tune_interval=5, seed=rank, cuda_graph=cuda_graph)
        m.set_proposal_covariance(np.diag(((upper - lower) * 0.005) ** 2))
        m.beta = 0.1
        qs = head["q_dev"][0][:nchains].clone()
        lps, lks = m.initial_llk(qs)
        for _ in range(max(args.warmup, 4)):                               # graph mode captures at its third step
            qs, lps, lks, _ = m.step(qs, lps, lks)
        barrier()
        e0.record()
        for _ in range(args.steps):
            qs, lps, lks, _ = m.step(qs, lps, lks)
        e1.record()
        barrier()
        return n_gpus * nchains * args.steps / (max_over_ranks(e0.elapsed_time(e1)) / 1e3), m, (qs, lps, lks)

    sampler_eager, mh, (qs, lps, lks) = sampler_rate(False)
    try:
        sampler_value, mh_g, _ = sampler_rate(True)
        sampler_mode = "one CUDA graph per step"
    except Exception as e:                                                  # graph capture is an optimisation, never a requirement
        log("graph-captured sampler step failed (%r); reporting the eager step" % (e,))
        sampler_value, sampler_mode = sampler_eager, "eager (graph capture failed: %r)" % (e,)
    if strong is not None:
        try:
            strong["sampler_step"] = {"value": sampler_rate(True, strong["chains_per_gpu"])[0], "unit": "chain-steps/s",
                                      "eager_value": sampler_rate(False, strong["chains_per_gpu"])[0], "mode": "one CUDA graph per step"}
        except Exception as e:
            strong["sampler_step"] = {"value": None, "error": repr(e)}
    sampler_traced = None
    if not args.no_trace_writer:
        try:
            sampler_traced = sampler_with_trace_writer(args, prob, mh, qs, lps, lks, B, n_gpus, rank, barrier, max_over_ranks)
        except Exception as e:                                           # a full disk must not take the benchmark down
            log("trace-writer leg failed: %r" % (e,))
            sampler_traced = {"value": None, "error": repr(e)}

    pt_leg = None
    if CONFIG == "c5" and not args.quick:
        pt_leg = pt_driver_leg(args, ev, prob, lower, upper, B * n_gpus, n_gpus, rank, device, barrier, max_over_ranks)

    line = None
    if rank == 0:
        _pk = load_peaks()
        peak, peak_src = _pk["hbm_gbs"], _pk["source"]
        bytes_launch = algorithmic_bytes_per_eval(a, args.store, args.interpolation) * B
        k_ms = head["kernel_ms"]
        try:
            row_bytes = min(16384, max(16, (a["ns"] * (4 if args.store == "f32" else 8)) // 16 * 16))
            gather_peak = max(ev.ctx.probe_gather(0, 30 << 20, row_bytes, 2048, 3) for _ in range(2))
        except Exception as e:  # diagnostics only
            log("gather probe failed: %s" % e)
            gather_peak = None
        achieved = bytes_launch / (k_ms / 1e3) / 1e9
        # DRAM traffic of the stack kernel: from the committed ncu capture ONLY if it was taken with this very build
        traffic, tnote, l2_bytes = None, None, None
        build_hash = beatlib.source_hash()
        from beat_b200.build import kernel_hash
        tpath = os.path.join(ROOT, "profiles", "stack_kernel_traffic.json")
        if os.path.exists(tpath):
            try:
                tj = json.load(open(tpath))
                if (CONFIG == "c3" and NOISE == "exponential" and not args.quick and tj.get("chains") == B and tj.get("store") == args.store
                        and tj.get("interpolation") == args.interpolation and tj.get("stack_cuh_hash") == kernel_hash()
                        and tj.get("l2_blocking") == blocking):
                    traffic, tnote, l2_bytes = tj.get("dram_bytes_per_launch"), tj.get("note"), tj.get("l2_to_sm_bytes")
                else:
                    tnote = ("profiles/stack_kernel_traffic.json was captured on another kernel source / configuration (stack.cuh %s, blocking %s; "
                             "this run %s, %s): traffic withheld" % (tj.get("stack_cuh_hash"), tj.get("l2_blocking"), kernel_hash(), blocking))
            except Exception:
                pass
        n_sm, dev_name = ev.ctx.device_info()
        clocks = head["clocks"]
        line = {
            "metric": metric_name(), "value": head["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_LABEL[args.store], "data": "synthetic", "config": workload_config(args, n_gpus, B),
            "e2e": {"value": head["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": B * prob["n_params"] * 8,
                    "d2h_bytes_per_step": B * (ev.n_out + 1) * 8, "ms_per_step": head["e2e_ms_per_step"],
                    "vs_resident": head["e2e_value"] / head["value"], "host_affinity": host_affinity},
            "gpu_launches": head["launches"],
            "sampler_step": {"value": sampler_value, "unit": "chain-steps/s", "mode": sampler_mode, "eager_value": sampler_eager,
                             "what": "lock-step Metropolis step (proposal + bounds + batched eval + accept) with the population resident on the device",
                             "with_trace_writer": sampler_traced},
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "gf_stack_chunk_kernel+misfit_kernel (GF gather/stack + misfit)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bytes_launch, "kernel_ms": k_ms,
                         "kernel_ms_source": "mean of the %d CUDA-event pairs recorded around the kernels inside the timed loop" % head["kernel_evals"],
                         "kernel_share_of_step": head["kernel_share_of_step"],
                         "build_source_hash": build_hash,
                         "l2_blocking": blocking,
                         # frac > 1 is expected here, not a measurement error: the gather is L2-blocked, DRAM moves
                         # `traffic` bytes (ncu) for `algorithmic_bytes_per_launch` requested; the binding resource
                         # is the L2->SM fabric, reported below against 64 B/clk/SM
                         "l2_fabric": {"bytes_per_launch": l2_bytes,
                                       "achieved_GBps": (l2_bytes / (k_ms / 1e3) / 1e9) if l2_bytes else None,
                                       "peak_GBps": n_sm * 64 * ((clocks or {}).get("sm_mhz") or 1965.0) * 1e6 / 1e9,
                                       "peak_source": "n_sm x 64 B/clk x measured SM clock",
                                       # measured live: pseudo-random rows out of a 30 MB (L2-resident) working
                                       # set with the kernel's own access pattern (csrc/probe.cuh)
                                       "measured_gather_peak_GBps": gather_peak,
                                       "frac_of_measured_gather_peak": (achieved / gather_peak) if gather_peak else None},
                         "note": tnote},
        }
        if strong is not None:
            line["strong"] = strong
        if pt_leg is not None:
            line["pt_sampler"] = pt_leg
        if global_chains:
            line["scaling"] = "strong"
            line["config"]["global_chains"] = global_chains
            line["config"]["scaling_protocol"] = "strong: n_chains = %d partitioned over %d ranks" % (global_chains, n_gpus)
        if cpu_info is not None:
            line["cpu_baseline"] = cpu_info
    del head
    ev.close()
    del ev
    torch.cuda.empty_cache()

    # ---------------- strict mode: the library stored in f64, every operation in f64 (the reference's precision end to end)
    if not args.no_strict_f64 and args.store == "f32" and not args.quick and CONFIG == "c3":
        try:
            ev64 = make_evaluator("f64")
            s64 = measure(ev64, "f64", B, args.steps, args.warmup, gather_every_step=True)
            if rank == 0:
                b64 = algorithmic_bytes_per_eval(a, "f64", args.interpolation) * B
                peak = load_peaks()["hbm_gbs"]
                ach = b64 / (s64["kernel_ms"] / 1e3) / 1e9
                line["strict_f64"] = {"dtype": DTYPE_LABEL["f64"], "value": s64["value"], "unit": UNIT, "ms_per_step": s64["ms_per_step"],
                                      "e2e": s64["e2e_value"], "gf_library_GB_per_gpu": lib_bytes(a, "f64") / 1e9,
                                      "l2_blocking": ev64.ctx.stack_blocking(ev64.wmap_ids[0], len(prob["slip_vars"])),
                                      "roofline": {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                                                   "traffic": None, "algorithmic_bytes_per_launch": b64, "kernel_ms": s64["kernel_ms"],
                                                   "kernel_share_of_step": s64["kernel_share_of_step"]}}
                log("strict f64: %.0f evals/s (%.2f ms/step)" % (s64["value"], s64["ms_per_step"]))
            del s64
            ev64.close()
        except Exception as e:
            log("strict f64 leg failed: %r" % (e,))
            if rank == 0:
                line["strict_f64"] = {"value": None, "error": repr(e)}
    if n_gpus > 1:
        dist.destroy_process_group()
    return line


def pt_driver_leg(args, ev, prob, lower, upper, n_chains, n_gpus, rank, device, barrier, max_over_ranks):
    """BASELINE config 5 as named: the parallel-tempering driver with `n_chains` chains on a beta ladder, sharded over the
    ranks (beat_b200.sampler.pt_sample; swap rule of the reference, beat/sampler/pt.py:442-446; its master / worker
    processes :472-704 become one all-gather of (llk, scaling, acceptance) per swap interval)."""
    import torch
    from beat_b200 import sampler as S
    from beat_b200 import synthetic
    n_samples = max(200, 50 * args.steps)                      # long enough to amortise the per-call graph capture and warm-up
    pop = synthetic.draw_chains(prob, n_chains, seed=97)                  # identical on every rank
    kw = dict(device=device, swap_interval=(10, 15), n_chains_posterior=max(1, n_chains // 8), t_scale=1.2,
              beta_tune_interval=4 * n_chains, proposal_cov=np.diag(((upper - lower) * 0.005) ** 2), tune_interval=20,
              initial_population=pop, record_every=10 ** 9, cuda_graph=True)
    S.pt_sample(ev.eval_device, lower, upper, n_chains, 30, seed=1, **kw)          # warm-up (scratch, NCCL, allocator)
    barrier()
    t0 = time.perf_counter()
    out = S.pt_sample(ev.eval_device, lower, upper, n_chains, n_samples, seed=2, **kw)
    barrier()
    dt = max_over_ranks(time.perf_counter() - t0)
    return {"value": n_chains * n_samples / dt, "unit": "chain-steps/s", "n_chains": n_chains, "chains_per_gpu": n_chains // n_gpus,
            "n_samples_per_chain": n_samples, "ms_per_step": 1e3 * dt / n_samples, "swap_acceptance": out["swap_acceptance"],
            "n_evals": out["n_evals"], "t_scales": out["t_scales"][-3:],
            "what": "lock-step PT: Metropolis step for every chain of the ladder + swap proposals every 10-15 steps "
                    "(one all-gather of 3 doubles per chain); wall clock incl. the initial evaluation, max over ranks"}


def sampler_with_trace_writer(args, prob, mh, qs, lps, lks, B, n_gpus, rank, barrier, max_over_ranks):
    """The lock-step Metropolis step with every chain's record of every step written in the reference's binary trace
    format (beat/backend.py:651-897; beat_b200.backend.BatchedNumpyChains): the step's outputs are packed in record
    order on the device, copied into the writer's page-locked step buffers on a side stream (DeviceRecorder) and appended
    to the chain files by writer threads while the sampler goes on (`value`); for comparison the synchronous path of
    round 1 -- D2H, host-side packing, appends before the next step (`synchronous`).  4000 chain files per rank."""
    import shutil
    import tempfile
    from collections import OrderedDict
    import torch
    from beat_b200.backend import BatchedNumpyChains, DeviceRecorder
    shapes = OrderedDict()
    for name, n in prob["var_order"]:
        shapes[name] = (int(n),)
    shapes["seis_like"] = (int(lps.shape[1]),)
    shapes["like"] = ()
    off = prob["offsets"]
    device = qs.device
    out = {}
    d = tempfile.mkdtemp(prefix="beat_b200_trace_r%d_" % rank)
    try:
        # ---- asynchronous: device-packed records, page-locked double buffers, writer threads
        steps, n_thr = max(16, args.steps), 8
        w = BatchedNumpyChains(os.path.join(d, "async"), shapes, B, buffer_size=8, n_io_threads=n_thr, pinned=True)
        w.setup()
        rec = DeviceRecorder(w, torch, device)
        for _ in range(2):
            qs, lps, lks, _ = mh.step(qs, lps, lks)
            rec.record(qs, lps, lks)
        rec.finish()
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            qs, lps, lks, _ = mh.step(qs, lps, lks)
            rec.record(qs, lps, lks)
        rec.finish()                                                  # every record of every chain is in its file
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        w.close()
        nbytes = steps * B * w.data_structure.itemsize
        out = {"value": n_gpus * B * steps / dt, "unit": "chain-steps/s", "steps": steps, "writer_threads": n_thr,
               "bytes_written_per_rank": int(nbytes), "write_MBps_per_rank": nbytes / dt / 1e6,
               "what": "same step + records packed on the device, D2H into page-locked step buffers on a side stream, one NumpyChain-format "
                       "record per chain per step appended to %d chain files per rank by %d native writer threads (one writev per chain per flush of 8 steps; wall clock incl. the final flush)" % (B, n_thr)}
        shutil.rmtree(os.path.join(d, "async"), ignore_errors=True)

        # ---- synchronous (round-1 path)
        steps = max(2, min(args.steps, 4))
        w = BatchedNumpyChains(os.path.join(d, "sync"), shapes, B, buffer_size=steps)
        w.setup()
        host_q = torch.empty(qs.shape, dtype=torch.float64).pin_memory()
        host_lp = torch.empty(lps.shape, dtype=torch.float64).pin_memory()
        host_lk = torch.empty(lks.shape, dtype=torch.float64).pin_memory()

        def record(q, lp, lk):
            host_q.copy_(q, non_blocking=True); host_lp.copy_(lp, non_blocking=True); host_lk.copy_(lk, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            qn = host_q.numpy()
            vals = {name: qn[:, off[name]:off[name] + n] for name, n in prob["var_order"]}
            vals["seis_like"], vals["like"] = host_lp.numpy(), host_lk.numpy()
            w.write(vals)

        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            qs, lps, lks, _ = mh.step(qs, lps, lks)
            record(qs, lps, lks)
        w.flush()
        barrier()
        dt = max_over_ranks(time.perf_counter() - t0)
        out["synchronous"] = {"value": n_gpus * B * steps / dt, "steps": steps,
                              "what": "D2H, host-side packing and the appends inside the step (round-1 path)"}
        return out
    finally:
        shutil.rmtree(d, ignore_errors=True)


def run_c2_llk(args):
    """For the record (not the BASELINE metric): the log-likelihood stage of config 2 (DC point source, 32 stations x 3
    components x 2048 samples, Toeplitz covariance, 2000 chains) with residuals resident in HBM.  The geometry-mode
    forward model itself lives in pyrocko and is out of scope."""
    import torch
    from beat_b200.covariance import Covariance, exponential_data_covariance
    from beat_b200.lib import Context
    dev = torch.device("cuda", 0)
    nt, ns, B = 96, 2048, args.chains if args.chains != 4000 else 2000
    ctx = Context(0)
    wid = ctx.add_wavemap(nt, ns, "nearest_neighbor", None, np.zeros(nt, np.int32), np.full(nt, ns, np.int32))
    cov = Covariance(data=exponential_data_covariance(ns, 0.5, 2.0) * 0.05 ** 2)
    U = np.broadcast_to(cov.chol_inverse, (nt, ns, ns)) if args.noise != "dense" else None
    if args.noise == "dense":
        rng = np.random.default_rng(0)
        a = rng.random((ns, ns))
        cov = Covariance(data=(a.T.dot(a) + np.eye(ns) * 0.3) * 1e-3)
        U = np.broadcast_to(cov.chol_inverse, (nt, ns, ns))
    ctx.update_weights(wid, np.ascontiguousarray(U), np.full(nt, cov.log_pdet))
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream, external=True)
    res = torch.randn((B, nt, ns), dtype=torch.float64, device=dev) * 0.05
    hyp = torch.zeros((B, 1), dtype=torch.float64, device=dev)
    out = torch.empty((B, nt), dtype=torch.float64, device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(args.warmup):
        ctx.misfit_batch_dev(wid, B, res.data_ptr(), hyp.data_ptr(), 1, out.data_ptr())
    torch.cuda.synchronize()
    e0.record()
    for _ in range(args.steps):
        ctx.misfit_batch_dev(wid, B, res.data_ptr(), hyp.data_ptr(), 1, out.data_ptr())
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    bytes_step = B * nt * ns * 8
    print(json.dumps({"metric": "loglike evals/sec (C2 shapes: 96 datasets x 2048 samples, %s covariance; llk stage only, not the BASELINE metric)" % args.noise,
                      "value": B / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
                      "config": {"workload": "C2 llk stage", "chains_per_gpu": B},
                      "roofline": {"bound": "hbm" if args.noise != "dense" else "tensor(f64)", "achieved": bytes_step / (ms / 1e3) / 1e9, "unit": "GB/s",
                                   "algorithmic_bytes_per_launch": bytes_step}}))
    ctx.close()


def c2_problem(quick=False, ragged=False):
    """BASELINE config 2: DC point source, 32 stations x 3 components x 2048 samples (2 Hz, b..c = 1024 s), stepwise
    Butterworth filter, Toeplitz (exponential) covariance.  Synthetic type-A GF store (10 components), ~0.6 GB > L2."""
    from beat_b200 import synthetic
    if quick:
        return synthetic.make_geometry_problem(n_stations=4, seed=7)
    # default: every record spans the windows that can be asked of it (what a store built for the wavemap provides);
    # --ragged: records of uneven length that start late / end early, so most rows take the end-value-repeating path
    span = dict(nrec=2300, lead=60.0, ragged=True) if ragged else dict(nrec=2400, lead=90.0, ragged=False)
    return synthetic.make_geometry_problem(
        n_stations=32, ns=2048, taper=(-34.0, -24.0, 1000.0, 1010.0), dist_range=(2000e3, 4000e3),
        dx=4e3, dz=2.5e3, depth_range_km=(5.0, 30.0), duration_bounds=(0.0, 10.0), seed=7,
        filterer=[dict(kind="stepwise", order=4, lower_corner=0.005, upper_corner=0.2)], **span)


def load_peaks():
    """Roofline denominator: the driver-written measured HBM copy bandwidth, else the profiling recipe's fallback."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return {"hbm_gbs": json.load(open(peaks_path))["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json hbm_gbs)"}
    return {"hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


_C2 = {}


def _c2_cpu_eval(i):
    from oracle import geom_oracle
    from beat_b200 import synthetic
    g = _C2["gprob"]
    return geom_oracle.geometry_seismic_eval(g, synthetic.split_point(g, _C2["Q"][i % len(_C2["Q"])])).sum()


def run_c2(args):
    """For the record (not the BASELINE metric): the geometry-mode seismic forward model + llk of config 2 end to end
    on one GPU -- plan, TMA-staged delay-and-sum over the GF store, IIR filter + taper + chop + banded misfit."""
    import torch
    from beat_b200 import synthetic
    from beat_b200.geometry import BatchedGeometryLogLike
    dev = torch.device("cuda", 0)
    B = args.chains if args.chains != 4000 else 2000
    log("building the config-2 problem")
    gprob = c2_problem(args.quick, args.ragged)
    wm = gprob["wavemaps"][0]
    ev = BatchedGeometryLogLike.from_problem(gprob, device=0, upload_data=False)
    q_true = synthetic.draw_chains(gprob, 1, seed=1)
    synthetic.attach_geometry_data(gprob, ev.get_synthetics(q_true[0]))
    # banded weights straight from the Toeplitz covariance: the dense [nt, ns, ns] array is only materialised per target
    ev.upload_data(0, wm["data"], wm["U"], wm["slog_pdet"])
    Qs = [synthetic.draw_chains(gprob, B, seed=100 + i) for i in range(4)]
    q_dev = [torch.from_numpy(q).to(dev) for q in Qs]
    lp = torch.empty((B, ev.n_out), dtype=torch.float64, device=dev)
    lk = torch.empty((B,), dtype=torch.float64, device=dev)
    sampler = ClockSampler(0)
    for i in range(args.warmup):
        ev.eval_device(q_dev[i % 4], lp, lk)
    torch.cuda.synchronize()
    viol = ev.ctx.index_violations()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sum_ms = []
    launches0 = ev.ctx.launch_count()
    t_wall0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        ev.eval_device(q_dev[i % 4], lp, lk)
    e1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop(t_wall0, time.perf_counter())
    launches = ev.ctx.launch_count() - launches0
    ms = e0.elapsed_time(e1) / args.steps
    for i in range(args.steps):                                # kernel time of the delay-and-sum, one step at a time
        ev.eval_device(q_dev[i % 4], lp, lk)
        torch.cuda.synchronize()
        sum_ms.append(ev.ctx.last_stack_ms())
    k_ms = float(np.mean(sum_ms))
    finite = bool(torch.isfinite(lk).all().item())
    # e2e: pinned host q -> H2D -> kernels -> D2H(logpts, like)
    q_pin = [torch.from_numpy(q).pin_memory() for q in Qs]
    lp_pin = torch.empty((B, ev.n_out), dtype=torch.float64).pin_memory()
    lk_pin = torch.empty((B,), dtype=torch.float64).pin_memory()
    ev.ctx.set_stream(0, external=False)
    for i in range(args.warmup):
        ev.eval_pinned(B, q_pin[i % 4].data_ptr(), lp_pin.data_ptr(), lk_pin.data_ptr())
    t0 = time.perf_counter()
    for i in range(args.steps):
        ev.eval_pinned(B, q_pin[i % 4].data_ptr(), lp_pin.data_ptr(), lk_pin.data_ptr())
    e2e_s = (time.perf_counter() - t0) / args.steps
    peaks = load_peaks()
    alg = ev._bytes_per_eval * B
    line = {"metric": "forward+loglike evals/sec (C2 geometry-mode DC point source; not the BASELINE metric)",
            "value": B / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms,
            "higher_is_better": True, "dtype": "f64 (GF sum in f32 like the store)", "data": "synthetic",
            "config": {"workload": "C2 seismic DC point source: %d stations x 3 components x %d samples, %s, stepwise Butterworth "
                                   "order 4, exponential covariance" % (wm["nt"] // 3, wm["ns"], wm["interpolation"]),
                       "gf_records": "ragged (end values repeated inside most windows)" if args.ragged else "span every window",
                       "chains_per_gpu": B, "gf_store_MB": gprob["store"]["traces"].nbytes / 1e6,
                       "l2": "q rotates between steps; raw-trace scratch %.1f GB > L2" % (B * wm["nt"] * 2200 * 4 / 1e9)},
            "e2e": {"value": B / e2e_s, "unit": UNIT, "h2d_bytes_per_step": int(Qs[0].nbytes), "d2h_bytes_per_step": int(lp_pin.numel() * 8 + lk_pin.numel() * 8)},
            "gpu_launches": int(launches), "all_finite": finite, "index_violations_warmup": int(viol), "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "geom_plan_kernel+gf_delay_sum_kernel", "achieved": alg / (k_ms / 1e3) / 1e9,
                         "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": alg / (k_ms / 1e3) / 1e9 / peaks["hbm_gbs"], "traffic": None,
                         "peak_source": peaks["source"], "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
                         "kernel_share_of_step": k_ms / ms,
                         "note": "rows of one receiver are shared by all chains and by its 3 channels, so they are served from L2"}}
    if not args.no_cpu_baseline:
        n = 4 if not args.quick else 8
        cores = min(usable_cores(), n)
        _C2["gprob"], _C2["Q"] = gprob, Qs[0][:n]
        t0 = time.perf_counter()
        _c2_cpu_eval(0)
        one = time.perf_counter() - t0
        dt = pool_map(_c2_cpu_eval, list(range(n)), cores, timeout=600)
        line["cpu_baseline"] = {"value": n / dt, "unit": UNIT, "cores": cores, "kind": "port", "value_1core": 1.0 / one,
                                "sample": "%d chains, oracle restatement (numpy delay-and-sum + scipy lfilter + numpy llk), fork pool" % n}
    ev.close()
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="beat_b200", choices=["beat_b200", "reference"])
    ap.add_argument("--chains", type=int, default=4000, help="chains per GPU")
    ap.add_argument("--global-chains", type=int, default=0, help="population size partitioned over the ranks (c4: 2000, c5: 512 by default)")
    ap.add_argument("--store", default="f32", choices=["f32", "f64"], help="GF library storage dtype in HBM")
    ap.add_argument("--interpolation", default="multilinear", choices=["multilinear", "nearest_neighbor"])
    ap.add_argument("--quick", action="store_true", help="tiny shapes (development only; not a valid benchmark)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strict-f64", action="store_true", help="skip the strict (f64 library) leg of the C3 line")
    ap.add_argument("--no-trace-writer", action="store_true", help="skip the sampler step with the trace writer on")
    ap.add_argument("--noise", default="exponential", choices=["exponential", "variance", "dense"],
                    help="data covariance structure (exponential = BASELINE config; dense = full non-Toeplitz, for the record)")
    ap.add_argument("--ragged", action="store_true", help="config c2: GF records of uneven span (exercises the end-value path)")
    ap.add_argument("--config", default="c3", choices=["c3", "c3big", "c4", "c5", "c2llk", "c2"], help="c3 = BASELINE.json metric; c4/c5 for the record")
    args = ap.parse_args()
    global CONFIG, NOISE
    CONFIG, NOISE = args.config, args.noise
    args.warmup = max(args.warmup, 3)
    if args.config == "c2llk":
        run_c2_llk(args)
    elif args.config == "c2":
        run_c2(args)
    elif args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
