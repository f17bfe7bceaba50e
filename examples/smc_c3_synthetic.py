#!/usr/bin/env python
"""
End-to-end example: SMC sampling of a synthetic FFI seismic problem with the batched GPU evaluator.

    python examples/smc_c3_synthetic.py [--small] [--chains 2000] [--steps 50]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 examples/smc_c3_synthetic.py

Builds the problem (synthetic library generated in HBM), runs the lock-step SMC driver with the whole population
resident on the device(s), optionally writes NumpyChain-compatible traces, prints stage statistics and evals/s.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--small", action="store_true", help="60 patches x 16 targets instead of the full C3 shapes")
    ap.add_argument("--chains", type=int, default=2000)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--trace-dir", default=None, help="write chain-<i>.bin traces (reference 'bin' backend layout)")
    args = ap.parse_args()

    import torch
    from beat_b200 import distributed as D
    from beat_b200 import sampler, synthetic
    from beat_b200.devlib import fill_library_on_device
    from beat_b200.engine import BatchedFFILogLike

    rank, local_rank, world = D.env_world()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    D.init_process_group(device=dev)
    shapes = dict(nt=16, subfaults=((6, 10, 2.0),), ns=64, ndur=9) if args.small else \
        dict(nt=64, subfaults=((10, 20, 2.0),), ns=120, ndur=17, nst=64)
    prob = synthetic.make_problem(build_library=False, seed=1234, **shapes)
    ev = BatchedFFILogLike.from_problem(prob, device=local_rank, store_dtype="float32", upload_libraries=False)
    fill_library_on_device(ev, prob, torch, dev, "f32")
    lower = np.concatenate([prob["priors"][n][0] for n, _ in prob["var_order"]])
    upper = np.concatenate([prob["priors"][n][1] for n, _ in prob["var_order"]])

    on_step = None
    if args.trace_dir:                                   # every rank writes the files of its own chains
        from collections import OrderedDict
        from beat_b200.backend import BatchedNumpyChains, DeviceRecorder
        shapes_out = OrderedDict([(n, (s,)) for n, s in prob["var_order"]] + [("seis_like", (ev.n_out,)), ("like", ())])
        writers, recorders = {}, {}

        def on_step(stage, step, q, logpts, like):
            # records packed on the device (q's columns are in var_order: record = q | seis_like | like), copied into
            # page-locked step buffers on a side stream, appended to the chain files by writer threads
            if stage not in writers:
                for st in writers:
                    recorders[st].finish()
                w = writers[stage] = BatchedNumpyChains(os.path.join(args.trace_dir, "stage_%d" % stage), shapes_out, q.shape[0],
                                                        buffer_size=min(16, args.steps), chain_offset=rank * q.shape[0],
                                                        n_io_threads=4, pinned=True)
                w.setup()
                recorders[stage] = DeviceRecorder(w, torch, q.device)
            recorders[stage].record(q, logpts, like)

    t0 = time.perf_counter()
    out = sampler.smc_sample(ev.eval_device, lower, upper, n_chains=args.chains, n_steps=args.steps, device=dev, seed=1,
                             log=(print if rank == 0 else None), on_step=on_step)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if on_step:
        for st, w in writers.items():
            recorders[st].finish()
            w.close()
    if rank == 0:
        print("stages %d, %d forward+loglike evaluations in %.1f s = %.0f evals/s on %d GPU(s); median llk %.2f"
              % (out["n_stages"], out["n_evals"], dt, out["n_evals"] / dt, world, np.median(out["likelihoods"])))
    ev.close()


if __name__ == "__main__":
    main()
