#!/usr/bin/env python
"""
End-to-end example: SMC sampling of a synthetic geometry-mode problem (double-couple point source, BASELINE config 2
shapes scaled down by default) with the batched GPU evaluator.

    python examples/smc_c2_geometry.py [--full] [--chains 1000] [--steps 30] [--checkpoint-dir DIR [--resume]]
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 examples/smc_c2_geometry.py

Builds a synthetic type-A GF store and stations, makes data from a reference source with the GPU forward model itself,
runs the lock-step SMC driver (population resident on the device, chains sharded over ranks) and reports how well the
source parameters are recovered.
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="32 stations x 3 components x 2048 samples (config 2) instead of 8 x 3 x 120")
    ap.add_argument("--chains", type=int, default=1000)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--checkpoint-dir", default=None)
    ap.add_argument("--resume", action="store_true")
    args = ap.parse_args()

    import torch
    from beat_b200 import distributed as D
    from beat_b200 import sampler, synthetic
    from beat_b200.geometry import BatchedGeometryLogLike

    rank, local_rank, world = D.env_world()
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    D.init_process_group(device=dev)
    if args.full:
        gprob = synthetic.make_geometry_problem(
            n_stations=32, ns=2048, taper=(-34.0, -24.0, 1000.0, 1010.0), nrec=2400, lead=90.0, ragged=False,
            dist_range=(2000e3, 4000e3), dx=4e3, dz=2.5e3, depth_range_km=(5.0, 30.0), duration_bounds=(0.0, 10.0), seed=7,
            filterer=[dict(kind="stepwise", order=4, lower_corner=0.005, upper_corner=0.2)])
    else:
        gprob = synthetic.make_geometry_problem(n_stations=8, ns=120, taper=(-15.0, -10.0, 50.0, 55.0), nrec=400, lead=40.0,
                                                seed=7, filterer=[dict(kind="stepwise", order=4, lower_corner=0.01, upper_corner=0.5)])
    ev = BatchedGeometryLogLike.from_problem(gprob, device=local_rank, upload_data=False)
    q_true = synthetic.draw_chains(gprob, 1, seed=1)[0]
    q_true[gprob["offsets"]["hypers"]] = 0.0
    synthetic.attach_geometry_data(gprob, ev.get_synthetics(q_true), rel_sigma=0.1)      # same data on every rank (seeded)
    wm = gprob["wavemaps"][0]
    ev.upload_data(0, wm["data"], wm["U"], wm["slog_pdet"])
    lower = np.concatenate([gprob["priors"][n][0] for n, _ in gprob["var_order"]])
    upper = np.concatenate([gprob["priors"][n][1] for n, _ in gprob["var_order"]])

    t0 = time.perf_counter()
    out = sampler.smc_sample(ev.eval_device, lower, upper, n_chains=args.chains, n_steps=args.steps, device=dev, seed=3,
                             checkpoint_dir=args.checkpoint_dir, resume=args.resume,
                             log=(lambda m: print("[smc] " + m, flush=True)) if rank == 0 else None)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if rank == 0:
        print("stages %d, %d evaluations in %.2f s = %.0f evals/s on %d GPU(s)" % (out["n_stages"], out["n_evals"], dt, out["n_evals"] / dt, world))
        best = out["population"][int(np.argmax(out["likelihoods"]))]
        names = [n for n, s in gprob["var_order"] for _ in range(s)]
        print("%-14s %10s %10s %10s" % ("variable", "true", "MAP", "post. std"))
        for i, n in enumerate(names):
            print("%-14s %10.3f %10.3f %10.3f" % (n, q_true[i], best[i], out["population"][:, i].std()))
        print("llk(true) %.2f  llk(MAP) %.2f" % (ev(q_true[None, :])[1][0], out["likelihoods"].max()))
    ev.close()
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
