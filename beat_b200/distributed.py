"""
Multi-GPU plumbing: one process per GPU, chains sharded in contiguous blocks, ONE all-gather of per-chain
log-likelihoods where the sampler needs the whole population (the SMC stage boundary).

Reference behaviour this replaces: chains are fanned out over a fork pool in contiguous chunks
(beat/sampler/base.py:518-571, ``chunksize = ceil(n_chains / n_jobs)``), ``n_chains`` must be divisible by the
worker count (beat/sampler/smc.py:425-427), results travel through the file system and are reloaded by the
parent before ``select_end_points -> calc_beta -> resample`` (smc.py:486-524).  Here every rank keeps the static
operands resident in its own HBM (replicated), evaluates its block of chains, and the population-wide quantities
are exchanged with NCCL over NVLink (``torch.distributed``; ``gloo`` on CPU for the tests).  Within a Metropolis
step there is no communication at all.
"""
from __future__ import annotations

import os

import numpy as np


def env_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend=None, device=None):
    """Initialise torch.distributed from the torchrun environment (idempotent).  Returns (rank, world)."""
    import torch
    import torch.distributed as dist
    rank, local_rank, world = env_world()
    if world == 1:
        return 0, 1
    if not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {}
        if backend == "nccl":
            kw["device_id"] = device if device is not None else torch.device("cuda", local_rank)
        dist.init_process_group(backend, **kw)
    return dist.get_rank(), dist.get_world_size()


def shard_range(n_chains, rank, world):
    """Contiguous block of chains owned by ``rank``: chain c -> rank c // (n_chains / world)."""
    if n_chains % world != 0:
        raise ValueError("n_chains / n_jobs has to be a whole number!")         # smc.py:425-427
    per = n_chains // world
    return rank * per, (rank + 1) * per


def allgather_chains(local, world=None):
    """All-gather along the chain axis: local [B_local, ...] tensor -> [world * B_local, ...] on every rank.

    One collective (``all_gather_into_tensor``); rank order == chain order because shards are contiguous."""
    import torch
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    w = dist.get_world_size()
    out = torch.empty((w * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, local.contiguous())
    return out


def gather_objects(obj):
    """List of one small picklable object per rank, on every rank (checkpoint metadata; not on the data path)."""
    import torch.distributed as dist
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return [obj]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, obj)
    return out


class ShardedPopulation:
    """The population of one SMC stage, sharded over ranks.

    ``evaluator``: callable on a local parameter block.  For the GPU path pass
    ``BatchedFFILogLike.eval_device`` (torch CUDA tensors, NCCL); the CPU tests pass a numpy-backed callable with
    CPU tensors (gloo)."""

    def __init__(self, n_chains, evaluator):
        import torch.distributed as dist
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.n_chains = n_chains
        self.lo, self.hi = shard_range(n_chains, self.rank, self.world)
        self.evaluator = evaluator

    def local(self, population):
        """Slice this rank's block out of a replicated [n_chains, ...] array/tensor."""
        return population[self.lo:self.hi]

    def evaluate(self, q_local):
        """Evaluate the local block -> (logpts_local, like_local); no communication."""
        return self.evaluator(q_local)

    def gather_llk(self, like_local):
        """The per-stage exchange: every rank gets llk of all n_chains chains, in chain order."""
        return allgather_chains(like_local)

    def gather_population(self, q_local):
        """For resampling every rank also needs the particle matrix [n_chains, n_params]."""
        return allgather_chains(q_local)


def nonfinite_guard(like):
    """The sampler raises on a non-finite llk at stage 0 (beat/sampler/metropolis.py:279-284)."""
    arr = like.detach().cpu().numpy() if hasattr(like, "detach") else np.asarray(like)
    if not np.isfinite(arr).all():
        raise ValueError("Initial likelihood is not finite for chains %s" % np.flatnonzero(~np.isfinite(arr))[:8].tolist())
