"""world_size-2 gloo tests (CPU) of the N>1 host logic: contiguous chain shards + one all-gather of llk reproduce
the single-process population exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_eval(q):
    """Stand-in evaluator with the engine's return convention (logpts [B, n_out], like [B])."""
    logpts = torch.stack([-(q ** 2).sum(dim=1), -q.abs().sum(dim=1), q[:, 0] * 0.5], dim=1)
    return logpts, logpts.sum(dim=1)


def _worker(rank, world, port, n_chains, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from beat_b200 import distributed as D
    r, w = D.init_process_group(backend="gloo")
    assert (r, w) == (rank, world)
    rng = np.random.default_rng(7)
    Q = torch.from_numpy(rng.standard_normal((n_chains, 11)))           # replicated population
    pop = D.ShardedPopulation(n_chains, _fake_eval)
    q_local = pop.local(Q)
    assert q_local.shape[0] == n_chains // world
    logpts_local, like_local = pop.evaluate(q_local)
    like_all = pop.gather_llk(like_local)
    Q_all = pop.gather_population(q_local)
    torch.save({"like": like_all, "Q": Q_all, "lo": pop.lo, "hi": pop.hi}, os.path.join(out_dir, "r%d.pt" % rank))
    with pytest.raises(ValueError):
        D.shard_range(n_chains + 1, rank, world)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_sharded_population_matches_single_process(tmp_path, world):
    n_chains = 12
    port = _free_port()
    mp.spawn(_worker, args=(world, port, n_chains, str(tmp_path)), nprocs=world, join=True)
    rng = np.random.default_rng(7)
    Q = torch.from_numpy(rng.standard_normal((n_chains, 11)))
    _, like_ref = _fake_eval(Q)
    for r in range(world):
        got = torch.load(os.path.join(str(tmp_path), "r%d.pt" % r))
        assert torch.equal(got["like"], like_ref)
        assert torch.equal(got["Q"], Q)
        assert (got["lo"], got["hi"]) == (r * n_chains // world, (r + 1) * n_chains // world)


def test_shard_range_contract():
    from beat_b200.distributed import shard_range
    assert [shard_range(4000, r, 8) for r in (0, 7)] == [(0, 500), (3500, 4000)]
    with pytest.raises(ValueError, match="whole number"):
        shard_range(10, 0, 4)


def _smc_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from beat_b200 import distributed as D
    from beat_b200 import sampler as S
    D.init_process_group(backend="gloo")
    n = 4
    mu1 = torch.ones(n, dtype=torch.float64) * 0.5

    def ev(q):          # the two-Gaussian toy posterior of the reference's sampler tests
        l1 = -0.5 * 100.0 * ((q - mu1) ** 2).sum(dim=1)
        l2 = -0.5 * 100.0 * ((q + mu1) ** 2).sum(dim=1)
        like = torch.logsumexp(torch.stack([np.log(0.1) + l1, np.log(0.9) + l2]), dim=0)
        return like[:, None].clone(), like
    out = S.smc_sample(ev, -2.0 * np.ones(n), 2.0 * np.ones(n), n_chains=600, n_steps=60, tune_interval=20, seed=11)
    np.savez(os.path.join(out_dir, "smc_r%d.npz" % rank), pop=out["population"], like=out["likelihoods"], betas=np.array(out["betas"]),
             n_evals=out["n_evals"])
    dist.barrier()
    dist.destroy_process_group()


def test_smc_sharded_over_two_ranks(tmp_path):
    """The lock-step SMC driver with chains sharded over 2 ranks (gloo): every rank ends with the same replicated
    population (one all-gather per stage), and the posterior is recovered."""
    port = _free_port()
    mp.spawn(_smc_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(os.path.join(str(tmp_path), "smc_r0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "smc_r1.npz"))
    assert np.array_equal(r0["pop"], r1["pop"]) and np.array_equal(r0["like"], r1["like"])
    assert np.array_equal(r0["betas"], r1["betas"]) and r0["betas"][-1] == 1.0
    assert r0["pop"].shape == (600, 4) and int(r0["n_evals"]) == int(r1["n_evals"]) > 600
    np.testing.assert_allclose(np.abs(r0["pop"]).mean(axis=0), 0.5, atol=0.04)


def _smc_ckpt_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from beat_b200 import distributed as D
    from beat_b200 import sampler as S
    D.init_process_group(backend="gloo")
    n = 3
    mu = torch.ones(n, dtype=torch.float64) * 0.3

    def ev(q):
        like = -0.5 * 400.0 * ((q - mu) ** 2).sum(dim=1)
        return like[:, None].clone(), like
    kw = dict(n_chains=120, n_steps=8, seed=3)
    lo, hi = -2.0 * np.ones(n), 2.0 * np.ones(n)
    full = S.smc_sample(ev, lo, hi, **kw)
    S.smc_sample(ev, lo, hi, checkpoint_dir=os.path.join(out_dir, "ck"), max_stages=2, **kw)
    dist.barrier()                                                  # rank 0 has written stage_2.npz
    res = S.smc_sample(ev, lo, hi, checkpoint_dir=os.path.join(out_dir, "ck"), resume=True, **kw)
    np.savez(os.path.join(out_dir, "ck_r%d.npz" % rank), full=full["population"], res=res["population"],
             fb=np.array(full["betas"]), rb=np.array(res["betas"]), fe=full["n_evals"], re=res["n_evals"])
    dist.barrier()
    dist.destroy_process_group()


def test_smc_checkpoint_resume_two_ranks(tmp_path):
    """Stage checkpoint written by rank 0 (end points of all chains, host RNG, per-rank proposal scaling + device RNG):
    a 2-rank run resumed after stage 2 reproduces the uninterrupted 2-rank run bit for bit on both ranks."""
    port = _free_port()
    mp.spawn(_smc_ckpt_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), "ck_r%d.npz" % r))
        assert np.array_equal(z["fb"], z["rb"]) and z["fb"][-1] == 1.0
        assert np.array_equal(z["full"], z["res"])
        assert int(z["fe"]) == int(z["re"])


def _pt_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from beat_b200 import distributed as D
    from beat_b200 import sampler as S
    D.init_process_group(backend="gloo")
    n = 4
    mu1 = torch.ones(n, dtype=torch.float64) * 0.5

    def ev(q):          # the two-Gaussian toy posterior of the reference's test/test_pt.py
        l1 = -0.5 * 100.0 * ((q - mu1) ** 2).sum(dim=1)
        l2 = -0.5 * 100.0 * ((q + mu1) ** 2).sum(dim=1)
        like = torch.logsumexp(torch.stack([np.log(0.1) + l1, np.log(0.9) + l2]), dim=0)
        return like[:, None].clone(), like
    out = S.pt_sample(ev, -2.0 * np.ones(n), 2.0 * np.ones(n), n_chains=16, n_samples=5000, swap_interval=(10, 15),
                      n_chains_posterior=4, t_scale=1.6, beta_tune_interval=200, seed=5)
    np.savez(os.path.join(out_dir, "pt_r%d.npz" % rank), samples=out["samples"], levels=out["levels"], betas=out["betas"],
             chain_betas=out["chain_betas"], acc=out["swap_acceptance"], n_evals=out["n_evals"], pop=out["population"])
    dist.barrier()
    dist.destroy_process_group()


def test_pt_sharded_over_two_ranks(tmp_path):
    """Parallel tempering with the chains sharded over 2 ranks (gloo): the swap phase is one all-gather of (llk, scaling,
    acceptance) per chain, decisions are identical on both ranks (shared host RNG), chains trade temperature LEVELS so no
    state crosses ranks -- and the beta = 1 samples still recover the two-mode posterior of the reference's test
    (swap rule: /root/reference/beat/sampler/pt.py:442-446)."""
    port = _free_port()
    mp.spawn(_pt_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(os.path.join(str(tmp_path), "pt_r0.npz"))
    r1 = np.load(os.path.join(str(tmp_path), "pt_r1.npz"))
    for k in ("samples", "levels", "betas", "chain_betas", "pop"):
        assert np.array_equal(r0[k], r1[k]), k                      # both ranks hold the same global picture
    assert sorted(r0["levels"].tolist()) == list(range(16))         # the ladder is a permutation of the chains
    assert r0["betas"][0] == 1.0 and np.all(np.diff(r0["betas"][3:]) < 0)
    assert 0.0 < float(r0["acc"]) <= 1.0 and int(r0["n_evals"]) > 16 * 1000
    assert (r0["levels"] != np.arange(16)).any()                    # swaps did happen, across the rank boundary too
    x = r0["samples"][len(r0["samples"]) // 5:]
    assert x.shape[0] > 10000
    np.testing.assert_allclose(np.abs(x).mean(axis=0), 0.5, rtol=0.0, atol=0.05)
    assert 0.02 < (x[:, 0] > 0).mean() < 0.35


def _trace_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), LOCAL_RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from collections import OrderedDict
    from beat_b200 import backend as bk
    from beat_b200 import distributed as D
    from beat_b200 import sampler as S
    D.init_process_group(backend="gloo")
    n, n_chains, n_steps = 3, 40, 9
    mu = torch.tensor([0.3, -0.2, 0.1], dtype=torch.float64)

    def ev(q):
        lp = -0.5 * ((q - mu) ** 2 / 0.05).sum(dim=1, keepdim=True)
        return lp, lp[:, 0]

    shapes = OrderedDict([("x", (n,)), ("seis_like", (1,)), ("like", ())])
    writers = {}

    def on_step(stage, step, q, logpts, like):
        # every rank writes the files of ITS chains: chain_offset = first global chain of the shard; native appends
        if stage not in writers:
            n_local = q.shape[0]
            writers[stage] = bk.BatchedNumpyChains(os.path.join(out_dir, "stage_%d" % stage), shapes, n_local, buffer_size=4,
                                                   chain_offset=rank * n_local, n_io_threads=2)
            writers[stage].setup()
        writers[stage].write_records(torch.cat([q, logpts, like[:, None]], dim=1).numpy())

    res = S.smc_sample(ev, -np.ones(n), np.ones(n), n_chains, n_steps, seed=3, on_step=on_step)
    for w in writers.values():
        w.close()
    np.savez(os.path.join(out_dir, "trace_r%d.npz" % rank), pop=res["population"], like=res["likelihoods"], n_stages=res["n_stages"])
    dist.barrier()
    dist.destroy_process_group()


def test_trace_files_written_by_two_ranks(tmp_path):
    """Two ranks stream the steps of their chain shards into one stage directory (disjoint chain numbers through
    ``chain_offset``, background native appends): afterwards every chain of the population has its file with n_steps
    records whose last one is that chain's end point."""
    from beat_b200 import backend as bk
    port = _free_port()
    mp.spawn(_trace_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0 = np.load(os.path.join(str(tmp_path), "trace_r0.npz"))
    n_stages, pop, like = int(r0["n_stages"]), r0["pop"], r0["like"]
    assert n_stages >= 2 and pop.shape == (40, 3)
    last = os.path.join(str(tmp_path), "stage_%d" % n_stages)
    assert sorted(os.listdir(last)) == sorted("chain-%d.bin" % c for c in range(40))
    for c in range(40):
        x = bk.get_values(os.path.join(last, "chain-%d.bin" % c), "x")
        assert x.shape == (9, 3)
        np.testing.assert_array_equal(x[-1], pop[c])
        np.testing.assert_array_equal(bk.get_values(os.path.join(last, "chain-%d.bin" % c), "like")[-1], like[c])
