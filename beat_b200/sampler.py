"""
Lock-step batched Metropolis / SMC driver: the shim that lets the batched GPU evaluator serve the reference's
sampler logic.

The reference advances ONE chain per ``Metropolis.astep`` call (beat/sampler/metropolis.py:276-422) and
parallelises chains over forked processes (beat/sampler/base.py:428-595).  Here every chain of the population
takes the same step at the same time:

    for step in range(n_steps):
        q      = q0 + proposal_sample * scaling                    (metropolis.py:311,325)
        inside = prior bounds check                                 (metropolis.py:341-343; Uniform priors ->
                                                                     finite prior logp <=> inside the box)
        lp     = ONE batched GPU evaluation of all chains           (metropolis.py:349)
        accept = log(u) < beta * (lp - l0)   (pymc metrop_select)   (metropolis.py:355-358)
        per-chain scale tuning every tune_interval steps            (metropolis.py:294-306, pymc ``tune``)

Stage logic is the reference's, restated on whole-population arrays: ``calc_beta`` (beat/sampler/smc.py:133-165),
``calc_covariance`` (:167-186), Kitagawa ``resample`` (:290-324), the stage loop of ``smc_sample`` (:459-546)
including the final stage.  Per-chain RNG streams are not bitwise those of the reference (it seeds each forked
chain separately, sampler/base.py:515-516) -- the claim is distributional equivalence, tested on the reference's
own toy posterior (test/test_smc.py).  Trace storage / checkpointing stay with the reference (out of scope).

All population state lives in torch tensors on the evaluator's device; the only host round-trip per stage is the
scalar bisection for beta.  With several ranks (torch.distributed) each rank advances its shard and the
population-wide quantities are all-gathered once per stage (beat_b200.distributed).
"""
from __future__ import annotations

import numpy as np

from .covariance import ensure_cov_psd        # the one faithful restatement of beat/utility.py:1034-1056 (+ repair_covariance)


def tune_scale(scale, acc_rate):
    """pymc's Metropolis ``tune`` rule, used by the reference as ``step_tune`` (metropolis.py:300-302), vectorised.

    Rate    Variance adaptation
    <0.001  x 0.1 ; <0.05  x 0.5 ; <0.2  x 0.9 ; >0.5  x 1.1 ; >0.75  x 2 ; >0.95  x 10"""
    import torch
    s = scale.clone()
    s = torch.where(acc_rate < 0.001, scale * 0.1, s)
    s = torch.where((acc_rate >= 0.001) & (acc_rate < 0.05), scale * 0.5, s)
    s = torch.where((acc_rate >= 0.05) & (acc_rate < 0.2), scale * 0.9, s)
    s = torch.where(acc_rate > 0.95, scale * 10.0, s)
    s = torch.where((acc_rate > 0.75) & (acc_rate <= 0.95), scale * 2.0, s)
    s = torch.where((acc_rate > 0.5) & (acc_rate <= 0.75), scale * 1.1, s)
    return s


def tune_pt_scale(scale, acc_rate):
    """Temperature-scale adaptation of the PT ladder (beat/sampler/pt.py:37-73 ``tune``; applied by
    ``TemperingManager.tune_betas``, :331-352, and clipped there to [1.01, 2.0]): gentler than the Metropolis step rule.

    Rate    scale
    <0.001  x 0.85 ; <0.05  x 0.9 ; <0.2  x 0.95 ; >0.5  x 1.05 ; >0.75  x 1.10 ; >0.95  x 1.15"""
    if acc_rate < 0.001:
        return scale * 0.85
    if acc_rate < 0.05:
        return scale * 0.9
    if acc_rate < 0.2:
        return scale * 0.95
    if acc_rate > 0.95:
        return scale * 1.15
    if acc_rate > 0.75:
        return scale * 1.10
    if acc_rate > 0.5:
        return scale * 1.05
    return scale


def calc_beta(likelihoods, beta, coef_variation=1.0):
    """Next tempering beta + importance weights by bisection (beat/sampler/smc.py:133-165), numpy on [n_chains]."""
    likelihoods = np.asarray(likelihoods, dtype=np.float64)
    low_beta, up_beta, old_beta = beta, 2.0, beta
    lmax = likelihoods.max()
    while up_beta - low_beta > 1e-6:
        current_beta = (low_beta + up_beta) / 2.0
        temp = np.exp((current_beta - beta) * (likelihoods - lmax))
        cov_temp = np.std(temp) / np.mean(temp)
        if cov_temp > coef_variation:
            up_beta = current_beta
        else:
            low_beta = current_beta
    return current_beta, old_beta, temp / np.sum(temp)


def proposal_factor(cov):
    """A factor F with F F^T = cov for the MultivariateNormal proposal (sampler/base.py:163-167 draws with
    numpy's SVD-based ``multivariate_normal``, which accepts a semi-definite covariance): Cholesky when it exists,
    else the symmetric eigen-factor of the repaired matrix -- same distribution either way."""
    cov = ensure_cov_psd(np.atleast_2d(cov))
    try:
        return np.linalg.cholesky(cov)
    except np.linalg.LinAlgError:
        w, v = np.linalg.eigh((cov + cov.T) / 2.0)
        return v * np.sqrt(np.clip(w, 0.0, None))


def calc_covariance(array_population, weights):
    """Weighted sample covariance of the population (beat/sampler/smc.py:167-186)."""
    cov = np.cov(array_population, aweights=np.asarray(weights).ravel(), bias=False, rowvar=0)
    cov = ensure_cov_psd(np.atleast_2d(cov))
    if np.isnan(cov).any() or np.isinf(cov).any():
        raise ValueError("Sample covariances contains Inf or NaN! Please try reducing the upper and lower bounds of hyper parameters!")
    return cov


def resample(weights, rng):
    """Kitagawa's deterministic resampling (beat/sampler/smc.py:290-324), vectorised: same output for the same u."""
    n = len(weights)
    cum_dist = np.cumsum(weights)
    u = (np.arange(n) + rng.random()) / n
    j = np.searchsorted(cum_dist, u, side="left")      # first j with cum_dist[j] >= u  <=>  `while u > cum_dist[j]: j += 1`
    j = np.minimum(j, n - 1)
    n_childs = np.bincount(j, minlength=n)
    return np.repeat(np.arange(n), n_childs)


class BatchedMetropolis:
    """All chains of one rank advanced in lock-step.  ``evaluator(q_dev) -> (logpts, like)`` on torch tensors.

    ``cuda_graph=True`` (CUDA devices only): after two eager steps the whole step -- proposal draw, bounds check, the
    batched evaluation (sweep -> stack -> misfit -> sum, launched by libbeatgpu on the capturing stream), accept / reject
    and the in-place state update -- is captured into ONE CUDA graph and replayed.  With few chains per GPU (n_chains =
    4000 over 8 GPUs = 500 per rank) the ~30 launches of a step otherwise cost more host time than the kernels take.
    The population state then lives in fixed buffers that ``step`` updates in place and returns."""

    def __init__(self, evaluator, lower, upper, n_chains, device=None, scale=1.0, tune=True, tune_interval=100, seed=0,
                 cuda_graph=False):
        import torch
        self.torch = torch
        self.evaluator = evaluator
        self.device = device if device is not None else torch.device("cpu")
        self.lower = torch.as_tensor(np.asarray(lower, dtype=np.float64), device=self.device)
        self.upper = torch.as_tensor(np.asarray(upper, dtype=np.float64), device=self.device)
        self.n_chains = n_chains
        self.n_params = self.lower.numel()
        self.scale0 = float(scale)
        self.scaling = torch.full((n_chains,), float(scale), dtype=torch.float64, device=self.device)
        self.tune, self.tune_interval = tune, tune_interval
        self.steps_until_tune = tune_interval
        self.accepted = torch.zeros(n_chains, dtype=torch.float64, device=self.device)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed))
        self._beta_t = torch.ones(n_chains, dtype=torch.float64, device=self.device)   # per-chain beta (SMC: all equal; PT: ladder)
        self._n_evals = torch.zeros((), dtype=torch.float64, device=self.device)   # device-side counter: no host sync per step
        self.chol = None
        self.cuda_graph = bool(cuda_graph) and torch.device(self.device).type == "cuda"
        self._graph, self._gstream, self._eager_steps, self._state = None, None, 0, None

    @property
    def beta(self):
        return self._beta_t

    @beta.setter
    def beta(self, value):
        """Scalar (SMC stage beta) or one value per chain (PT ladder); written in place (the captured graph reads it)."""
        torch = self.torch
        v = torch.as_tensor(value, dtype=torch.float64, device=self.device) if not torch.is_tensor(value) else value.to(self.device, torch.float64)
        self._beta_t.copy_(v.expand(self.n_chains) if v.ndim == 0 else v)

    @property
    def n_evals(self):
        return int(self._n_evals.item())

    def start_stage(self):
        """Every chain of a stage starts from the PARENT step object's state: the reference pickles ``step`` into each
        forked chain (sampler/base.py:260-313), so per-chain tuning of the previous stage is never carried over --
        scaling back to the constructor's value, tuning counters cleared (metropolis.py:97-106)."""
        self.scaling.fill_(self.scale0)
        self.steps_until_tune = self.tune_interval
        self.accepted.zero_()

    def set_proposal_covariance(self, cov):
        """MultivariateNormal proposal (sampler/base.py:163-167): draws = z @ chol(cov).T."""
        f = self.torch.as_tensor(proposal_factor(cov), device=self.device)
        if self.chol is not None and self.chol.shape == f.shape:
            self.chol.copy_(f)                       # in place: a captured graph keeps reading the same buffer
        else:
            self.chol = f.contiguous()
            self._graph = None

    def initial_llk(self, q):
        """Stage 0: evaluate the start population; non-finite llk raises (metropolis.py:277-284)."""
        logpts, like = self.evaluator(q)
        self._n_evals += q.shape[0]
        if not bool(self.torch.isfinite(like).all()):
            raise ValueError("Got NaN in likelihood evaluation! Invalid model definition? Or starting point outside prior bounds!")
        return logpts, like

    def _tune_if_due(self):
        if self.tune and self.steps_until_tune == 0:
            self.scaling.copy_(tune_scale(self.scaling, self.accepted / float(self.tune_interval)))
            self.steps_until_tune = self.tune_interval
            self.accepted.zero_()

    def _step_body(self, q0, logpts0, like0):
        torch = self.torch
        z = torch.randn((self.n_chains, self.n_params), dtype=torch.float64, device=self.device, generator=self.gen)
        delta = (z @ self.chol.T) * self.scaling[:, None]
        q = q0 + delta
        inside = ((q >= self.lower) & (q <= self.upper)).all(dim=1)
        # out-of-prior proposals are rejected without being trusted: evaluate a safe copy (the previous point) there
        q_eval = torch.where(inside[:, None], q, q0).contiguous()
        logpts, like = self.evaluator(q_eval)
        self._n_evals += inside.sum()
        log_u = torch.log(torch.rand(self.n_chains, dtype=torch.float64, device=self.device, generator=self.gen))
        ratio = self._beta_t * (like - like0)
        accept = inside & torch.isfinite(ratio) & (log_u < ratio)          # pymc metrop_select
        q_new = torch.where(accept[:, None], q, q0)
        logpts_new = torch.where(accept[:, None], logpts, logpts0)
        like_new = torch.where(accept, like, like0)
        self.accepted += accept.to(torch.float64)
        return q_new, logpts_new, like_new, accept

    def step(self, q0, logpts0, like0):
        """One lock-step Metropolis step for every chain.  Returns (q_new, logpts_new, like_new, accepted_mask)."""
        self._tune_if_due()
        self.steps_until_tune -= 1
        if not self.cuda_graph:
            return self._step_body(q0, logpts0, like0)
        return self._step_graphed(q0, logpts0, like0)

    # ------------------------------------------------------------------ CUDA-graph path
    def _step_graphed(self, q0, logpts0, like0):
        torch = self.torch
        if self._gstream is None:
            self._gstream = torch.cuda.Stream(device=self.device)
        cur = torch.cuda.current_stream(self.device)
        self._gstream.wait_stream(cur)
        with torch.cuda.stream(self._gstream):
            if self._state is None or self._state[0].shape != q0.shape or self._state[1].shape != logpts0.shape:
                self._state = (q0.clone(), logpts0.clone(), like0.clone(), torch.zeros(self.n_chains, dtype=torch.bool, device=self.device))
                self._graph, self._eager_steps = None, 0
            sq, slp, slk, sacc = self._state
            if q0.data_ptr() != sq.data_ptr():                    # a new population was handed in (stage start, swap, resample)
                sq.copy_(q0); slp.copy_(logpts0); slk.copy_(like0)
            if self._graph is None and self._eager_steps >= 2:
                g = torch.cuda.CUDAGraph()
                g.register_generator_state(self.gen)
                with torch.cuda.graph(g, stream=self._gstream):
                    qn, lpn, lkn, acc = self._step_body(sq, slp, slk)
                    sq.copy_(qn); slp.copy_(lpn); slk.copy_(lkn); sacc.copy_(acc)
                self._graph = g
            if self._graph is not None:
                self._graph.replay()
            else:                                               # the first steps run eagerly on the same stream (warm-up)
                qn, lpn, lkn, acc = self._step_body(sq, slp, slk)
                sq.copy_(qn); slp.copy_(lpn); slk.copy_(lkn); sacc.copy_(acc)
                self._eager_steps += 1
        cur.wait_stream(self._gstream)
        return sq, slp, slk, sacc


def _drain_diagnostics(evaluator):
    """Per-stage read-out of the evaluator's device-side counters (``drain_diagnostics`` of the object a bound
    ``eval_device`` belongs to); plain callables (tests, toy posteriors) have none."""
    owner = getattr(evaluator, "__self__", None)
    fn = getattr(owner, "drain_diagnostics", None) or getattr(evaluator, "drain_diagnostics", None)
    return fn() if fn is not None else {}


def _checkpoint_path(checkpoint_dir, stage):
    import os
    return os.path.join(checkpoint_dir, "stage_%d.npz" % stage)


def save_stage(checkpoint_dir, stage, q_all, like_all, logpts_all, beta, betas, acc_hist, rng, mh_states):
    """Per-stage sampler state (the reference pickles ``sample.params`` per stage, beat/sampler/smc.py:549-557,
    beat/backend.py:1043-1077): end points of all chains, beta history, host RNG state and per-rank proposal state, so
    that a run can be resumed at a stage boundary (beat/sampler/base.py:618-661) and continues bit-identically."""
    import os
    import pickle
    os.makedirs(checkpoint_dir, exist_ok=True)
    tmp = _checkpoint_path(checkpoint_dir, stage) + ".tmp.npz"
    np.savez(tmp, population=q_all, likelihoods=like_all, logpts=logpts_all, beta=np.float64(beta), betas=np.asarray(betas),
             acceptance=np.asarray(acc_hist), stage=np.int64(stage),
             rng_state=np.frombuffer(pickle.dumps(rng.bit_generator.state), dtype=np.uint8),
             mh_states=np.frombuffer(pickle.dumps(mh_states), dtype=np.uint8))
    os.replace(tmp, _checkpoint_path(checkpoint_dir, stage))           # atomic: an interrupted write never leaves a half stage


def load_last_stage(checkpoint_dir):
    """Latest complete stage file of ``checkpoint_dir`` as a dict, or None."""
    import glob
    import os
    import pickle
    import re
    best = None
    for f in glob.glob(os.path.join(checkpoint_dir, "stage_*.npz")):
        m = re.search(r"stage_(\d+)\.npz$", f)
        if m and (best is None or int(m.group(1)) > best[0]):
            best = (int(m.group(1)), f)
    if best is None:
        return None
    z = np.load(best[1])
    out = {k: z[k] for k in ("population", "likelihoods", "logpts", "betas", "acceptance")}
    out["beta"], out["stage"] = float(z["beta"]), int(z["stage"])
    out["rng_state"] = pickle.loads(z["rng_state"].tobytes())
    out["mh_states"] = pickle.loads(z["mh_states"].tobytes())
    return out


def smc_sample(evaluator, lower, upper, n_chains, n_steps, device=None, coef_variation=1.0, tune_interval=None, seed=0,
               sample_factor_final_stage=1, max_stages=200, initial_population=None, update_weights=None, log=None,
               on_step=None, checkpoint_dir=None, resume=False, cuda_graph=False):
    """Batched restatement of ``smc_sample``'s stage loop (beat/sampler/smc.py:459-546).

    Returns dict(population [n_chains, n_params], likelihoods [n_chains], logpts, betas, n_evals, acceptance).
    ``update_weights(map_point) -> None`` mirrors the ``update`` hook (smc.py:492-503): called with the MAP end
    point after each stage; the caller re-uploads weights (``BatchedFFILogLike.update_weights``) and the end points
    are re-evaluated.  ``on_step(stage, step, q, logpts, like)`` is called after every lock-step Metropolis step with
    this rank's device tensors -- the hook for a trace backend (``beat_b200.backend.BatchedNumpyChains``).
    ``cuda_graph``: capture the Metropolis step into a CUDA graph (see ``BatchedMetropolis``); ``on_step`` then receives the
    state buffers the graph updates in place -- consume them before returning.
    ``checkpoint_dir``: rank 0 writes ``stage_<k>.npz`` after every stage; ``resume=True`` continues from the latest one
    (same results as an uninterrupted run: host and per-rank device RNG states are part of the checkpoint)."""
    import torch
    from . import distributed as D
    device = device if device is not None else torch.device("cpu")
    rng = np.random.default_rng(seed)
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    n_params = lower.size
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    lo, hi = D.shard_range(n_chains, rank, world)
    n_local = hi - lo
    mh = BatchedMetropolis(evaluator, lower, upper, n_local, device=device, tune=True,
                           tune_interval=tune_interval or max(1, n_steps // 4 or 1), seed=seed * 7919 + rank, cuda_graph=cuda_graph)

    # stage 0: population from the prior (metropolis.py:128-152), identical on all ranks (shared seed)
    if initial_population is None:
        pop = rng.uniform(lower, upper, (n_chains, n_params))
    else:
        pop = np.array(initial_population, dtype=np.float64, copy=True)
    ckpt = load_last_stage(checkpoint_dir) if (resume and checkpoint_dir) else None
    if ckpt is not None:
        if ckpt["population"].shape != (n_chains, n_params):
            raise ValueError("checkpoint holds a %s population, run expects %s" % (ckpt["population"].shape, (n_chains, n_params)))
        q = torch.as_tensor(ckpt["population"][lo:hi], device=device).contiguous()
        like = torch.as_tensor(ckpt["likelihoods"][lo:hi], device=device).contiguous()
        logpts = torch.as_tensor(ckpt["logpts"][lo:hi], device=device).contiguous()
        beta, betas, stage = ckpt["beta"], list(ckpt["betas"]), ckpt["stage"]
        acc_hist = list(ckpt["acceptance"])
        rng.bit_generator.state = ckpt["rng_state"]
        st = ckpt["mh_states"][rank] if rank < len(ckpt["mh_states"]) else None
        if st is not None:
            mh.scaling.copy_(torch.as_tensor(st["scaling"], device=device))
            mh.gen.set_state(torch.as_tensor(st["gen"], dtype=torch.uint8))
            mh._n_evals += st["n_evals"]
        if log:
            log("resumed after stage %d (beta %.6f)" % (stage, beta))
    else:
        q = torch.as_tensor(pop[lo:hi], device=device).contiguous()
        logpts, like = mh.initial_llk(q)
        beta, betas, stage = 0.0, [0.0], 0
        acc_hist = []
    while beta < 1.0 and stage < max_stages:
        like_all = D.allgather_chains(like).cpu().numpy()                    # THE per-stage exchange
        q_all = D.allgather_chains(q).cpu().numpy()
        logpts_all = D.allgather_chains(logpts)
        if update_weights is not None:
            update_weights(q_all[int(np.argmax(like_all))])
            logpts, like = mh.initial_llk(q)
            like_all = D.allgather_chains(like).cpu().numpy()
            logpts_all = D.allgather_chains(logpts)
        new_beta, old_beta, weights = calc_beta(like_all, beta, coef_variation)
        final = new_beta > 1.0
        if final:                                                            # smc.py:507-513,522-524
            new_beta = 1.0
            temp = np.exp((1.0 - old_beta) * (like_all - like_all.max()))
            weights = temp / temp.sum()
        cov = calc_covariance(q_all, weights)
        idx = resample(weights, rng)                                         # identical on all ranks (shared rng)
        mh.set_proposal_covariance(cov)
        mh.beta = new_beta
        sel = torch.as_tensor(idx[lo:hi], device=device)
        q = torch.as_tensor(q_all, device=device)[sel].contiguous()
        like = torch.as_tensor(like_all, device=device)[sel].contiguous()
        logpts = logpts_all[sel].contiguous()
        draws = n_steps * (sample_factor_final_stage if final else 1)
        mh.start_stage()
        n_acc = torch.zeros((), dtype=torch.float64, device=device)
        for istep in range(draws):
            q, logpts, like, acc = mh.step(q, logpts, like)
            n_acc += acc.double().mean()                      # stays on the device; read once per stage
            if on_step is not None:
                on_step(stage + 1, istep, q, logpts, like)
        acc_hist.append(float(n_acc.item()) / max(1, draws))
        diag = _drain_diagnostics(evaluator)                   # once per stage: the device-pointer path never syncs
        if diag.get("geom_timeouts"):
            raise RuntimeError("stage %d: %d GF-store bulk copies timed out on the device (affected chains were rejected "
                               "with NaN); stopping" % (stage + 1, diag["geom_timeouts"]))
        if log and diag.get("index_violations"):
            log("stage %d: %d library / grid indices out of range (those proposals were rejected)" % (stage + 1, diag["index_violations"]))
        beta = new_beta
        betas.append(beta)
        stage += 1
        if log:
            log("stage %d beta %.6f acceptance %.3f" % (stage, beta, acc_hist[-1]))
        if checkpoint_dir:
            mine = dict(scaling=mh.scaling.cpu().numpy(), gen=mh.gen.get_state().cpu().numpy(), n_evals=mh.n_evals)
            states = D.gather_objects(mine)
            ck_like, ck_q, ck_lp = (D.allgather_chains(x).cpu().numpy() for x in (like, q, logpts))
            if rank == 0:
                save_stage(checkpoint_dir, stage, ck_q, ck_like, ck_lp, beta, betas, acc_hist, rng, states)
    like_all = D.allgather_chains(like).cpu().numpy()
    q_all = D.allgather_chains(q).cpu().numpy()
    logpts_all = D.allgather_chains(logpts).cpu().numpy()
    n_evals = torch.tensor([mh.n_evals], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(n_evals)
    return dict(population=q_all, likelihoods=like_all, logpts=logpts_all, betas=betas, n_stages=stage,
                n_evals=int(n_evals.item()), acceptance=acc_hist)


def pt_betas(n_chains, n_chains_posterior, t_scale):
    """Temperature ladder of the reference's ``TemperingManager.update_betas`` (beat/sampler/pt.py:179-221):
    ``n_chains_posterior`` chains at beta = 1, the others at beta = 1 / t_scale**k, k = 1.."""
    n_temp = n_chains - n_chains_posterior
    return np.concatenate([np.ones(n_chains_posterior), 1.0 / np.power(t_scale, np.arange(1, n_temp + 1))])


def pt_sample(evaluator, lower, upper, n_chains, n_samples, device=None, swap_interval=(10, 15), n_chains_posterior=1,
              t_scale=1.2, beta_tune_interval=None, proposal_cov=None, tune_interval=50, seed=0, initial_population=None,
              record_every=1, cuda_graph=False):
    """Lock-step parallel tempering with a batched evaluator (restating beat/sampler/pt.py:100-469,472-704 without MPI).

    The reference runs one Metropolis chain per MPI worker at its own beta, lets each run a random number of steps
    drawn from ``swap_interval`` and has the master propose a state swap between two finished workers with
    ``alpha = (beta2 - beta1) * (llk1 - llk2)`` (pt.py:428-455), counting acceptances per pair and re-scaling the
    temperature ladder every ``beta_tune_interval`` samples.  Here all chains advance together (one batched evaluation
    per step); after each interval the chains are paired at random and every pair proposes a swap with the same rule.

    Sharded over ranks (torch.distributed, one process per GPU): every rank advances ``n_chains / world`` chains.  A
    swap exchanges the two chains' TEMPERATURE LEVELS (beta and the Metropolis scaling tuned at that level) instead of
    their states -- the same Markov chain on (state, level) pairs, but no particle ever moves between GPUs.  The swap
    phase is one all-gather of (llk, scaling, acceptance counter) per chain -- 3 x 8 bytes x n_chains -- and the
    decisions are taken identically on every rank from a shared host RNG.

    Returns the recorded samples of the chains that were at beta = 1 when recorded (all ranks' records, rank order)."""
    import torch
    from . import distributed as D
    device = device if device is not None else torch.device("cpu")
    rng = np.random.default_rng(seed)                                        # shared by all ranks: identical decisions
    lower = np.asarray(lower, dtype=np.float64)
    upper = np.asarray(upper, dtype=np.float64)
    n_params = lower.size
    rank, world = 0, 1
    if torch.distributed.is_available() and torch.distributed.is_initialized():
        rank, world = torch.distributed.get_rank(), torch.distributed.get_world_size()
    lo, hi = D.shard_range(n_chains, rank, world)
    n_local = hi - lo
    mh = BatchedMetropolis(evaluator, lower, upper, n_local, device=device, tune=True, tune_interval=tune_interval,
                           seed=seed * 7919 + 1 + rank, cuda_graph=cuda_graph)
    if proposal_cov is None:
        proposal_cov = np.diag(((upper - lower) * 0.05) ** 2)
    mh.set_proposal_covariance(proposal_cov)
    ladder = pt_betas(n_chains, n_chains_posterior, t_scale)                 # beta of every temperature level
    level = np.arange(n_chains)                                              # level held by each chain (replicated)

    def set_local_betas():
        mh.beta = torch.as_tensor(ladder[level[lo:hi]], device=device)

    set_local_betas()
    pop = rng.uniform(lower, upper, (n_chains, n_params)) if initial_population is None else np.array(initial_population, dtype=np.float64)
    q = torch.as_tensor(pop[lo:hi], device=device).contiguous()
    logpts, like = mh.initial_llk(q)

    recorded, rec_like = [], []
    n_swaps = n_acc = 0
    since_tune_swaps = since_tune_acc = 0
    done = 0
    scales = [t_scale]
    while done < n_samples:
        draws = int(rng.integers(swap_interval[0], swap_interval[1]))        # pt.py:146-149 (DiscreteBoundedUniform)
        draws = min(draws, n_samples - done)
        post_local = torch.as_tensor(np.flatnonzero(level[lo:hi] < n_chains_posterior), device=device)
        for i in range(draws):
            q, logpts, like, _ = mh.step(q, logpts, like)
            if (done + i) % record_every == 0 and post_local.numel():
                recorded.append(q[post_local].clone())
                rec_like.append(like[post_local].clone())
        done += draws
        # ---- swap proposals between randomly paired chains (pt.py:428-455); THE exchange of this sampler
        packed = torch.stack([like, mh.scaling, mh.accepted], dim=1)         # [n_local, 3]
        allp = D.allgather_chains(packed).cpu().numpy()                      # [n_chains, 3] on every rank
        like_all, scal_all, accd_all = allp[:, 0].copy(), allp[:, 1].copy(), allp[:, 2].copy()
        perm = rng.permutation(n_chains)
        a, b = perm[0: 2 * (n_chains // 2): 2], perm[1: 2 * (n_chains // 2): 2]
        beta_all = ladder[level]
        alpha = (beta_all[b] - beta_all[a]) * (like_all[a] - like_all[b])
        with np.errstate(invalid="ignore"):
            acc = np.log(rng.random(a.size)) < alpha
        ia, ib = a[acc], b[acc]
        if ia.size:
            # accepted: the two chains trade places on the ladder; the step size tuned for a level stays with the level
            level[ia], level[ib] = level[ib].copy(), level[ia].copy()
            scal_all[ia], scal_all[ib] = scal_all[ib].copy(), scal_all[ia].copy()
            accd_all[ia], accd_all[ib] = accd_all[ib].copy(), accd_all[ia].copy()
            mh.scaling.copy_(torch.as_tensor(scal_all[lo:hi], device=device))
            mh.accepted.copy_(torch.as_tensor(accd_all[lo:hi], device=device))
            set_local_betas()
        k = int(acc.sum())
        n_swaps += a.size; n_acc += k
        since_tune_swaps += a.size; since_tune_acc += k
        if beta_tune_interval and since_tune_swaps >= beta_tune_interval:
            rate = since_tune_acc / float(since_tune_swaps)
            t_scale = float(np.clip(tune_pt_scale(t_scale, rate), 1.01, 2.0))            # pt.py:344-348 (limits :126-127)
            ladder = pt_betas(n_chains, n_chains_posterior, t_scale)
            set_local_betas()
            scales.append(t_scale)
            since_tune_swaps = since_tune_acc = 0
    if recorded:
        samples_local = torch.cat(recorded).reshape(-1, n_params).cpu().numpy()
        like_local = torch.cat(rec_like).reshape(-1).cpu().numpy()
    else:
        samples_local, like_local = np.zeros((0, n_params)), np.zeros(0)
    parts = D.gather_objects((samples_local, like_local))                     # ragged per rank: once, at the end
    samples = np.concatenate([p[0] for p in parts])
    likes = np.concatenate([p[1] for p in parts])
    n_evals = torch.tensor([mh.n_evals], dtype=torch.float64, device=device)
    if world > 1:
        torch.distributed.all_reduce(n_evals)
    return dict(samples=samples, likelihoods=likes, betas=ladder, chain_betas=ladder[level], levels=level.copy(),
                swap_acceptance=n_acc / max(1, n_swaps), t_scales=scales, n_evals=int(n_evals.item()),
                population=D.allgather_chains(q).cpu().numpy(), population_likelihoods=D.allgather_chains(like).cpu().numpy())
