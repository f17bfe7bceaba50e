"""CPU tests of the lock-step SMC driver shim (row f1) on the reference's own toy posterior
(test/test_smc.py:38-104: 4-d mixture of two Gaussians, Uniform(-2, 2) prior; |x| mean recovered to atol 0.03),
plus unit checks of the stage functions against straightforward restatements of the reference loops."""
import numpy as np
import pytest
import torch

from beat_b200 import sampler as S


def _two_gaussians_evaluator(n=4, stdev=0.1):
    mu1 = torch.ones(n, dtype=torch.float64) * 0.5
    mu2 = -mu1
    isig = 1.0 / stdev ** 2
    logdet = n * np.log(stdev ** 2)
    w1, w2 = stdev, 1 - stdev

    def ev(q):
        l1 = -0.5 * n * np.log(2 * np.pi) - 0.5 * logdet - 0.5 * isig * ((q - mu1) ** 2).sum(dim=1)
        l2 = -0.5 * n * np.log(2 * np.pi) - 0.5 * logdet - 0.5 * isig * ((q - mu2) ** 2).sum(dim=1)
        like = torch.logsumexp(torch.stack([np.log(w1) + l1, np.log(w2) + l2]), dim=0)
        return like[:, None].clone(), like
    return ev, mu1.numpy()


def test_smc_two_gaussians_like_reference_test():
    ev, mu1 = _two_gaussians_evaluator()
    n = 4
    out = S.smc_sample(ev, -2.0 * np.ones(n), 2.0 * np.ones(n), n_chains=1000, n_steps=100, tune_interval=25, seed=3)
    assert out["betas"][-1] == 1.0 and out["n_stages"] >= 3
    x = out["population"]
    np.testing.assert_allclose(np.abs(x).mean(axis=0), mu1, rtol=0.0, atol=0.03)
    # both modes are populated roughly 10 % / 90 % (w1 = 0.1)
    frac_pos = (x[:, 0] > 0).mean()
    assert 0.02 < frac_pos < 0.25
    assert out["n_evals"] > 0


def test_resample_equals_reference_loop():
    rng = np.random.default_rng(0)
    for n in (5, 64, 1000):
        w = rng.random(n) ** 3
        w /= w.sum()

        class R:                      # same u for both implementations
            def __init__(self, v): self.v = v
            def random(self): return self.v
        aux = rng.random()
        got = S.resample(w, R(aux))
        # the reference's loops, beat/sampler/smc.py:300-324
        parents = np.arange(n)
        N_childs = np.zeros(n, dtype=int)
        cum_dist = np.cumsum(w)
        u = (parents + aux) / n
        j = 0
        for i in parents:
            while u[i] > cum_dist[j] and j < n - 1:
                j += 1
            N_childs[j] += 1
        ref = np.repeat(parents, N_childs)
        assert np.array_equal(got, ref)


def test_calc_beta_and_covariance():
    rng = np.random.default_rng(1)
    like = rng.normal(-500, 30, 400)
    beta, old, w = S.calc_beta(like, 0.0, coef_variation=1.0)
    assert 0 < beta < 1 and old == 0.0 and abs(w.sum() - 1) < 1e-12
    temp = np.exp((beta - 0.0) * (like - like.max()))
    assert abs(np.std(temp) / np.mean(temp) - 1.0) < 1e-2            # bisection target: COV == coef_variation
    pop = rng.standard_normal((400, 6))
    cov = S.calc_covariance(pop, w)
    np.testing.assert_allclose(cov, np.cov(pop, aweights=w, bias=False, rowvar=0))
    np.linalg.cholesky(cov)


def test_tune_scale_table():
    s = torch.ones(6, dtype=torch.float64)
    acc = torch.tensor([0.0005, 0.03, 0.1, 0.3, 0.6, 0.99], dtype=torch.float64)
    np.testing.assert_allclose(S.tune_scale(s, acc).numpy(), [0.1, 0.5, 0.9, 1.0, 1.1, 10.0])


def test_pt_two_gaussians_like_reference_test():
    """Lock-step parallel tempering on the toy posterior of test/test_pt.py (same two-Gaussian mixture): the beta = 1
    chains visit both modes in the right proportion (w1 = 0.1) and recover |x| ~ 0.5."""
    ev, mu1 = _two_gaussians_evaluator()
    n = 4
    out = S.pt_sample(ev, -2.0 * np.ones(n), 2.0 * np.ones(n), n_chains=16, n_samples=6000, swap_interval=(10, 15),
                      n_chains_posterior=4, t_scale=1.6, beta_tune_interval=200, seed=5)
    x = out["samples"][len(out["samples"]) // 5:]                      # drop burn-in
    np.testing.assert_allclose(np.abs(x).mean(axis=0), mu1, rtol=0.0, atol=0.05)
    frac_pos = (x[:, 0] > 0).mean()
    assert 0.02 < frac_pos < 0.35, frac_pos
    assert 0.0 < out["swap_acceptance"] <= 1.0
    assert out["betas"][0] == 1.0 and np.all(np.diff(out["betas"][3:]) < 0)
    np.testing.assert_allclose(S.pt_betas(5, 2, 2.0), [1, 1, 0.5, 0.25, 0.125])


def test_smc_checkpoint_resume_is_bit_identical(tmp_path):
    """Stage checkpoints (beat/sampler/smc.py:549-557, backend.py:1043-1077 write ``sample.params`` per stage): a run
    interrupted after k stages and resumed (sampler/base.py:618-661) ends exactly where the uninterrupted run ends."""
    import torch
    from beat_b200 import sampler as S
    n = 4
    mu1, mu2 = torch.full((n,), 0.5, dtype=torch.float64), torch.full((n,), -0.5, dtype=torch.float64)

    def evaluator(q):
        a = -0.5 * ((q - mu1) ** 2).sum(1) / 0.01
        b = -0.5 * ((q - mu2) ** 2).sum(1) / 0.01
        like = torch.logsumexp(torch.stack([a + np.log(0.1), b + np.log(0.9)]), 0)
        return like[:, None].clone(), like

    lower, upper = -2.0 * np.ones(n), 2.0 * np.ones(n)
    kw = dict(n_chains=200, n_steps=10, seed=11)
    full = S.smc_sample(evaluator, lower, upper, checkpoint_dir=str(tmp_path / "a"), **kw)
    assert full["n_stages"] >= 4
    part = S.smc_sample(evaluator, lower, upper, checkpoint_dir=str(tmp_path / "b"), max_stages=2, **kw)
    assert part["n_stages"] == 2 and part["betas"][-1] < 1.0
    ck = S.load_last_stage(str(tmp_path / "b"))
    assert ck["stage"] == 2 and ck["population"].shape == (200, n)
    np.testing.assert_array_equal(ck["population"], part["population"])
    resumed = S.smc_sample(evaluator, lower, upper, checkpoint_dir=str(tmp_path / "b"), resume=True, **kw)
    assert resumed["betas"] == full["betas"] and resumed["n_stages"] == full["n_stages"]
    np.testing.assert_array_equal(resumed["population"], full["population"])
    np.testing.assert_array_equal(resumed["likelihoods"], full["likelihoods"])
    assert resumed["n_evals"] == full["n_evals"]
    with pytest.raises(ValueError, match="checkpoint holds"):
        S.smc_sample(evaluator, lower, upper, n_chains=100, n_steps=10, seed=11, checkpoint_dir=str(tmp_path / "b"), resume=True)


def _sampler_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampler_golden.npz"))


def test_stage_functions_match_the_references_own():
    """tests/golden/sampler_golden.npz: outputs of the reference's OWN SMC.calc_beta / calc_covariance / resample
    (beat/sampler/smc.py:133-186,290-324) and pt.tune (beat/sampler/pt.py:37-73), generated by importing them from
    /root/reference (tests/golden/make_sampler_golden.py).  Integer work (resampling indexes) is exact; the bisection is
    the same sequence of floating-point operations, so beta and weights are bit-equal; the covariance to rounding."""
    g = _sampler_golden()
    for i in range(int(g["cb_n"])):
        beta_in, cv = g["cb%d_in" % i]
        b, old, w = S.calc_beta(g["cb%d_like" % i], float(beta_in), coef_variation=float(cv))
        assert np.array_equal(np.array([b, old]), g["cb%d_beta" % i])
        assert np.array_equal(w, g["cb%d_weights" % i])
    for i in range(int(g["cov_n"])):
        cov = S.calc_covariance(g["cov%d_pop" % i], g["cov%d_w" % i])
        ref = g["cov%d_out" % i]
        np.testing.assert_allclose(cov, ref, rtol=1e-12, atol=1e-14 * np.abs(ref).max())
        assert np.all(np.linalg.eigvalsh((cov + cov.T) / 2.0) >= -1e-12 * np.abs(ref).max())

    class FixedOffset:                                    # the reference draws np.random.rand(1); the fixture stores that draw
        def __init__(self, u):
            self.u = u

        def random(self):
            return self.u
    for i in range(int(g["rs_n"])):
        idx = S.resample(g["rs%d_w" % i], FixedOffset(float(g["rs%d_aux" % i])))
        assert np.array_equal(idx, g["rs%d_idx" % i])
    got = np.array([S.tune_pt_scale(1.3, float(a)) for a in g["pt_acc"]])
    assert np.array_equal(got, g["pt_scale"])


def test_pt_ladder_and_swap_rule_match_the_references_own():
    """The temperature ladder of TemperingManager.update_betas (beat/sampler/pt.py:179-214) and the swap decision of
    propose_chain_swap (:428-455: alpha = (beta2 - beta1) * (llk1 - llk2), accepted when log(u) < alpha), both run from
    /root/reference (tests/golden/make_sampler_golden.py): pt_betas is bit-equal, and the vectorised decision pt_sample
    takes for a pair of chains equals the reference's for the same betas, likelihoods and uniform draws."""
    g = _sampler_golden()
    for i in range(int(g["ladder_n"])):
        n, n_post, t_scale = g["ladder%d_in" % i]
        assert np.array_equal(S.pt_betas(int(n), int(n_post), float(t_scale)), g["ladder%d_betas" % i])
    # pt_sample: alpha = (beta_all[b] - beta_all[a]) * (like_all[a] - like_all[b]); acc = log(u) < alpha   (a = chain 1, b = chain 2)
    alpha = (g["swap_b2"] - g["swap_b1"]) * (g["swap_l1"] - g["swap_l2"])
    with np.errstate(invalid="ignore"):
        acc = np.log(g["swap_u"]) < alpha
    assert np.array_equal(acc, g["swap_acc"]) and 0 < acc.sum() < acc.size
