"""GPU integration test of the lock-step SMC driver (row f1) on a small FFI problem with the real batched evaluator."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from beat_b200 import synthetic  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402


def test_smc_ffi_small_posterior_concentrates():
    import torch
    from beat_b200 import sampler as S
    from beat_b200.covariance import Covariance, exponential_data_covariance
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=8, subfaults=((3, 4, 4.0),), ns=48, ndur=4, seed=5, slip_vars=("uparr",))
    q_true = synthetic.draw_chains(prob, 1, seed=6)[0]
    q_true[prob["offsets"]["hypers"]] = 0.0
    _, synths, _ = O.ffi_seismic_eval(prob, synthetic.split_point(prob, q_true), impl="port", return_synth=True)
    rng = np.random.default_rng(1)
    wm = prob["wavemaps"][0]
    sigma = 0.1 * np.abs(synths[0]).max(axis=1)
    for t in range(wm["nt"]):
        C = exponential_data_covariance(wm["ns"], 0.5, 2.0) * sigma[t] ** 2
        cov = Covariance(data=C)
        wm["U"][t], wm["slog_pdet"][t] = cov.chol_inverse, cov.log_pdet
        wm["data"][t] = synths[0][t] + np.linalg.cholesky(C).dot(rng.standard_normal(wm["ns"]))
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    dev = torch.device("cuda", 0)
    lower = np.concatenate([prob["priors"][n][0] for n, _ in prob["var_order"]])
    upper = np.concatenate([prob["priors"][n][1] for n, _ in prob["var_order"]])
    n_chains = 512
    prior_like = ev(synthetic.draw_chains(prob, n_chains, seed=2))[1]
    true_like = ev(q_true[None, :])[1][0]
    out = S.smc_sample(ev.eval_device, lower, upper, n_chains=n_chains, n_steps=25, device=dev, seed=4, max_stages=60)
    assert out["betas"][-1] == 1.0
    post = out["likelihoods"]
    # the population moved from the prior towards the data-generating model
    assert np.median(post) > np.median(prior_like) + 50.0
    assert np.median(post) > true_like - 0.6 * (true_like - np.median(prior_like))
    # stored llk of the final population equals a fresh evaluation (accept/reject bookkeeping is consistent) ...
    fresh_logpts, fresh = ev(out["population"])
    np.testing.assert_allclose(post, fresh, rtol=1e-12)
    np.testing.assert_allclose(out["logpts"], fresh_logpts, rtol=1e-12)
    # ... and the oracle's value for a few of them
    for c in (0, 100, 511):
        ref = O.ffi_seismic_eval(prob, synthetic.split_point(prob, out["population"][c]), impl="port").sum()
        assert abs(post[c] - ref) <= 1e-9 * abs(ref)
    # all end points inside the prior box (the bounds check of metropolis.py:341-343)
    assert (out["population"] >= lower).all() and (out["population"] <= upper).all()
    assert out["n_evals"] > n_chains
    ev.close()
