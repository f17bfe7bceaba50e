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


def test_smc_ffi_with_trace_writer_on_the_gpu(tmp_path):
    """The sampler run a user would do: lock-step SMC on the batched CUDA evaluator with every step of every chain
    streamed into the reference's binary trace format (row f4 driven by row f1).  The files hold n_steps records per
    stage; the last record of each chain is its end point, and re-evaluating a stored point reproduces its stored logpts."""
    import torch
    from collections import OrderedDict
    from beat_b200 import backend as bk
    from beat_b200 import sampler as S
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=5, subfaults=((3, 4, 4.0),), ns=40, ndur=4, seed=15)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    dev = torch.device("cuda", 0)
    lower = np.concatenate([prob["priors"][n][0] for n, _ in prob["var_order"]])
    upper = np.concatenate([prob["priors"][n][1] for n, _ in prob["var_order"]])
    n_chains, n_steps = 64, 5
    shapes = OrderedDict((name, (int(n),)) for name, n in prob["var_order"])
    shapes["seis_like"] = (ev.n_out,)
    shapes["like"] = ()
    off = prob["offsets"]
    writers = {}

    def on_step(stage, step, q, logpts, like):
        if stage not in writers:
            writers[stage] = bk.BatchedNumpyChains(str(tmp_path / ("stage_%d" % stage)), shapes, n_chains, buffer_size=3)
            writers[stage].setup()
        qn = q.cpu().numpy()
        vals = {name: qn[:, off[name]:off[name] + n] for name, n in prob["var_order"]}
        vals["seis_like"], vals["like"] = logpts.cpu().numpy(), like.cpu().numpy()
        writers[stage].write(vals)

    out = S.smc_sample(ev.eval_device, lower, upper, n_chains=n_chains, n_steps=n_steps, device=dev, seed=2, max_stages=4, on_step=on_step)
    for w in writers.values():
        w.flush()
    assert len(writers) == out["n_stages"]
    last = writers[max(writers)]
    for c in (0, 17, 63):
        rec, _ = bk.read_chain(last.filename(c))
        assert rec.shape[0] == n_steps
        q_last = np.concatenate([rec[name][-1].ravel() for name, _ in prob["var_order"]])
        np.testing.assert_array_equal(q_last, out["population"][c])
        np.testing.assert_array_equal(rec["like"][-1], out["likelihoods"][c])
        lp, lk = ev(q_last[None, :])
        np.testing.assert_allclose(rec["seis_like"][-1].ravel(), lp[0], rtol=1e-12)
        ref = O.ffi_seismic_eval(prob, synthetic.split_point(prob, q_last), impl="port")
        np.testing.assert_allclose(rec["seis_like"][-1].ravel(), ref, rtol=1e-9)
    assert ev.drain_diagnostics()["index_violations"] == 0
    ev.close()


def test_device_recorder_streams_the_same_files_as_the_host_path(tmp_path):
    """DeviceRecorder (records packed on the device, copied into the writer's page-locked buffers on a side stream, appended
    by writer threads) against the synchronous dict-fed writer on the same SMC run: byte-identical chain files."""
    from collections import OrderedDict
    import torch
    from beat_b200 import backend as bk
    from beat_b200 import sampler as S
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=4, subfaults=((3, 4, 4.0),), ns=32, ndur=4, seed=16)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    dev = torch.device("cuda", 0)
    lower = np.concatenate([prob["priors"][n][0] for n, _ in prob["var_order"]])
    upper = np.concatenate([prob["priors"][n][1] for n, _ in prob["var_order"]])
    n_chains, n_steps = 48, 7
    shapes = OrderedDict((name, (int(n),)) for name, n in prob["var_order"])
    shapes["seis_like"] = (ev.n_out,)
    shapes["like"] = ()
    off = prob["offsets"]
    host_w, dev_w, recs = {}, {}, {}

    def on_step(stage, step, q, logpts, like):
        if stage not in host_w:
            host_w[stage] = bk.BatchedNumpyChains(str(tmp_path / ("host_%d" % stage)), shapes, n_chains, buffer_size=100)
            dev_w[stage] = bk.BatchedNumpyChains(str(tmp_path / ("dev_%d" % stage)), shapes, n_chains, buffer_size=3, n_io_threads=2, pinned=True)
            host_w[stage].setup(); dev_w[stage].setup()
            recs[stage] = bk.DeviceRecorder(dev_w[stage], torch, dev)
        recs[stage].record(q, logpts, like)                       # q's columns are in var_order: the record is q | logpts | like
        qn = q.cpu().numpy()
        vals = {name: qn[:, off[name]:off[name] + n] for name, n in prob["var_order"]}
        vals["seis_like"], vals["like"] = logpts.cpu().numpy(), like.cpu().numpy()
        host_w[stage].write(vals)

    out = S.smc_sample(ev.eval_device, lower, upper, n_chains=n_chains, n_steps=n_steps, device=dev, seed=5, max_stages=3, on_step=on_step)
    assert out["n_stages"] >= 2
    for st in host_w:
        host_w[st].flush()
        recs[st].finish()
        dev_w[st].close()
        assert dev_w[st].stored_samples == host_w[st].stored_samples == n_steps
        for c in range(n_chains):
            assert open(dev_w[st].filename(c), "rb").read() == open(host_w[st].filename(c), "rb").read(), (st, c)
    with pytest.raises(ValueError):
        recs[max(recs)].record(torch.zeros((n_chains, 3), dtype=torch.float64, device=dev))
    ev.close()


def test_metropolis_step_as_one_cuda_graph():
    """cuda_graph=True: proposal, bounds check, the batched evaluation (libbeatgpu's kernels captured on torch's stream),
    accept / reject and the in-place state update replay as ONE CUDA graph.  The run is deterministic, its bookkeeping
    is consistent (stored llk == fresh evaluation of the stored point, oracle included) and it samples the same
    posterior as the eager path."""
    import torch
    from beat_b200 import sampler as S
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=6, subfaults=((3, 5, 3.0),), ns=40, ndur=4, seed=31)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    dev = torch.device("cuda", 0)
    lower = np.concatenate([prob["priors"][n][0] for n, _ in prob["var_order"]])
    upper = np.concatenate([prob["priors"][n][1] for n, _ in prob["var_order"]])
    kw = dict(n_chains=256, n_steps=12, device=dev, seed=9, max_stages=5)
    launches0 = ev.ctx.launch_count()
    g1 = S.smc_sample(ev.eval_device, lower, upper, cuda_graph=True, **kw)
    launched_graph = ev.ctx.launch_count() - launches0
    g2 = S.smc_sample(ev.eval_device, lower, upper, cuda_graph=True, **kw)
    launches1 = ev.ctx.launch_count()
    eager = S.smc_sample(ev.eval_device, lower, upper, cuda_graph=False, **kw)
    launched_eager = ev.ctx.launch_count() - launches1
    np.testing.assert_array_equal(g1["population"], g2["population"])           # deterministic
    np.testing.assert_array_equal(g1["likelihoods"], g2["likelihoods"])
    assert g1["n_stages"] == eager["n_stages"] and np.isfinite(g1["likelihoods"]).all()
    fresh_lp, fresh = ev(g1["population"])
    np.testing.assert_allclose(g1["likelihoods"], fresh, rtol=1e-12)
    np.testing.assert_allclose(g1["logpts"], fresh_lp, rtol=1e-12)
    for c in (0, 255):
        ref = O.ffi_seismic_eval(prob, synthetic.split_point(prob, g1["population"][c]), impl="port").sum()
        assert abs(g1["likelihoods"][c] - ref) <= 1e-9 * abs(ref)
    assert (g1["population"] >= lower).all() and (g1["population"] <= upper).all()
    # same sampler, same posterior: the tempered populations agree in their llk statistics
    assert abs(np.median(g1["likelihoods"]) - np.median(eager["likelihoods"])) < 0.25 * np.std(eager["likelihoods"]) + 5.0
    assert g1["n_evals"] > 256 and 0.0 < np.mean(g1["acceptance"]) < 1.0
    # replays do not go through the library's host entry: a fraction of the eager run's host-side launches
    assert launched_graph < 0.5 * launched_eager, (launched_graph, launched_eager)
    assert ev.drain_diagnostics()["index_violations"] >= 0
    ev.close()
