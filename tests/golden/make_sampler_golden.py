"""
tests/golden/make_sampler_golden.py -- regenerates tests/golden/sampler_golden.npz.

Stage functions of the reference's samplers, imported from /root/reference in the dev container (pymc / pytensor /
pyrocko replaced by the inert stand-ins of tests/golden/_refshim.py -- none of them is touched by these functions):

    beat.sampler.smc.SMC.calc_beta          (smc.py:133-165)   bisection for the next tempering beta + importance weights
    beat.sampler.smc.SMC.calc_covariance    (smc.py:167-186)   weighted population covariance + utility.ensure_cov_psd
    beat.sampler.smc.SMC.resample           (smc.py:290-324)   Kitagawa's deterministic resampling
    beat.sampler.pt.tune                    (pt.py:37-73)      temperature-scale adaptation of the PT ladder
    beat.sampler.pt.TemperingManager.update_betas / propose_chain_swap   (pt.py:179-214, 428-455)   ladder and swap decision

run on seeded inputs; the vectors pin beat_b200.sampler.calc_beta / calc_covariance / resample / tune_pt_scale (row f1).

    python tests/golden/make_sampler_golden.py
"""
import os
import sys
import types
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _refshim  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402

warnings.simplefilter("ignore")


def main():
    _refshim.install(fast_sweep_ext=O.load_reference_ext())
    from beat.sampler import pt as rpt
    from beat.sampler import smc as rsmc
    rng = np.random.default_rng(2026)
    out = {}

    # ---- calc_beta: broad, narrow and near-degenerate likelihood populations, first and later stages
    cases = [(rng.normal(-500.0, 40.0, 300), 0.0, 1.0), (rng.normal(-50.0, 2.0, 128), 0.3, 1.0),
             (rng.normal(-1e4, 900.0, 1000), 0.0, 0.6), (np.r_[rng.normal(-20.0, 0.5, 63), -5.0], 0.8, 1.0)]
    for i, (lk, beta, cv) in enumerate(cases):
        b, old, w = rsmc.SMC.calc_beta(types.SimpleNamespace(beta=beta, likelihoods=lk, coef_variation=cv))
        out["cb%d_like" % i], out["cb%d_in" % i] = lk, np.array([beta, cv])
        out["cb%d_beta" % i], out["cb%d_weights" % i] = np.array([b, old]), w
    out["cb_n"] = np.int64(len(cases))

    # ---- calc_covariance: well-conditioned, and rank-deficient (fewer effective samples than dimensions -> repaired)
    pops = [(rng.standard_normal((200, 6)) * np.arange(1, 7), rng.random(200)),
            (np.outer(rng.standard_normal(40), np.ones(5)) + 1e-9 * rng.standard_normal((40, 5)), rng.random(40)),
            (rng.standard_normal((12, 3)), np.r_[1.0, np.zeros(11)] + 1e-3)]
    for i, (pop, w) in enumerate(pops):
        w = w / w.sum()
        out["cov%d_pop" % i], out["cov%d_w" % i] = pop, w
        out["cov%d_out" % i] = rsmc.SMC.calc_covariance(types.SimpleNamespace(array_population=pop, weights=w))
    out["cov_n"] = np.int64(len(pops))

    # ---- resample: the reference draws its offset from numpy's global RNG; the drawn value is stored with the result
    for i, n in enumerate((7, 64, 500)):
        w = rng.random(n) ** 3
        w /= w.sum()
        np.random.seed(100 + i)
        aux = np.random.rand(1)[0]
        np.random.seed(100 + i)
        idx = rsmc.SMC.resample(types.SimpleNamespace(n_chains=n, weights=w))
        out["rs%d_w" % i], out["rs%d_aux" % i], out["rs%d_idx" % i] = w, np.float64(aux), idx
    out["rs_n"] = np.int64(3)

    # ---- PT temperature-scale tuning
    acc = np.array([0.0, 0.0005, 0.001, 0.02, 0.05, 0.1, 0.2, 0.35, 0.5, 0.6, 0.75, 0.8, 0.95, 0.99, 1.0])
    out["pt_acc"] = acc
    out["pt_scale"] = np.array([rpt.tune(1.3, float(a)) for a in acc])

    # ---- PT ladder (TemperingManager.update_betas, pt.py:179-214) and the swap decision (propose_chain_swap, :428-455)
    for i, (n, n_post, t_scale) in enumerate(((16, 4, 1.6), (512, 64, 1.2), (5, 1, 2.0))):
        tm = types.SimpleNamespace(n_workers_posterior=n_post, n_workers_tempered=n - n_post, current_scale=None, _betas=None,
                                   _worker_package_mapping={})
        rpt.TemperingManager.update_betas(tm, t_scale)
        out["ladder%d_in" % i], out["ladder%d_betas" % i] = np.array([n, n_post, t_scale]), np.asarray(tm._betas, dtype=np.float64)
    out["ladder_n"] = np.int64(3)
    n_sw = 200
    b1, b2 = rng.uniform(0.05, 1.0, n_sw), rng.uniform(0.05, 1.0, n_sw)
    l1, l2 = rng.normal(-300.0, 30.0, n_sw), rng.normal(-300.0, 30.0, n_sw)
    np.random.seed(77)
    u = np.random.uniform(size=n_sw)
    np.random.seed(77)
    acc = np.zeros(n_sw, dtype=bool)
    for k in range(n_sw):
        steps = {1: types.SimpleNamespace(beta=b1[k], _llk_index=0), 2: types.SimpleNamespace(beta=b2[k], _llk_index=0)}
        reg = {}
        tm = types.SimpleNamespace(worker_a2l=lambda m, source: [m], worker2package=lambda source: {"step": steps[source]},
                                   register_swap=lambda s1, s2, accepted: reg.update(acc=accepted))
        rpt.TemperingManager.propose_chain_swap(tm, l1[k], l2[k], 1, 2)
        acc[k] = reg["acc"]
    out["swap_b1"], out["swap_b2"], out["swap_l1"], out["swap_l2"], out["swap_u"], out["swap_acc"] = b1, b2, l1, l2, u, acc

    np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "sampler_golden.npz"), {k: np.asarray(v).shape for k, v in out.items() if k.endswith(("beta", "out", "idx", "scale"))})


if __name__ == "__main__":
    main()
