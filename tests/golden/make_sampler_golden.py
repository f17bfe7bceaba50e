"""
tests/golden/make_sampler_golden.py -- regenerates tests/golden/sampler_golden.npz.

Stage functions of the reference's samplers, imported from /root/reference in the dev container (pymc / pytensor /
pyrocko replaced by the inert stand-ins of tests/golden/_refshim.py -- none of them is touched by these functions):

    beat.sampler.smc.SMC.calc_beta          (smc.py:133-165)   bisection for the next tempering beta + importance weights
    beat.sampler.smc.SMC.calc_covariance    (smc.py:167-186)   weighted population covariance + utility.ensure_cov_psd
    beat.sampler.smc.SMC.resample           (smc.py:290-324)   Kitagawa's deterministic resampling
    beat.sampler.pt.tune                    (pt.py:37-73)      temperature-scale adaptation of the PT ladder

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

    np.savez_compressed(os.path.join(HERE, "sampler_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "sampler_golden.npz"), {k: np.asarray(v).shape for k, v in out.items() if k.endswith(("beta", "out", "idx", "scale"))})


if __name__ == "__main__":
    main()
