"""Randomised shape / option sweep of the fused evaluation against the oracle (GPU): ragged sample counts, patch
counts around the chunk and plan-chunk boundaries, 1-3 slip components, both interpolations and storage types,
both execution modes.  Seeds are fixed; every case is small enough for the numpy oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from beat_b200 import synthetic  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402


def _case(i):
    rng = np.random.default_rng(1000 + i)
    nsf = int(rng.integers(1, 3))
    subfaults = tuple((int(rng.integers(1, 9)), int(rng.integers(1, 12)), float(rng.choice([1.0, 2.0, 2.5]))) for _ in range(nsf))
    return dict(
        nt=int(rng.integers(1, 6)), subfaults=subfaults, ns=int(rng.choice([1, 3, 8, 30, 33, 64, 127, 128, 129, 140])),
        ndur=int(rng.integers(2, 6)), slip_vars=("uparr", "uperp", "utens")[: int(rng.integers(1, 4))],
        interpolation=str(rng.choice(["multilinear", "nearest_neighbor"])),
        noise=str(rng.choice(["variance", "exponential", "dense"])), station_corrections=bool(rng.integers(0, 2)),
        hp_specific=bool(rng.integers(0, 2)), n_wavemaps=int(rng.integers(1, 3)), seed=2000 + i)


@pytest.mark.parametrize("i", range(24))
def test_random_configuration(i, monkeypatch):
    from beat_b200.engine import BatchedFFILogLike
    args = _case(i)
    if args["ns"] < 5 and args["noise"] != "variance":
        args["noise"] = "variance"
    monkeypatch.setenv("BEATGPU_STACK_MODE", "chunked" if i % 2 else "fused")
    monkeypatch.setenv("BEATGPU_CHUNK", str([5, 32, 9, 16][i % 4]))
    prob = synthetic.make_problem(**args)
    B = [1, 3, 17][i % 3]
    Q = synthetic.draw_chains(prob, B, seed=i)
    ref = np.array([O.ffi_seismic_eval(prob, synthetic.split_point(prob, q), impl="port") for q in Q])
    for store, rtol in (("float64", 1e-9), ("float32", 2e-5)):
        ev = BatchedFFILogLike.from_problem(prob, store_dtype=store)
        logpts, like = ev(Q)
        np.testing.assert_allclose(logpts, ref, rtol=rtol, err_msg=str(args))
        # `like` can cancel to ~0 (log-densities of either sign): tolerance relative to the magnitude of its terms
        np.testing.assert_allclose(like, ref.sum(axis=1), rtol=rtol, atol=rtol * np.abs(ref).sum(axis=1).max())
        ev.close()
