"""GPU parity at the full sizes of BASELINE.json's configs 4 and 5 (the C3 headline has its own file):
C5 = 2 subfaults x 150 patches, 64 targets x 120 samples, library 17 x 64 (20 GB f32 in HBM): every logpt of 256 chains
against the oracle (fanned out over the host cores); C4 = C3 seismic + geodetic static (500 observations, dense non-Toeplitz
covariance) + laplacian prior: the geodetic and laplacian terms of all 2000 chains (DMMA GEMM paths) against the oracle.
Tolerances: f32 library rtol 1e-5 (north star), f64 terms rtol 1e-10."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from beat_b200 import synthetic  # noqa: E402
from oracle import ffi_oracle as O  # noqa: E402

C5 = dict(nt=64, subfaults=((10, 15, 2.0), (10, 15, 2.0)), ns=120, ndur=17, nst=64)
C4 = dict(nt=64, subfaults=((10, 20, 2.0),), ns=120, ndur=17, nst=64, geodetic=dict(nobs=[500]), laplacian=True)


def test_c5_two_subfaults_all_logpts(tmp_path):
    import torch
    from oracle import parallel_check as PC
    from beat_b200.devlib import fill_library_on_device
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(interpolation="multilinear", seed=4321, build_library=False, **C5)
    ev = BatchedFFILogLike.from_problem(prob, device=0, store_dtype="float32", upload_libraries=False)
    try:
        dev = torch.device("cuda", 0)
        fill_library_on_device(ev, prob, torch, dev, "f32")
        B = 256
        Q = synthetic.draw_chains(prob, B, seed=77)
        logpts, like = ev(Q)
        with PC.pool(workers=min(PC.n_workers(), 12)) as ex:            # 300 patches: ~2.5 GB of numpy temporaries per worker
            ref = PC.full_size_logpts(prob, Q, ex)
        assert logpts.shape == ref.shape == (B, 64) and np.isfinite(ref).all()
        np.testing.assert_allclose(logpts, ref, rtol=1e-5)
        np.testing.assert_allclose(like, ref.sum(axis=1), rtol=1e-5)
        # each subfault's rupture front starts at its own nucleation point and time: bit-equal to the sequential C restatement
        st = ev.starttimes(B)
        for c in (0, 100, 255):
            pt = synthetic.split_point(prob, Q[c])
            for sf, (nd, ns_, h) in enumerate(prob["subfaults"]):
                hr, hc = O.fault_locations2idxs(pt["nucleation_dip"][sf], pt["nucleation_strike"][sf], h, h)
                t0 = O.fast_sweep(1.0 / pt["velocities"][sf * 150:(sf + 1) * 150], h, hr, hc, nd, ns_, impl="port") + pt["time"][sf]
                assert np.array_equal(st[c, sf * 150:(sf + 1) * 150], t0)
    finally:
        ev.close()


def test_c4_geodetic_and_laplacian_terms_all_chains():
    import torch
    from beat_b200.devlib import fill_library_on_device
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(interpolation="multilinear", seed=1234, build_library=False, **C4)
    ev = BatchedFFILogLike.from_problem(prob, device=0, store_dtype="float32", upload_libraries=False)
    try:
        fill_library_on_device(ev, prob, torch, torch.device("cuda", 0), "f32")
        B = 2000
        Q = synthetic.draw_chains(prob, B, seed=78)
        logpts, like = ev(Q)
        assert logpts.shape == (B, 64 + 1 + 1) and np.isfinite(logpts).all()
        geo = np.array([O.ffi_geodetic_eval(prob["geodetic"], synthetic.split_point(prob, q))[0] for q in Q])
        lap = np.array([O.ffi_laplacian_eval(prob["laplacian"], synthetic.split_point(prob, q), prob["slip_vars"]) for q in Q])
        np.testing.assert_allclose(logpts[:, 64], geo, rtol=1e-10)
        np.testing.assert_allclose(logpts[:, 65], lap, rtol=1e-10)
        np.testing.assert_allclose(like, logpts.sum(axis=1), rtol=1e-13)
    finally:
        ev.close()
