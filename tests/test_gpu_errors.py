"""GPU tests of the C-ABI's error behaviour: call-order and argument errors are reported, never repaired."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_call_order_and_argument_errors():
    from beat_b200.lib import BeatGpuError, Context, Layout, F64
    c = Context(0)
    with pytest.raises(BeatGpuError, match="set_fault"):
        c.fast_sweep_batch(0, np.ones((1, 4)), [0], [0])
    with pytest.raises(ValueError):
        c.set_fault([0], [3], [1.0])                       # empty grid
    c.set_fault([2], [2], [1.0])
    L = Layout()
    L.n_params, L.n_slipvars = 4, 1
    L.off_slip[0] = 0
    L.off_durations = L.off_velocities = L.off_nucleation_strike = L.off_nucleation_dip = L.off_time = -1
    L.off_hypers, L.n_hypers, L.off_time_shifts, L.n_time_shifts = -1, 1, -1, 0
    with pytest.raises(ValueError, match="fixed"):
        c.set_layout(L, None)                              # fixed variables without the fixed vector
    L.off_slip[0] = 2
    with pytest.raises(ValueError, match="exceeds"):
        c.set_layout(L, np.zeros(4 + 4 + 4 + 3 + 1))
    L.off_slip[0] = 0
    c.set_layout(L, np.concatenate([np.zeros(4), np.full(4, 1.0), np.full(4, 3.0), [0.5], [0.5], [0.0], [0.0]]))
    with pytest.raises(BeatGpuError, match="no composite"):
        c.ffi_loglike_batch(np.ones((2, 4)))
    wid = c.add_wavemap(1, 8, "multilinear", None, [0], [8])
    with pytest.raises(BeatGpuError, match="not uploaded"):
        c.ffi_loglike_batch(np.ones((2, 4)))
    with pytest.raises(ValueError, match="do not match"):
        c.upload_gflib(wid, 0, np.zeros((2, 4, 2, 3, 8)), F64, 0.5, 0.5, 0.0, 0.5)     # wrong target count
    with pytest.raises(ValueError, match="patches"):
        c.upload_gflib(wid, 0, np.zeros((1, 5, 2, 3, 8)), F64, 0.5, 0.5, 0.0, 0.5)     # wrong patch count
    c.upload_gflib(wid, 0, np.zeros((1, 4, 2, 3, 8)), F64, 0.5, 0.5, 0.0, 0.5)
    with pytest.raises(BeatGpuError, match="data"):
        c.ffi_loglike_batch(np.ones((2, 4)))
    c.upload_data(wid, np.zeros((1, 8)))
    with pytest.raises(BeatGpuError, match="weights"):
        c.ffi_loglike_batch(np.ones((2, 4)))
    with pytest.raises(ValueError, match="NaN"):
        c.update_weights(wid, np.full((1, 8, 8), np.nan), [0.0])
    c.update_weights(wid, np.eye(8)[None], [0.0])
    logpts, like = c.ffi_loglike_batch(np.ones((2, 4)))
    assert logpts.shape == (2, 1) and np.isfinite(like).all()
    with pytest.raises(ValueError, match="unknown wavemap"):
        c.upload_data(7, np.zeros((1, 8)))
    c.close()


def test_host_register_roundtrip():
    from beat_b200 import synthetic
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=3, subfaults=((4, 5, 2.0),), ns=20, ndur=4, seed=2)
    Q = synthetic.draw_chains(prob, 10, seed=1)
    ev = BatchedFFILogLike.from_problem(prob, store_dtype="float64")
    ref, ref_like = ev(Q)
    Qp = ev.ctx.pin(np.ascontiguousarray(Q.copy()))
    logpts = ev.ctx.pin(np.empty((10, ev.n_out)))
    like = ev.ctx.pin(np.empty(10))
    ev.ctx.ffi_loglike_batch(Qp, logpts, like)
    assert np.array_equal(logpts, ref) and np.array_equal(like, ref_like)
    for a in (Qp, logpts, like):
        ev.ctx.unpin(a)
    with pytest.raises(Exception):
        ev.ctx.unpin(Qp)                 # not registered any more
    # a failed runtime call must not poison later launches (the CUDA last-error state is cleared)
    again, _ = ev(Q)
    assert np.array_equal(again, ref)
    ev.close()
