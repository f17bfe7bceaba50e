"""GPU tests of the C-ABI's error behaviour: call-order and argument errors are reported, never repaired."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_call_order_and_argument_errors():
    from beat_b200.lib import BeatGpuError, Context, GFLibraryError, Layout, F64
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
    with pytest.raises(GFLibraryError, match="1 targets x 8 samples"):
        c.upload_gflib(wid, 0, np.zeros((2, 4, 2, 3, 8)), F64, 0.5, 0.5, 0.0, 0.5)     # wrong target count
    with pytest.raises(ValueError, match="do not match"):                               # ... and the C side refuses it too
        c.alloc_gflib(wid, 0, F64, (2, 4, 2, 3, 8), 0.5, 0.5, 0.0, 0.5)
    # every operand is shape-checked before its pointer crosses the ABI (the C side reads nt*ns, nt*ns*ns, canon_len ... elements)
    with pytest.raises(ValueError, match="expected shape"):
        c.upload_data(wid, np.zeros((1, 7)))
    with pytest.raises(ValueError, match="expected shape"):
        c.update_weights(wid, np.zeros((1, 8, 7)), [0.0])
    with pytest.raises(ValueError, match="expected shape"):
        c.update_weights(wid, np.eye(8)[None], [0.0, 0.0])
    with pytest.raises(ValueError, match="expected shape"):
        c.set_layout(L, np.zeros(15))                                                   # fixed vector one short
    with pytest.raises(ValueError, match="expected shape"):
        c.set_laplacian(np.eye(5), 0.0, 0)
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


def test_probe_gather_entry():
    """The diagnostics entry measures something positive in every mode and rejects bad arguments."""
    from beat_b200.lib import Context
    c = Context(0)
    for mode in (0, 1, 2):
        for row_bytes in (480, 4096):
            assert c.probe_gather(mode, 8 << 20, row_bytes, 32, 1) > 1.0          # GB/s
    for mode in (3, 4, 5, 6, 7, 8):                                # batched bulk copies, shared memory, DSMEM of 2 / 4 / 8 CTAs
        for row_bytes in (480, 960):
            assert c.probe_gather(mode, 8 << 20, row_bytes, 64, 1) > 1.0
    with pytest.raises(ValueError):
        c.probe_gather(0, 8 << 20, 100, 32, 1)                 # row size not a multiple of 16
    with pytest.raises(ValueError):
        c.probe_gather(9, 8 << 20, 480, 32, 1)                 # unknown mode
    with pytest.raises(ValueError):
        c.probe_gather(3, 8 << 20, 8192, 32, 1)                # batched ring would not fit in shared memory
    c.close()


def test_upload_paths_agree(monkeypatch):
    """GF-library upload: pageable source through the pinned double buffers, direct copies, and a page-locked source give
    the same library (checked through stack_batch), also when the chunking splits the rows unevenly."""
    from beat_b200.lib import F32, F64, Context
    rng = np.random.default_rng(5)
    nt, npatch, ndur, nst, ns = 3, 7, 4, 9, 50
    G = rng.standard_normal((nt, npatch, ndur, nst, ns))
    dur = rng.uniform(0.5, 1.9, (6, npatch))
    st = rng.uniform(0.1, 3.9, (6, nt, npatch))
    slip = rng.uniform(0, 2, (1, 6, npatch))

    monkeypatch.setenv("BEATGPU_UPLOAD_CHUNK_KB", "37")        # 756 rows of 400 B in chunks of 94 rows: 9 chunks, the last one short

    def run(store, pinned=False, direct=False):
        if direct:
            monkeypatch.setenv("BEATGPU_UPLOAD_DIRECT", "1")
        else:
            monkeypatch.delenv("BEATGPU_UPLOAD_DIRECT", raising=False)
        c = Context(0)
        c.set_fault([1], [npatch], [1.0])
        wid = c.add_wavemap(nt, ns, "multilinear", None, np.zeros(nt, np.int32), np.full(nt, ns, np.int32))
        src = G.copy()
        if pinned:
            c.pin(src)
        c.upload_gflib(wid, 0, src, store, 0.5, 0.5, 0.0, 0.5)
        if pinned:
            c.unpin(src)
        out = c.stack_batch(wid, dur, st, slip, nt, ns)
        c.close()
        return out

    for store in (F64, F32):
        ref = run(store)
        np.testing.assert_array_equal(run(store, direct=True), ref)
        np.testing.assert_array_equal(run(store, pinned=True), ref)
