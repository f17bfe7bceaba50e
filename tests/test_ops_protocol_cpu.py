"""The Op contract of the reference's plugin boundary (beat/pytensorf.py:424 ``__props__``, :432-441 ``make_node``,
:443-500 ``perform``, :502-503 ``infer_shape``) exercised on the product's Op classes through a minimal eager runtime
that speaks pytensor's Op protocol (tests/_op_protocol.py; pytensor itself is not installed here).  No GPU: ``perform``
is fed by stand-in contexts backed by the CPU oracle, so what is tested is the host logic AROUND the kernels."""
import numpy as np
import pytest

from _op_protocol import Apply, Op, TensorVariable, op_protocol
from oracle import ffi_oracle as O


class _OracleSweepCtx(object):
    """Stand-in for lib.Context inside Sweeper.perform: same call, same return convention, CPU oracle underneath."""

    def __init__(self, nd, ns, h):
        self.nd, self.ns, self.h = nd, ns, h

    def fast_sweep_batch(self, sf, slowness, nuc_dip_idx, nuc_strike_idx, return_iters=False):
        out, it = O.fast_sweep_batch_port(slowness, self.h, nuc_dip_idx, nuc_strike_idx, self.nd, self.ns)
        return (out, it) if return_iters else out


def test_sweeper_is_an_op_with_the_references_contract():
    with op_protocol() as (ops, geometry):
        assert issubclass(ops.Sweeper, Op) and ops.HAVE_PYTENSOR
        # the reference's own test case (test/test_fastsweep.py:21-31, 84-112)
        patch_size, nuc_x, nuc_y, n_patch_strike, n_patch_dip = 10.0, 2, 3, 4, 6
        a = ops.Sweeper(patch_size, n_patch_dip, n_patch_strike, "cuda")
        b = ops.Sweeper(patch_size, n_patch_dip, n_patch_strike, "cuda")
        c = ops.Sweeper(patch_size, n_patch_dip, n_patch_strike + 1, "cuda")
        assert ops.Sweeper.__props__ == ("patch_size", "n_patch_dip", "n_patch_strike", "implementation")   # pytensorf.py:424
        assert a == b and hash(a) == hash(b) and a != c                   # graph merging relies on __props__ identity
        assert a.infer_shape() == [(24,)]                                  # pytensorf.py:502-503
        velocities = np.concatenate((np.ones((n_patch_dip, 2)), np.ones((n_patch_dip, 2)) * 3.5), axis=1)
        slow = (1.0 / velocities).flatten()
        node = a.make_node(slow, nuc_y, nuc_x)                             # pytensorf.py:432-441
        assert isinstance(node, Apply) and node.op is a and len(node.inputs) == 3 and len(node.outputs) == 1
        assert node.outputs[0].type.ndim == 1 and node.outputs[0].type.dtype == "float64"
        # __call__ = make_node + perform + type / shape check, like a compiled graph would run the Op
        a._ctx = _OracleSweepCtx(n_patch_dip, n_patch_strike, patch_size)
        out = a(slow, nuc_y, nuc_x)
        assert isinstance(out, TensorVariable) and out.owner.op is a
        ref = O.fast_sweep(slow, patch_size, nuc_y, nuc_x, n_patch_dip, n_patch_strike, impl="port")
        np.testing.assert_array_equal(out.eval(), ref)
        assert out.eval()[3 * 4 + 2] == 0.0                               # hypocentre patch (dip 3, strike 2)
        # the batch axis this Op adds declares (and produces) a matrix
        B = 5
        slows = np.tile(slow, (B, 1)) * np.linspace(1.0, 1.4, B)[:, None]
        outb = a(slows, np.full(B, nuc_y), np.full(B, nuc_x))
        assert outb.type.ndim == 2 and outb.eval().shape == (B, 24)
        np.testing.assert_array_equal(outb.eval()[0], ref)
        # error behaviour of the reference wrapper
        with pytest.raises(NotImplementedError):
            ops.Sweeper(patch_size, n_patch_dip, n_patch_strike, "fortran")(slow, nuc_y, nuc_x)   # pytensorf.py:494-498
        with pytest.raises(AttributeError, match="unexpected size"):
            a(slow[:-1], nuc_y, nuc_x)                                     # fast_sweep_ext.c:36-39


def test_fused_loglike_op_declares_what_it_returns():
    with op_protocol() as (ops, geometry):
        class _Ev(object):
            n_out = 3

            def __call__(self, q):
                q = np.atleast_2d(q)
                lp = np.stack([-(q ** 2).sum(1), -np.abs(q).sum(1), q[:, 0]], axis=1)
                return lp, lp.sum(1)

        op = ops.FFILogLike(_Ev(), name="ffi")
        assert op == ops.FFILogLike(_Ev(), name="ffi") and op != ops.FFILogLike(_Ev(), name="other")
        q = np.arange(4.0)
        logpts, like = op(q)
        assert logpts.type.ndim == 1 and like.type.ndim == 0 and logpts.eval().shape == (3,)
        assert like.eval() == logpts.eval().sum()
        Q = np.arange(8.0).reshape(2, 4)
        logpts, like = op(Q)
        assert logpts.type.ndim == 2 and like.type.ndim == 1 and logpts.eval().shape == (2, 3) and like.eval().shape == (2,)


def test_seis_synthesizer_make_node_takes_the_references_dict():
    """SeisSynthesizer.make_node (pytensorf.py:215-239): a DICT of named tensors in, varnames remembered in call order,
    (synthetics matrix, tmins vector) out."""
    with op_protocol() as (ops, geometry):
        s = object.__new__(geometry.SeisSynthesizer)            # no device here: only the graph-building half
        s.nt, s.ns, s.varnames = 6, 40, []
        inputs = {"time": 1.0, "depth": 5.0, "east_shift": 0.0, "north_shift": 0.0, "strike": 10.0, "dip": 40.0,
                  "rake": 90.0, "magnitude": 6.0, "duration": 2.0}
        node = s.make_node(inputs)
        assert s.varnames == list(inputs.keys())                           # order of the dict, not of the source class
        assert isinstance(node, Apply) and len(node.inputs) == 9
        assert [o.type.ndim for o in node.outputs] == [2, 1]
        assert s.infer_shape() == [(6, 40), (6,)]                          # pytensorf.py:303-311
        assert geometry.SeisSynthesizer.__props__[:3] == ("store", "event", "targets")
