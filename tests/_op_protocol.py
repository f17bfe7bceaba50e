"""
tests/_op_protocol.py -- TEST INFRASTRUCTURE: a minimal eager runtime that speaks pytensor's ``Op`` protocol.

pytensor is not installed in the build container nor on the GPU box, so the product's Op classes normally fall back to
a plain calling convention and their ``make_node`` / ``infer_shape`` / ``__props__`` never execute.  This module
installs a stand-in ``pytensor`` package whose ``Op.__call__`` does what pytensor's graph machinery does around a Python
Op (pytensor/graph/op.py: ``__call__`` -> ``make_node`` -> ``Apply``; ``perform(node, inputs, output_storage)`` at run
time; ``__props__`` -> generated ``__eq__`` / ``__hash__``; outputs checked against the declared ``TensorType``), eagerly
on numpy arrays.  Nothing here computes anything of the hot path; it only drives the Ops the way a compiled graph does.

    with op_protocol() as (ops, geometry):      # beat_b200.ops / beat_b200.geometry re-imported against the stand-in
        sweeper = ops.Sweeper(...)
        t0 = sweeper(slowness, nuc_dip_idx, nuc_strike_idx)      # make_node -> perform -> type check
"""
import contextlib
import importlib
import sys
import types

import numpy as np


class TensorType(object):
    def __init__(self, dtype, ndim, shape=None):
        self.dtype, self.ndim, self.shape = str(dtype), int(ndim), shape

    def __call__(self):                      # ``outv.type()`` creates a fresh variable of that type
        return TensorVariable(self, None)

    def __eq__(self, other):
        return isinstance(other, TensorType) and (self.dtype, self.ndim) == (other.dtype, other.ndim)

    def __hash__(self):
        return hash((self.dtype, self.ndim))

    def filter(self, value):
        """What the compiled function does with ``output_storage`` contents: the array must match the declared type."""
        value = np.asarray(value)
        if value.ndim != self.ndim:
            raise TypeError("Op output has %d dimensions, node declared %d" % (value.ndim, self.ndim))
        if str(value.dtype) != self.dtype:
            raise TypeError("Op output has dtype %s, node declared %s" % (value.dtype, self.dtype))
        return value


class TensorVariable(object):
    def __init__(self, type_, value, owner=None, index=None):
        self.type, self.value, self.owner, self.index = type_, value, owner, index

    @property
    def ndim(self):
        return self.type.ndim

    def eval(self):
        return self.value


def as_tensor_variable(x):
    if isinstance(x, TensorVariable):
        return x
    arr = np.asarray(x)
    return TensorVariable(TensorType(arr.dtype, arr.ndim, arr.shape), arr)


class Apply(object):
    def __init__(self, op, inputs, outputs):
        self.op, self.inputs, self.outputs = op, list(inputs), list(outputs)
        for i, o in enumerate(self.outputs):
            if not isinstance(o, TensorVariable):
                raise TypeError("Apply outputs must be variables")
            o.owner, o.index = self, i


class Op(object):
    """Python Op: ``__props__`` identity, ``__call__`` = make_node + (eager) perform."""
    __props__ = ()

    def _props(self):
        return tuple(getattr(self, p) for p in self.__props__)

    def __eq__(self, other):
        if type(self) is not type(other):
            return False
        try:
            return all(_same(a, b) for a, b in zip(self._props(), other._props()))
        except Exception:
            return False

    def __hash__(self):
        return hash((type(self), tuple(_hashable(p) for p in self._props())))

    def __call__(self, *inputs, **kwargs):
        node = self.make_node(*inputs, **kwargs)
        if not isinstance(node, Apply) or node.op is not self:
            raise TypeError("make_node must return an Apply node of this Op")
        storage = [[None] for _ in node.outputs]
        self.perform(node, [v.eval() for v in node.inputs], storage)
        shapes = self.infer_shape(None, node, [np.shape(v.eval()) for v in node.inputs])
        outs = []
        for var, cell, shp in zip(node.outputs, storage, shapes):
            if cell[0] is None:
                raise ValueError("perform left an output unset")
            var.value = var.type.filter(cell[0])
            if tuple(shp) != tuple(var.value.shape):
                raise ValueError("infer_shape says %s, perform produced %s" % (tuple(shp), var.value.shape))
            outs.append(var)
        return outs[0] if len(outs) == 1 else outs


def _same(a, b):
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.array_equal(a, b)
    return a is b or a == b


def _hashable(p):
    if isinstance(p, np.ndarray):
        return p.tobytes()
    try:
        hash(p)
        return p
    except TypeError:
        return id(p)


def _vector(dtype, ndim):
    def make(name=None):
        return TensorVariable(TensorType(dtype, ndim), None)
    return make


def _build_modules():
    pt = types.ModuleType("pytensor")
    tt = types.ModuleType("pytensor.tensor")
    gr = types.ModuleType("pytensor.graph")
    tt.Op, tt.as_tensor_variable, tt.TensorType, tt.TensorVariable = Op, as_tensor_variable, TensorType, TensorVariable
    tt.dscalar, tt.dvector, tt.dmatrix = _vector("float64", 0), _vector("float64", 1), _vector("float64", 2)
    gr.Apply, gr.Op = Apply, Op
    pt.tensor, pt.graph = tt, gr
    return {"pytensor": pt, "pytensor.tensor": tt, "pytensor.graph": gr}


@contextlib.contextmanager
def op_protocol():
    """Re-import beat_b200.ops / beat_b200.geometry against the stand-in pytensor; restore the plain modules afterwards."""
    saved = {k: sys.modules.get(k) for k in ("pytensor", "pytensor.tensor", "pytensor.graph")}
    sys.modules.update(_build_modules())
    import beat_b200.geometry as geometry
    import beat_b200.ops as ops
    try:
        ops = importlib.reload(ops)
        geometry = importlib.reload(geometry)
        assert ops.HAVE_PYTENSOR and issubclass(ops.Sweeper, Op)
        yield ops, geometry
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
        importlib.reload(ops)
        importlib.reload(geometry)
