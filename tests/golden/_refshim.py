"""
tests/golden/_refshim.py -- import machinery used ONLY by tests/golden/make_golden.py in the dev container.

The reference (``/root/reference``, BEAT 2.0.5) cannot be imported as-is here: ``pytensor``, ``pymc``,
``pyrocko`` and ``matplotlib`` are not installed and there is no network.  To still run the reference's OWN
source for the hot-path functions (and so pin the oracle against real reference output), this module installs
stand-ins for exactly those absent third-party packages:

* ``pytensor`` / ``pytensor.tensor``: a numpy-backed eager shim of the handful of graph primitives the hot
  path uses (``zeros, cast, dot, exp, set_subtensor, shared, batched_dot, tile, repeat, concatenate, round,
  ceil``).  The formulas executed are the reference's; only the array backend is numpy.
* ``pyrocko.guts``: a minimal declarative ``Object`` (class-level ``X.T(...)`` specs become ``None`` / default
  attributes, ``__init__(**kwargs)`` sets them), enough for ``beat.heart.Covariance``.
* everything else under ``pyrocko``, ``pymc``, ``matplotlib``, ``arviz``, ``mpi4py`` ...: inert placeholder
  modules whose attributes are subclassable/callable dummies (never executed on the paths we call).

Nothing here is shipped, imported by the product, or used on the GPU box.
"""
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np

STUB_ROOTS = ("pytensor", "pyrocko", "pymc", "matplotlib", "arviz", "mpi4py", "cutde", "pygmsh", "tqdm_missing")


# ---------------------------------------------------------------- generic inert dummies
class _DummyMeta(type):
    def __getattr__(cls, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        return _new_dummy(name)

    def __call__(cls, *a, **k):
        inst = type.__call__(cls)
        inst.__dict__.update(k)
        return inst

    def __iter__(cls):
        return iter(())

    def __getitem__(cls, k):
        return cls

    def __len__(cls):
        return 0

    def __mro_entries__(cls, bases):  # pragma: no cover
        return (cls,)


def _new_dummy(name="Dummy"):
    return _DummyMeta(name, (), {
        "__init__": lambda self, *a, **k: None,
        "__call__": lambda self, *a, **k: _new_dummy(name)(),
        "__getattr__": lambda self, n: (_ for _ in ()).throw(AttributeError(n)) if n.startswith("__") else _new_dummy(n),
        "__iter__": lambda self: iter(()),
        "__len__": lambda self: 0,
        "__mul__": lambda self, o: self, "__rmul__": lambda self, o: self,
        "__add__": lambda self, o: self, "__radd__": lambda self, o: self,
        "__truediv__": lambda self, o: self, "__rtruediv__": lambda self, o: self,
        "__sub__": lambda self, o: self, "__rsub__": lambda self, o: self,
        "__getitem__": lambda self, k: self,
    })


class _StubModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__") and name.endswith("__"):
            raise AttributeError(name)
        d = _new_dummy(name)
        setattr(self, name, d)
        return d


# ---------------------------------------------------------------- pyrocko.guts (minimal)
class _TSpec:
    def __init__(self, *a, **k):
        self.default = k.get("default", None)
        if callable(self.default) and not isinstance(self.default, type):
            pass


class _GutsType:
    @classmethod
    def T(cls, *a, **k):
        return _TSpec(*a, **k)

    @classmethod
    def D(cls, *a, **k):
        return None


class _ObjectMeta(_DummyMeta):
    def __getattr__(cls, name):
        raise AttributeError(name)

    def __call__(cls, *a, **k):
        return type.__call__(cls, *a, **k)

    def __new__(mcls, name, bases, ns):
        specs = {}
        for b in bases:
            specs.update(getattr(b, "_guts_specs", {}))
        for k, v in list(ns.items()):
            if isinstance(v, _TSpec):
                specs[k] = v
                ns[k] = None
        ns["_guts_specs"] = specs
        return super().__new__(mcls, name, bases, ns)


class GutsObject(_GutsType, metaclass=_ObjectMeta):
    def __init__(self, **kwargs):
        for k, spec in self._guts_specs.items():
            d = spec.default
            setattr(self, k, d() if callable(d) else d)
        for k, v in kwargs.items():
            setattr(self, k, v)

    def regularize(self):
        pass

    def validate(self):
        pass


def _make_guts():
    m = _StubModule("pyrocko.guts")
    m.Object = GutsObject
    for n in ("Bool", "Dict", "Float", "Int", "List", "String", "StringChoice", "StringUnion", "Tuple",
              "Timestamp", "Any", "Choice", "Unicode", "Complex", "DateTimestamp", "SObject", "StringPattern"):
        setattr(m, n, type(n, (_GutsType,), {}))
    m.load = lambda *a, **k: None
    m.dump = lambda *a, **k: None
    m.ArgumentError = type("ArgumentError", (Exception,), {})
    m.ValidationError = type("ValidationError", (Exception,), {})
    ga = _StubModule("pyrocko.guts_array")
    ga.Array = type("Array", (_GutsType,), {})
    return m, ga


# ---------------------------------------------------------------- pytensor (numpy eager shim)
class Shared(np.ndarray):
    """numpy array with the get/set_value surface of a pytensor shared variable."""

    def __new__(cls, value, name=None, borrow=False, **kw):
        arr = np.array(value, copy=True)
        obj = arr.view(cls)
        obj.name = name
        return obj

    def __array_finalize__(self, obj):
        self.name = getattr(obj, "name", None)

    def set_value(self, v, borrow=False):
        v = np.asarray(v)
        if v.shape != self.shape:
            raise ValueError("Shared.set_value shape change unsupported in shim")
        self[...] = v

    def get_value(self, borrow=False):
        return np.asarray(self)

    def dimshuffle(self, *pattern):
        return _dimshuffle(np.asarray(self), *pattern)


def _dimshuffle(a, *pattern):
    if len(pattern) == 1 and isinstance(pattern[0], (tuple, list)):
        pattern = tuple(pattern[0])
    return np.transpose(a, pattern)


class _ND(np.ndarray):
    """ndarray view that knows .dimshuffle (used by stack_all's pytensor branch)."""

    def dimshuffle(self, *pattern):
        return _dimshuffle(np.asarray(self), *pattern).view(_ND)


def _set_subtensor(view, val):
    base = view.base if view.base is not None else view
    view[...] = val
    return base


def _batched_dot(a, b):
    # pytensor.tensor.batched_dot: batch over axis 0
    a, b = np.asarray(a), np.asarray(b)
    return np.einsum("bij,bj->bi", a, b) if b.ndim == 2 else np.einsum("bij,bjk->bik", a, b)


def _make_pytensor():
    pt = _StubModule("pytensor")
    tt = _StubModule("pytensor.tensor")
    cfg = types.SimpleNamespace(floatX="float64", compute_test_value="off")
    pt.config = cfg
    pt.shared = Shared
    pt.tensor = tt
    tt.zeros = lambda shape, dtype="float64": np.zeros(shape, dtype=dtype)
    tt.ones = lambda shape, dtype="float64": np.ones(shape, dtype=dtype)
    tt.cast = lambda x, dtype: np.asarray(x).astype(dtype)
    tt.dot = np.dot
    tt.exp = np.exp
    tt.log = np.log
    tt.round = np.round      # pytensor default rounding mode is half_to_even, as numpy
    tt.ceil = np.ceil
    tt.tile = np.tile
    tt.repeat = np.repeat
    tt.concatenate = lambda arrs, axis=0: np.concatenate(arrs, axis=axis).view(_ND)
    tt.set_subtensor = _set_subtensor
    tt.batched_dot = _batched_dot
    tt.as_tensor_variable = np.asarray
    tt.Op = type("Op", (), {"__call__": lambda self, *inputs: self._eager(*inputs)})

    def _eager(self, *inputs):
        out = [[None]]
        self.perform(None, [np.asarray(i) for i in inputs], out)
        return out[0][0]
    tt.Op._eager = _eager
    graph = _StubModule("pytensor.graph")
    graph.Apply = _new_dummy("Apply")
    pt.graph = graph
    return pt, tt, graph


# ---------------------------------------------------------------- finder
class _StubFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    def __init__(self):
        self.fixed = {}
        guts, ga = _make_guts()
        pt, tt, graph = _make_pytensor()
        self.fixed.update({"pyrocko.guts": guts, "pyrocko.guts_array": ga, "pytensor": pt,
                           "pytensor.tensor": tt, "pytensor.graph": graph})

    def find_spec(self, fullname, path=None, target=None):
        root = fullname.split(".")[0]
        if root in STUB_ROOTS or fullname in ("beat.info", "beat.defaults"):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        if spec.name in self.fixed:
            m = self.fixed[spec.name]
        else:
            m = _StubModule(spec.name)
        m.__path__ = []
        if spec.name == "beat.info":
            m.version = "2.0.5"
            m.__all__ = ["version"]
        return m

    def exec_module(self, module):
        pass


def install(reference_root="/root/reference", fast_sweep_ext=None):
    """Make ``import beat...`` resolve to the reference sources with absent third-party deps stubbed."""
    finder = _StubFinder()
    sys.meta_path.insert(0, finder)
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    if fast_sweep_ext is not None:
        sys.modules["fast_sweep_ext"] = fast_sweep_ext
    return finder
