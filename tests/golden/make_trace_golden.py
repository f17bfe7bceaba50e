"""
tests/golden/make_trace_golden.py -- regenerates tests/golden/trace_golden.npz by running the REFERENCE'S OWN trace
backend (beat/backend.py:651-897, ``NumpyChain``; ``FileChain.write`` / ``record_buffer`` :369-400) on small chains.

Run in the dev container only (needs /root/reference):

    python tests/golden/make_trace_golden.py

What runs is the reference's source, imported from /root/reference through ``_refshim`` (pymc / arviz / pyrocko are
absent here; the shim supplies inert stand-ins -- ``NumpyChain`` itself only needs numpy, json and os).  For each of
3 chains the reference class is set up (``setup(draws, chain, overwrite=True)``), fed 7 steps through ``write(lpoint,
draw)`` with a buffer of 3 (so ``record_buffer`` appends three times) and flushed; the bytes of the resulting
``chain-<i>.bin`` files and the values written are stored.  The script also lets the REFERENCE's reader
(``NumpyChain.get_values`` / ``point``) read files written by ``beat_b200.backend.BatchedNumpyChains`` and records that
they came back identical.
"""
import os
import shutil
import sys
import tempfile
import warnings
from collections import OrderedDict

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)

import _refshim  # noqa: E402

warnings.simplefilter("ignore")
_refshim.install()

from beat import backend as rb  # noqa: E402

from beat_b200 import backend as bk  # noqa: E402


class _Tag(object):
    def __init__(self, v):
        self.test_value = v


class _Var(object):
    """What BaseChain.__init__ reads of a model variable: .name and .tag.test_value (backend.py:208-214)."""

    def __init__(self, name, shape):
        self.name = name
        self.tag = _Tag(np.zeros(shape, dtype=np.float64))


def main():
    rng = np.random.default_rng(20261017)
    # the unobserved RVs of a small FFI problem in model order (beat/sampler/metropolis.py:160-162)
    var_shapes = OrderedDict([("uparr", (6,)), ("uperp", (6,)), ("durations", (6,)), ("velocities", (6,)),
                              ("nucleation_strike", (1,)), ("nucleation_dip", (1,)), ("time", (1,)),
                              ("h_any_P_0_Z", (1,)), ("seis_like", (4,)), ("like", ())])
    value_vars = [_Var(k, s) for k, s in var_shapes.items()]
    n_chains, n_steps, buffer_size = 3, 7, 3
    steps = [{k: rng.standard_normal((n_chains,) + s) for k, s in var_shapes.items()} for _ in range(n_steps)]
    out = {"n_chains": n_chains, "n_steps": n_steps, "buffer_size": buffer_size,
           "varnames": np.array(list(var_shapes.keys())), "shapes": np.array([str(s) for s in var_shapes.values()])}
    for i, st in enumerate(steps):
        for k, v in st.items():
            out["step%d_%s" % (i, k)] = v
    d = tempfile.mkdtemp()
    try:
        ref_dir, our_dir = os.path.join(d, "ref"), os.path.join(d, "ours")
        for c in range(n_chains):
            chain = rb.NumpyChain(ref_dir, model=None, value_vars=value_vars, buffer_size=buffer_size)
            chain.setup(n_steps, c, overwrite=True)
            for i, st in enumerate(steps):
                chain.write([st[k][c] for k in var_shapes], i)          # the lpoint of one step (sampler/base.py:364)
            chain.record_buffer()                                        # chain end (sampler/base.py:390)
            out["file_chain%d" % c] = np.frombuffer(open(chain.filename, "rb").read(), dtype=np.uint8)
            out["header"] = np.frombuffer(chain.file_header.encode(), dtype=np.uint8)
        # our batched writer, read back by the reference's reader
        w = bk.BatchedNumpyChains(our_dir, var_shapes, n_chains, buffer_size=buffer_size)
        w.setup()
        for st in steps:
            w.write(st)
        w.flush()
        ok = True
        for c in range(n_chains):
            reader = rb.NumpyChain(our_dir, model=None, value_vars=value_vars)
            reader.setup(n_steps, c, overwrite=False)                    # existing file: "Found existing trace, appending!"
            for k, s in var_shapes.items():
                got = reader.get_values(k)
                want = np.array([st[k][c] for st in steps]).reshape((n_steps,) + s)
                ok = ok and got.shape == want.shape and np.array_equal(got, want)
            pt = reader.point(n_steps - 1)
            ok = ok and all(np.array_equal(pt[k], steps[-1][k][c]) for k in var_shapes)
            ok = ok and len(reader) == n_steps
            ok = ok and open(w.filename(c), "rb").read() == out["file_chain%d" % c].tobytes()
        out["reference_reader_reads_our_files"] = np.array(bool(ok))
        assert ok, "the reference's NumpyChain could not read back what BatchedNumpyChains wrote"
    finally:
        shutil.rmtree(d, ignore_errors=True)
    path = os.path.join(HERE, "trace_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; reference reader ok:", bool(out["reference_reader_reads_our_files"]))


if __name__ == "__main__":
    main()
