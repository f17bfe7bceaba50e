"""
Batched writer / reader for the reference's binary trace format.

The reference appends one structured record per Metropolis step to ``chain-<i>.bin`` (beat/backend.py:651-897,
``NumpyChain``): a one-line JSON header (``flat_names``, ``var_shapes``, ``var_dtypes``; :765-782) followed by
``numpy.dtype({'names': varnames, 'formats': ['<shape><dtype>', ...]})`` records (:797-820) written with
``ndarray.tofile`` (:833-837).  With a lock-step batched evaluator every step yields the outputs of ALL chains at
once, so this writer buffers ``[n_steps, n_chains]`` records in one structured array and appends each chain's
column to its file when flushed (per stage): the files are byte-compatible with ``NumpyChain`` / ``load_multitrace``.
"""
from __future__ import annotations

import json
import os
from collections import OrderedDict

import numpy as np


def create_flat_names(varname, shape):
    """pymc's flat-name convention used in the header (``var__0``, ``var__0_1`` ...)."""
    if not shape:
        return [varname]
    labels = (np.ravel(xs).tolist() for xs in np.indices(shape))
    labels = (map(str, xs) for xs in labels)
    return ["{}__{}".format(varname, "_".join(idxs)) for idxs in zip(*labels)]


class BatchedNumpyChains(object):
    """Append the per-step outputs of all chains to ``<dir_path>/chain-<i>.bin`` in NumpyChain layout.

    var_shapes: OrderedDict name -> shape tuple (order = ``model.unobserved_RVs`` order, beat/sampler/metropolis.py:160-162);
    var_dtypes: name -> numpy dtype string (default float64)."""

    flat_names_tag, var_shape_tag, var_dtypes_tag = "flat_names", "var_shapes", "var_dtypes"   # backend.py:680-682

    def __init__(self, dir_path, var_shapes, n_chains, var_dtypes=None, buffer_size=5000, chain_offset=0):
        os.makedirs(dir_path, exist_ok=True)
        self.dir_path = dir_path
        self.var_shapes = OrderedDict((k, tuple(v)) for k, v in var_shapes.items())
        self.varnames = list(self.var_shapes.keys())
        self.var_dtypes = OrderedDict((k, str((var_dtypes or {}).get(k, "float64"))) for k in self.varnames)
        self.flat_names = OrderedDict((k, create_flat_names(k, s)) for k, s in self.var_shapes.items())
        self.n_chains, self.chain_offset = n_chains, chain_offset
        self.data_structure = np.dtype({"names": self.varnames,
                                        "formats": ["{}{}".format(self.var_shapes[n], self.var_dtypes[n]) for n in self.varnames]})
        self.buffer_size = buffer_size
        self._buf = np.zeros((buffer_size, n_chains), dtype=self.data_structure)
        self._n = 0
        self.stored_samples = 0

    def filename(self, chain):
        return os.path.join(self.dir_path, "chain-{}.bin".format(chain + self.chain_offset))

    def setup(self, overwrite=True):
        """Create the files with their headers (backend.py:735-782)."""
        header = (json.dumps({self.flat_names_tag: self.flat_names, self.var_shape_tag: self.var_shapes,
                              self.var_dtypes_tag: self.var_dtypes}) + "\n").encode()
        for c in range(self.n_chains):
            if overwrite or not os.path.exists(self.filename(c)):
                with open(self.filename(c), "wb") as fh:
                    fh.write(header)

    def write(self, values):
        """Buffer one step of all chains.  values: dict name -> array [n_chains, *shape]."""
        if self._n == self.buffer_size:
            self.flush()
        row = self._buf[self._n]
        for name in self.varnames:
            row[name] = np.asarray(values[name]).reshape((self.n_chains,) + self.var_shapes[name])
        self._n += 1

    def flush(self):
        """Append every chain's buffered records to its file (backend.py:822-845) and clear the buffer."""
        if self._n == 0:
            return
        block = self._buf[: self._n]
        for c in range(self.n_chains):
            with open(self.filename(c), mode="ab") as fh:
                np.ascontiguousarray(block[:, c]).tofile(fh)
        self.stored_samples += self._n
        self._n = 0


def read_chain(filename):
    """Read a ``chain-<i>.bin`` file (backend.py:784-866): returns (structured array, var_shapes)."""
    with open(filename, "rb") as fh:
        header = json.loads(fh.readline().decode(), object_pairs_hook=OrderedDict)
        var_shapes = OrderedDict((k, tuple(v)) for k, v in header[BatchedNumpyChains.var_shape_tag].items())
        dtypes = header[BatchedNumpyChains.var_dtypes_tag]
        names = list(header[BatchedNumpyChains.flat_names_tag].keys())
        dt = np.dtype({"names": names, "formats": ["{}{}".format(var_shapes[n], dtypes[n]) for n in names]})
        data = np.fromfile(fh, dtype=dt)
    return data, var_shapes


def get_values(filename, varname, burn=0, thin=1):
    """``NumpyChain.get_values`` (backend.py:868-878)."""
    data, var_shapes = read_chain(filename)
    vals = data[varname].ravel().reshape((data.shape[0],) + var_shapes[varname])
    return vals[burn::thin]
