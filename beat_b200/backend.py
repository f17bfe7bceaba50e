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
    var_dtypes: name -> numpy dtype string (default float64).

    ``n_io_threads = 0``: ``flush`` writes the files before it returns, chain by chain with ``ndarray.tofile`` (the
    reference's behaviour, pure numpy).  ``n_io_threads > 0``: two step buffers; a full buffer is handed to ONE background
    thread that calls the native ``beatgpu_trace_append`` (libbeatgpu: every chain's records leave with one ``writev``
    straight from the step-major buffer, chains split over ``n_io_threads`` native threads, no interpreter lock) while the
    sampler fills the other buffer -- the sampler only waits when the writers are a whole buffer behind.  Flushes are
    queued on that one thread, so a chain's appends keep their order.  ``pinned=True`` page-locks the buffers (torch) so that a GPU sampler
    can copy a step's packed records straight into ``slot()`` (see ``DeviceRecorder``) without any host-side packing."""

    flat_names_tag, var_shape_tag, var_dtypes_tag = "flat_names", "var_shapes", "var_dtypes"   # backend.py:680-682

    def __init__(self, dir_path, var_shapes, n_chains, var_dtypes=None, buffer_size=5000, chain_offset=0, n_io_threads=0,
                 pinned=False):
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
        self.n_io_threads = int(n_io_threads)
        self._pinned_keep = []
        n_buf = 2 if self.n_io_threads > 0 else 1
        self._bufs = [self._alloc(pinned) for _ in range(n_buf)]
        self._pending = [[] for _ in range(n_buf)]              # futures of the flush that last used each buffer
        self._cur = 0
        self._buf = self._bufs[0]
        self._n = 0
        self.stored_samples = 0
        self._pool = None
        if self.n_io_threads > 0:
            from concurrent.futures import ThreadPoolExecutor
            from . import lib as _beatlib
            _beatlib.load()                                  # fail here, not in the background thread, if the library is missing
            self._native_append = _beatlib.trace_append
            self._pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix="beat_b200_trace")   # one thread: flushes stay ordered

    # ------------------------------------------------------------------ buffers
    def _alloc(self, pinned):
        shape = (self.buffer_size, self.n_chains)
        if not pinned:
            return np.zeros(shape, dtype=self.data_structure)
        import torch
        t = torch.zeros(self.buffer_size * self.n_chains * self.data_structure.itemsize, dtype=torch.uint8).pin_memory()
        self._pinned_keep.append(t)
        return t.numpy().view(self.data_structure).reshape(shape)

    @property
    def record_width(self):
        """float64 values per record when every variable is float64 (the packed-record fast path), else None."""
        if any(np.dtype(d) != np.float64 for d in self.var_dtypes.values()):
            return None
        return self.data_structure.itemsize // 8

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

    # ------------------------------------------------------------------ filling
    def write(self, values):
        """Buffer one step of all chains.  values: dict name -> array [n_chains, *shape]."""
        if self._n == self.buffer_size:
            self.flush(wait=False)
        row = self._buf[self._n]
        for name in self.varnames:
            row[name] = np.asarray(values[name]).reshape((self.n_chains,) + self.var_shapes[name])
        self._n += 1

    def slot(self):
        """The next step's records as a float64 matrix [n_chains, record_width] INSIDE the buffer (all-float64 layouts):
        fill it (e.g. by a device-to-host copy), then call ``commit()``."""
        if self.record_width is None:
            raise TypeError("packed records need an all-float64 layout")
        if self._n == self.buffer_size:
            self.flush(wait=False)
        return self._buf[self._n].view(np.float64).reshape(self.n_chains, self.record_width)

    def commit(self):
        self._n += 1

    def write_records(self, rec):
        """One step of all chains already packed in record order: rec [n_chains, record_width] float64."""
        self.slot()[...] = rec
        self.commit()

    # ------------------------------------------------------------------ writing
    def _append_chains(self, block, c0, c1):
        for c in range(c0, c1):
            with open(self.filename(c), mode="ab") as fh:
                np.ascontiguousarray(block[:, c]).tofile(fh)         # backend.py:822-845: records appended with tofile

    def flush(self, wait=True):
        """Append every chain's buffered records to its file (backend.py:822-845) and clear the buffer.  With writer
        threads and ``wait=False`` the append runs in the background and filling continues in the other buffer."""
        if self._n > 0:
            block = self._buf[: self._n]
            if self._pool is None:
                self._append_chains(block, 0, self.n_chains)
            else:
                self._pending[self._cur] = [self._pool.submit(self._native_append, self.dir_path, self.chain_offset, block,
                                                              self.n_io_threads)]
                self._cur = (self._cur + 1) % len(self._bufs)
                self._wait(self._cur)                                   # the buffer we are about to fill must be on disk
                self._buf = self._bufs[self._cur]
            self.stored_samples += self._n
            self._n = 0
        if wait:
            for i in range(len(self._bufs)):
                self._wait(i)

    def _wait(self, i):
        for f in self._pending[i]:
            f.result()                                                  # re-raises a writer thread's exception here
        self._pending[i] = []

    def close(self):
        try:
            self.flush(wait=True)
        finally:
            if self._pool is not None:
                self._pool.shutdown(wait=True)
                self._pool = None


class DeviceRecorder(object):
    """``on_step`` companion for a GPU sampler: packs a step's outputs in record order ON THE DEVICE (one ``torch.cat``),
    copies the packed matrix into the writer's page-locked buffer on a side stream, and lets the writer's threads put it on
    disk -- the sampling stream never waits for the host.  ``record(*tensors)``: tensors in the writer's variable order,
    each [n_chains, k] or [n_chains].  ``finish()`` drains the copies and flushes."""

    def __init__(self, writer, torch, device):
        if writer.record_width is None:
            raise TypeError("DeviceRecorder needs an all-float64 record layout")
        self.w, self.torch, self.device = writer, torch, device
        self.stream = torch.cuda.Stream(device=device)
        self._last = None                                  # event of the newest copy
        self._keep = []                                    # packed device matrices whose copies may still be in flight

    def record(self, *tensors):
        torch, w = self.torch, self.w
        rec = torch.cat([t.reshape(w.n_chains, -1) for t in tensors], dim=1)
        if rec.shape[1] != w.record_width or rec.dtype != torch.float64:
            raise ValueError("packed record is %s %s, the layout wants [%d, %d] float64" % (tuple(rec.shape), rec.dtype, w.n_chains, w.record_width))
        if w._n == w.buffer_size:                          # the buffer is about to be flushed: its copies must have landed
            self._drain()
        dst = torch.from_numpy(w.slot())                   # view into the writer's page-locked buffer
        ready = torch.cuda.Event()
        ready.record(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(self.stream):
            self.stream.wait_event(ready)
            dst.copy_(rec, non_blocking=True)
            done = torch.cuda.Event()
            done.record(self.stream)
        self._last = done
        self._keep.append(rec)
        w.commit()

    def _drain(self):
        if self._last is not None:
            self._last.synchronize()
            self._last = None
        self._keep = []

    def finish(self):
        self._drain()
        self.w.flush(wait=True)


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
