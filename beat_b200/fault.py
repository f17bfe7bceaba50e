"""
The four index helpers of ``FaultGeometry`` that sit on the hot path, batched and GPU-backed.

Reference: beat/ffi/fault.py -- ``fault_locations2idxs`` (:866-894), ``vector2subfault`` (:610-612),
``point2starttimes`` (:614-632), ``get_subfault_starttimes`` (:722-752), the ordering of
``FaultOrdering`` (:1114-1169); ``positions2idxs`` is beat/utility.py:1542-1558.  Geometry construction,
discretisation and pyrocko sources stay with the reference.
"""
from __future__ import annotations

import numpy as np

from .lib import Context


def positions2idxs(positions, cell_size, min_pos=0.0, backend=np, dtype="int16"):
    """Index of the grid cell a position [km] falls into: round-half-even((pos - min - cell/2) / cell) (utility.py:1542-1558)."""
    return backend.round((positions - min_pos - (cell_size / 2.0)) / cell_size).astype(dtype)


class FaultOrdering(object):
    """Patch bookkeeping for faults with uniform grid size (fault.py:1114-1169)."""

    def __init__(self, npls, npws, patch_sizes_strike, patch_sizes_dip):
        self.patch_sizes_dip = list(patch_sizes_dip)
        self.patch_sizes_strike = list(patch_sizes_strike)
        self.shapes, self.slices = [], []
        dim = 0
        for npl, npw in zip(npls, npws):
            self.slices.append(slice(dim, dim + npl * npw))
            self.shapes.append((npw, npl))
            dim += npl * npw
        self.npatches = dim

    def get_subfault_discretization(self, index):
        """(n_patch_dip, n_patch_strike) of a subfault."""
        return self.shapes[index]


class FaultGeometry(object):
    """Subfault grids + the start-time helpers, rupture times computed by the batched CUDA sweeper."""

    def __init__(self, ordering, device=0):
        self.ordering = ordering
        self.nsubfaults = len(ordering.shapes)
        self.npatches = ordering.npatches
        self.cum_subfault_npatches = np.concatenate([[0], np.cumsum([s[0] * s[1] for s in ordering.shapes])])
        self._device = device
        self._ctx = None

    def _context(self):
        if self._ctx is None:
            ctx = Context(self._device)
            ctx.set_fault([s[0] for s in self.ordering.shapes], [s[1] for s in self.ordering.shapes], self.ordering.patch_sizes_dip)
            self._ctx = ctx
        return self._ctx

    def _check_index(self, index):
        if index > self.nsubfaults - 1:
            raise TypeError("Subfault with index %i not defined!" % index)

    def vector2subfault(self, index, vector):
        """Slice of a per-patch vector (last axis) that belongs to subfault ``index`` (fault.py:610-612)."""
        lo, hi = self.cum_subfault_npatches[index: index + 2]
        return vector[..., lo:hi]

    def fault_locations2idxs(self, index, positions_dip, positions_strike, backend="numpy"):
        """Patch indexes of positions on the fault [km] (fault.py:866-894)."""
        if backend != "numpy":
            raise NotImplementedError("Backend not supported! Options: numpy")
        return (positions2idxs(np.asarray(positions_dip), self.ordering.patch_sizes_dip[index]),
                positions2idxs(np.asarray(positions_strike), self.ordering.patch_sizes_strike[index]))

    def get_subfault_starttimes(self, index, rupture_velocities, nuc_dip_idx, nuc_strike_idx):
        """Rupture onset times of one subfault (fault.py:722-752), for one chain ([np_sf]) or a batch ([B, np_sf]).

        Returns [n_patch_dip, n_patch_strike] (single) or [B, n_patch_dip, n_patch_strike]."""
        self._check_index(index)
        npw, npl = self.ordering.get_subfault_discretization(index)
        v = np.asarray(rupture_velocities, dtype=np.float64)
        single = v.ndim == 1 or (v.ndim == 2 and v.shape == (npw, npl))
        v2 = v.reshape(1, -1) if single else v.reshape(v.shape[0], -1)
        t = self._context().fast_sweep_batch(index, 1.0 / v2, np.atleast_1d(nuc_dip_idx), np.atleast_1d(nuc_strike_idx))
        t = t.reshape((-1, npw, npl))
        return t[0] if single else t

    def point2starttimes(self, point, index=0):
        """Start times for a point (dict of variables; values may carry a leading chain axis) (fault.py:614-632)."""
        nuc_dip = np.asarray(point["nucleation_dip"])[..., index]
        nuc_strike = np.asarray(point["nucleation_strike"])[..., index]
        time = np.asarray(point["time"])[..., index]
        velocities = self.vector2subfault(index, np.asarray(point["velocities"]))
        nuc_dip_idx, nuc_strike_idx = self.fault_locations2idxs(index, nuc_dip, nuc_strike)
        st = self.get_subfault_starttimes(index, velocities, nuc_dip_idx, nuc_strike_idx)
        return st + (time[..., None, None] if np.ndim(time) else time)
