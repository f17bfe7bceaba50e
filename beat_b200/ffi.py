"""
GF-library loader / writer in the reference's on-disk format.

The reference stores one library per (datatype, slip component, wavemap, crust index) as three files with the
prefix ``get_gf_prefix`` builds (beat/ffi/base.py:157-158):

    <prefix>.traces.npy   float64, C-order (ntargets, npatches, ndurations, nstarttimes, nsamples)   (base.py:364-373)
    <prefix>.times.npy    float64 [ntargets]  trace tmin per target                                   (base.py:372)
    <prefix>.yaml         pyrocko-guts dump of SeismicGFLibraryConfig (dimensions, axis origin/step ...)  (base.py:113-120,
                          beat/config.py:1900-1919)

``load_gf_library`` mirrors beat/ffi/base.py:161-189: the traces are memory-mapped (never fully read into host
RAM) and streamed into HBM in 256 MiB chunks by ``beatgpu_upload_gflib``, stored as float32 (default, half the
bytes) or float64 (strict).  The YAML is read with a tag-tolerant loader -- pyrocko is not needed to parse the few
scalars the hot path uses.
"""
from __future__ import annotations

import os

import numpy as np
import yaml

from .lib import GFLibraryError


def get_gf_prefix(datatype, component, wavename, crust_ind):
    """beat/ffi/base.py:157-158."""
    return "%s_%s_%s_%i" % (datatype, component, wavename, crust_ind)


class _TolerantLoader(yaml.SafeLoader):
    """SafeLoader that turns guts' application tags (``!beat.SeismicGFLibraryConfig``, ``!pf.Event`` ...) into plain
    dicts / lists / scalars."""


def _construct_any(loader, tag_suffix, node):
    if isinstance(node, yaml.MappingNode):
        return loader.construct_mapping(node, deep=True)
    if isinstance(node, yaml.SequenceNode):
        return loader.construct_sequence(node, deep=True)
    return loader.construct_scalar(node)


_TolerantLoader.add_multi_constructor("!", _construct_any)


def read_library_config(path):
    """Parse a ``<prefix>.yaml`` written by the reference (guts) or by :func:`save_gf_library`."""
    try:
        with open(path) as f:
            cfg = yaml.load(f, Loader=_TolerantLoader)
    except IOError:
        raise IOError("Cannot load config, file %s does not exist!" % path)       # base.py:122-126
    if not isinstance(cfg, dict) or "dimensions" not in cfg:
        raise GFLibraryError("%s is not a GF library config" % path)
    return cfg


def load_gf_library(directory="", filename=None, device=0, store_dtype="float32", interpolation="multilinear"):
    """Load one seismic GF library from disk onto the GPU (reference: beat/ffi/base.py:161-189).

    Returns a :class:`beat_b200.ops.SeismicGFLibrary` whose ``stack_all`` runs on the device.  The wavemap-level
    engine (``BatchedFFILogLike``) takes the memory-mapped array directly through ``prob['wavemaps'][i]['G']``."""
    from .ops import SeismicGFLibrary
    inpath = os.path.join(directory, filename)
    datatype = filename.split("_")[0]
    if datatype != "seismic":
        raise ValueError('datatype "%s" not supported!' % datatype)               # base.py:185-186 (geodetic: load_geodetic)
    cfg = read_library_config(inpath + ".yaml")
    traces = np.load(inpath + ".traces.npy", mmap_mode="r", allow_pickle=False)
    tmins = np.load(inpath + ".times.npy", mmap_mode="r", allow_pickle=False)
    if tuple(traces.shape) != tuple(int(x) for x in cfg["dimensions"]):
        raise GFLibraryError("traces shape %s does not match config dimensions %s" % (traces.shape, cfg["dimensions"]))
    comp = cfg.get("component", "uparr")
    gfs = SeismicGFLibrary({comp: traces}, duration_min=float(cfg.get("duration_min", 0.1)),
                           duration_sampling=float(cfg.get("duration_sampling", 0.5)),
                           starttime_min=float(cfg.get("starttime_min", 0.0)),
                           starttime_sampling=float(cfg.get("starttime_sampling", 0.5)),
                           interpolation_default=interpolation, store_dtype=store_dtype, device=device)
    gfs.config = cfg
    gfs._tmins = np.asarray(tmins)
    return gfs


def load_geodetic_library(directory="", filename=None):
    """Geodetic library (npatches, nobs) -- beat/ffi/base.py:178-183; returns (config, matrix memmap)."""
    inpath = os.path.join(directory, filename)
    cfg = read_library_config(inpath + ".yaml")
    G = np.load(inpath + ".traces.npy", mmap_mode="r", allow_pickle=False)
    return cfg, G


def save_gf_library(outdir, traces, tmins, component="uparr", wavename="any_P", mapnumber=0, crust_ind=0,
                    duration_min=0.1, duration_sampling=0.5, starttime_min=0.0, starttime_sampling=0.5, datatype="seismic"):
    """Write a library in the reference's layout (beat/ffi/base.py:364-373 + save_config :113-120).  Used by tests and
    to hand synthetic libraries to a reference installation.  Returns the file prefix."""
    traces = np.asarray(traces)
    mapid = "_".join((wavename, str(mapnumber))) if mapnumber is not None else wavename        # config.py:1911-1916
    prefix = get_gf_prefix(datatype, component, mapid, crust_ind)
    outpath = os.path.join(outdir, prefix)
    np.save(outpath + ".traces", arr=traces.astype(np.float64, copy=False), allow_pickle=False)
    np.save(outpath + ".times", arr=np.asarray(tmins, dtype=np.float64), allow_pickle=False)
    cfg = dict(component=component, crust_ind=int(crust_ind), starttime_sampling=float(starttime_sampling),
               duration_sampling=float(duration_sampling), starttime_min=float(starttime_min),
               duration_min=float(duration_min), dimensions=[int(x) for x in traces.shape], datatype=datatype,
               mapnumber=mapnumber, wave_config=dict(name=wavename))
    with open(outpath + ".yaml", "w") as f:
        f.write("# beat.ffi.SeismicGFLibrary YAML Config\n--- !beat.SeismicGFLibraryConfig\n")
        body = yaml.safe_dump(cfg, default_flow_style=False, sort_keys=False)
        f.write(body.replace("wave_config:\n", "wave_config: !beat.WaveformFitConfig\n"))
    return prefix


def discover_seismic_libraries(gfpath, slip_vars, crust_ind=0):
    """Find the seismic libraries of a project's ``<project>/ffi/linear_gfs`` directory (reference naming:
    ``seismic_<component>_<mapid>_<crust_ind>``, beat/ffi/base.py:157-158, beat/models/seismic.py:1168-1208).

    Returns an ordered list of wavemaps: ``[(mapid, {component: prefix})]`` for every mapid that has all ``slip_vars``."""
    import glob
    import re
    found = {}
    for path in sorted(glob.glob(os.path.join(gfpath, "seismic_*_%i.yaml" % crust_ind))):
        prefix = os.path.basename(path)[: -len(".yaml")]
        for comp in slip_vars:
            m = re.match(r"^seismic_%s_(.+)_%i$" % (re.escape(comp), crust_ind), prefix)
            if m:
                found.setdefault(m.group(1), {})[comp] = prefix
    return [(mapid, comps) for mapid, comps in sorted(found.items()) if all(c in comps for c in slip_vars)]


def wavemaps_from_directory(gfpath, slip_vars, interpolation="multilinear", crust_ind=0):
    """Build the ``prob['wavemaps']`` entries (library part) of :class:`beat_b200.engine.BatchedFFILogLike` straight from
    the reference's library files: traces stay memory-mapped and are streamed to HBM at upload.  The caller adds the
    per-wavemap ``data``, ``U``, ``slog_pdet``, ``nsamples``, ``hyper_idx`` and ``station_idx`` (they come from the
    reference's datasets / Covariance objects)."""
    out = []
    for mapid, comps in discover_seismic_libraries(gfpath, slip_vars, crust_ind):
        cfg0, G, tmins = None, {}, None
        for comp in slip_vars:
            cfg = read_library_config(os.path.join(gfpath, comps[comp] + ".yaml"))
            G[comp] = np.load(os.path.join(gfpath, comps[comp] + ".traces.npy"), mmap_mode="r", allow_pickle=False)
            if tuple(G[comp].shape) != tuple(int(x) for x in cfg["dimensions"]):
                raise GFLibraryError("%s: traces shape %s != config dimensions %s" % (comps[comp], G[comp].shape, cfg["dimensions"]))
            if cfg0 is None:
                cfg0 = cfg
                tmins = np.load(os.path.join(gfpath, comps[comp] + ".times.npy"), allow_pickle=False)
            else:
                for key in ("dimensions", "duration_min", "duration_sampling", "starttime_min", "starttime_sampling"):
                    if cfg[key] != cfg0[key]:
                        raise GFLibraryError("libraries of wavemap %s differ in %s" % (mapid, key))
        nt, _, ndur, nst, ns = (int(x) for x in cfg0["dimensions"])
        out.append(dict(mapid=mapid, nt=nt, ns=ns, ndur=ndur, nst=nst, G=G, tmins=tmins, interpolation=interpolation,
                        dur_min=float(cfg0["duration_min"]), dur_step=float(cfg0["duration_sampling"]),
                        st_min=float(cfg0["starttime_min"]), st_step=float(cfg0["starttime_sampling"])))
    if not out:
        raise GFLibraryError("no seismic GF libraries for %s found in %s" % (list(slip_vars), gfpath))
    return out
