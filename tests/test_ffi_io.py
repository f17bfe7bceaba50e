"""On-disk GF-library format (reference layout: .traces.npy / .times.npy / .yaml): CPU round trip + GPU load."""
import numpy as np
import pytest

from beat_b200 import ffi

GUTS_YAML = """# beat.ffi.SeismicGFLibrary YAML Config
--- !beat.SeismicGFLibraryConfig
component: uperp
event: !pf.Event
  lat: 10.0
  lon: 10.0
  time: 1970-01-01 00:00:00
  depth: 2000.0
crust_ind: 0
reference_sources:
- !beat.sources.RectangularSource
  lat: 10.0
  lon: 10.0
  depth: 2000.0
  stf: !pf.HalfSinusoidSTF
    duration: 0.0
wave_config: !beat.WaveformFitConfig
  name: any_P
  arrival_taper: !beat.heart.ArrivalTaper
    a: -15.0
    b: -10.0
    c: 50.0
    d: 55.0
starttime_sampling: 0.5
duration_sampling: 0.25
starttime_min: -5.0
duration_min: 0.5
dimensions: [3, 4, 2, 5, 6]
datatype: seismic
mapnumber: 1
"""


def test_read_guts_style_yaml(tmp_path):
    p = tmp_path / "seismic_uperp_any_P_1_0.yaml"
    p.write_text(GUTS_YAML)
    cfg = ffi.read_library_config(str(p))
    assert cfg["dimensions"] == [3, 4, 2, 5, 6] and cfg["component"] == "uperp"
    assert cfg["starttime_min"] == -5.0 and cfg["duration_sampling"] == 0.25
    assert cfg["wave_config"]["arrival_taper"]["c"] == 50.0
    with pytest.raises(IOError):
        ffi.read_library_config(str(tmp_path / "missing.yaml"))


def test_save_then_read_roundtrip(tmp_path):
    rng = np.random.default_rng(0)
    G = rng.standard_normal((3, 4, 2, 5, 6))
    prefix = ffi.save_gf_library(str(tmp_path), G, np.arange(3.0), component="uparr", wavename="any_P", mapnumber=0,
                                 duration_min=0.5, duration_sampling=0.25, starttime_min=-1.0, starttime_sampling=0.5)
    assert prefix == ffi.get_gf_prefix("seismic", "uparr", "any_P_0", 0) == "seismic_uparr_any_P_0_0"
    cfg = ffi.read_library_config(str(tmp_path / (prefix + ".yaml")))
    assert cfg["dimensions"] == list(G.shape) and cfg["wave_config"]["name"] == "any_P"
    back = np.load(str(tmp_path / (prefix + ".traces.npy")), mmap_mode="r")
    assert back.dtype == np.float64 and np.array_equal(back, G)


@pytest.mark.gpu
def test_load_gf_library_on_gpu(tmp_path):
    from oracle import ffi_oracle as O
    rng = np.random.default_rng(1)
    G = rng.standard_normal((3, 9, 3, 6, 20))
    prefix = ffi.save_gf_library(str(tmp_path), G, np.zeros(3), duration_min=0.5, duration_sampling=0.25,
                                 starttime_min=-1.0, starttime_sampling=0.5)
    gfs = ffi.load_gf_library(str(tmp_path), prefix, store_dtype="float64")
    assert (gfs.ntargets, gfs.npatches, gfs.ndurations, gfs.nstarttimes, gfs.nsamples) == G.shape
    d = rng.uniform(0.51, 0.99, 9)
    st = rng.uniform(-0.9, 1.4, (3, 9))
    u = rng.uniform(0, 2, 9)
    tidx = np.atleast_2d(np.arange(3)).T
    for interp in ("nearest_neighbor", "multilinear"):
        got = gfs.stack_all(d, st, u, targetidxs=tidx, interpolation=interp)
        ref = O.stack_all(G, d, st, u, 0.5, 0.25, -1.0, 0.5, interp)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12 * np.abs(ref).max())
    with pytest.raises(ValueError):
        ffi.load_gf_library(str(tmp_path), "geodetic_uparr_static_0")


def test_discover_project_libraries(tmp_path):
    rng = np.random.default_rng(2)
    shape = (2, 6, 2, 4, 8)
    for comp in ("uparr", "uperp"):
        for mapn in (0, 1):
            ffi.save_gf_library(str(tmp_path), rng.standard_normal(shape), np.zeros(2), component=comp, wavename="any_P",
                                mapnumber=mapn, duration_min=0.5, duration_sampling=0.25, starttime_min=-1.0, starttime_sampling=0.5)
    ffi.save_gf_library(str(tmp_path), rng.standard_normal(shape), np.zeros(2), component="uparr", wavename="any_S", mapnumber=0)
    found = ffi.discover_seismic_libraries(str(tmp_path), ("uparr", "uperp"))
    assert [m for m, _ in found] == ["any_P_0", "any_P_1"]          # any_S_0 lacks uperp
    wms = ffi.wavemaps_from_directory(str(tmp_path), ("uparr", "uperp"))
    assert len(wms) == 2 and wms[0]["nt"] == 2 and wms[0]["ns"] == 8 and wms[0]["st_min"] == -1.0
    assert isinstance(wms[0]["G"]["uperp"], np.memmap)
    with pytest.raises(ffi.GFLibraryError):
        ffi.wavemaps_from_directory(str(tmp_path), ("utens",))


@pytest.mark.gpu
def test_engine_from_project_directory(tmp_path):
    """Libraries written in the reference's on-disk layout, memory-mapped and streamed to HBM, give the same
    log-likelihoods as the in-memory problem."""
    from beat_b200 import synthetic
    from beat_b200.engine import BatchedFFILogLike
    prob = synthetic.make_problem(nt=4, subfaults=((4, 6, 2.0),), ns=32, ndur=4, seed=17)
    wm = prob["wavemaps"][0]
    for comp in prob["slip_vars"]:
        ffi.save_gf_library(str(tmp_path), wm["G"][comp], np.zeros(wm["nt"]), component=comp, wavename="any_P", mapnumber=0,
                            duration_min=wm["dur_min"], duration_sampling=wm["dur_step"], starttime_min=wm["st_min"],
                            starttime_sampling=wm["st_step"])
    disk = ffi.wavemaps_from_directory(str(tmp_path), prob["slip_vars"])[0]
    disk.update({k: wm[k] for k in ("data", "U", "slog_pdet", "nsamples", "hyper_idx", "station_idx")})
    Q = synthetic.draw_chains(prob, 9, seed=3)
    a = BatchedFFILogLike.from_problem(prob, store_dtype="float32")
    b = BatchedFFILogLike.from_problem(dict(prob, wavemaps=[disk]), store_dtype="float32")
    assert np.array_equal(a(Q)[0], b(Q)[0])
    a.close()
    b.close()
