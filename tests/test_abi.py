"""CPU tests of the drop-in boundary: the C-ABI library builds, loads and exports every declared symbol.
No compute calls (there is no GPU here); the product must fail loudly without a device."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "beatgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(beatgpu_[a-z_0-9]+)\s*\(", src)))


@pytest.fixture(scope="module")
def built_lib():
    from beat_b200.build import build
    return build()


def test_header_symbols_all_exported(built_lib):
    lib = ctypes.CDLL(built_lib)
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), "libbeatgpu.so does not export %s" % n


def test_binding_covers_header(built_lib):
    from beat_b200 import lib as L
    assert sorted(L._SIGNATURES) == _declared_symbols()
    L.load()
    assert L.load().beatgpu_version() == 100


def test_layout_struct_matches_header():
    from beat_b200.lib import Layout
    # 2 + 3 + 9 int32 fields
    assert ctypes.sizeof(Layout) == 4 * 14
    from beat_b200.lib import GeomLayout
    assert ctypes.sizeof(GeomLayout) == 4 * 15          # beatgpu_geom_layout: 15 int32 fields


def test_no_cpu_fallback_without_device(built_lib):
    """Without a CUDA device the context refuses to exist -- nothing routes to a CPU path."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    from beat_b200.lib import BeatGpuError, Context
    with pytest.raises(BeatGpuError, match="no CPU fallback"):
        Context(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: no module of the package may import it."""
    pkg = os.path.join(ROOT, "beat_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M), f
                assert "libfsport" not in text, f


def test_integration_doc_lists_every_entry_point():
    """INTEGRATION.md is the reference-side binding guide: every exported entry must appear there (host/_dev pairs may
    share a row)."""
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    missing = [n for n in _declared_symbols() if n not in doc and n.replace("_dev", "") not in doc]
    assert not missing, missing


def test_parameter_matrix_width_is_checked_before_the_c_call():
    """The C entries trust [B, n_params]; the binding refuses any other width (no device needed to check that)."""
    from beat_b200.lib import Context
    c = Context.__new__(Context)
    c._n_params, c._geom_n_params = 7, 10
    c.n_outputs = lambda: 3
    import numpy as np
    for fn, args in ((c.ffi_loglike_batch, (np.zeros((4, 6)),)), (c.ffi_synthetics_batch, (0, np.zeros((4, 8)), 2, 5)),
                     (c.geom_loglike_batch, (np.zeros((4, 9)),)), (c.geom_synthetics_batch, (0, np.zeros(10), 2, 5))):
        with pytest.raises(ValueError, match="q must be"):
            fn(*args)
