"""The C-ABI library loads on a CPU-only box and exports every symbol include/somcb.h declares
(no compute calls here)."""
import ctypes
import os
import re

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "somcb.h")


def _declared():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(som_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_expected_entry_points():
    names = _declared()
    for must in ("som_bmu_nchw_f32", "som_accumulate_nchw_f32", "som_filter_f32", "som_histogram_i64",
                 "som_merge_candidates", "som_adam_f32", "som_quantize_nchw_f32",
                 "som_prepare_codebook_f32", "som_last_error", "som_version"):
        assert must in names


def test_library_exports_every_declared_symbol():
    from somcb import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declared():
        assert hasattr(lib, name), f"{name} declared in somcb.h but not exported"


def test_binding_table_matches_header():
    from somcb import _lib
    assert sorted(_lib.SIGNATURES) == _declared()
    lib = _lib.load()
    from somcb import SOM_ABI_VERSION
    assert lib.som_version() == SOM_ABI_VERSION == 2
    assert lib.som_last_error() is not None


def test_variant_rule_is_static_and_host_only():
    from somcb import _lib
    lib = _lib.load()
    v = lib.som_bmu_pick_variant(512, 64, 1024)
    assert v in (_lib.SOM_BMU_FFMA, _lib.SOM_BMU_TC3X)
    assert lib.som_bmu_pick_variant(512, 64, 1024) == v
