"""The C-ABI library loads and exports every symbol include/cytopath_b200.h declares (no GPU needed)."""
import ctypes
import os
import re

import pytest

from cytopath_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from cytopath_b200.build import build_library
        build_library()
    return _lib.load()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "cytopath_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cyp_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "missing export: " + name
    assert sorted(_lib.EXPORTS) == names


def test_version_and_error_string(lib):
    assert lib.cyp_version() == 100
    assert isinstance(lib.cyp_last_error(), bytes)


def test_no_gpu_is_a_loud_failure(lib):
    n = ctypes.c_int(-1)
    rc = lib.cyp_device_count(ctypes.byref(n))
    if rc == 0 and n.value > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _lib.require_device(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "cytopath_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "liboracle" not in src, f
