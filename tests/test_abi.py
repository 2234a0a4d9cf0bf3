"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports exactly the symbols that
include/pfd_b200.h declares, and fails loudly (no CPU fallback) when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g

    g.build()
    from pyflwdir_b200 import _lib

    return _lib


def _header_functions():
    src = open(os.path.join(ROOT, "include", "pfd_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pfd_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(built):
    names = _header_functions()
    assert len(names) >= 25
    lib = C.CDLL(built.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/pfd_b200.h but not exported by libpfd_b200.so"
    # the ctypes table binds every declared function, and nothing that is not declared
    assert sorted(built.SYMBOLS) == names


def test_version_and_status_strings(built):
    l = built.lib()
    assert b"sm_100a" in l.pfd_version()
    assert l.pfd_status_string(0) == b"ok"
    assert l.pfd_status_string(3) == b"invalid D8 data"


def test_enums_match_header(built):
    src = open(os.path.join(ROOT, "include", "pfd_b200.h")).read()
    for name, val in re.findall(r"(PFD_(?:I8|U8|I16|U16|I32|U32|I64|U64|F32|F64))\s*=\s*(\d+)", src):
        np_name = {"I": "int", "U": "uint", "F": "float"}[name[4]] + name[5:]
        assert built.DTYPES[np.dtype(np_name)] == int(val)
    for name, val in re.findall(r"(PFD_ARR_[A-Z0-9_]+)\s*=\s*(\d+)", src):
        assert getattr(built, name[4:]) == int(val)
    for name, val in re.findall(r"(PFD_ERR_[A-Z0-9_]+)\s*=\s*(\d+)", src):
        assert getattr(built, name[4:]) == int(val)


def test_no_cpu_fallback(built):
    """Without a CUDA device every compute path must raise, never silently compute on the CPU."""
    if built.device_count() > 0:
        pytest.skip("a GPU is present")
    import pyflwdir_b200 as pfb

    h = C.c_void_p()
    assert built.lib().pfd_create(0, C.byref(h)) == built.ERR_CUDA
    assert b"no CPU fallback" in built.lib().pfd_last_error(None)
    with pytest.raises(built.PfdError, match="no usable CUDA device"):
        pfb.from_array(np.zeros((8, 8), dtype=np.uint8), ftype="d8")


def test_product_does_not_import_oracle():
    """The product package must never route through the oracle (test infrastructure)."""
    pkg = os.path.join(ROOT, "pyflwdir_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f
                assert "libpfd_oracle" not in txt and "pfd_oracle" not in txt, f
