"""The C-ABI library builds, loads, and exports every symbol include/ss_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes
import os
import re

from conftest import ROOT
from smartstartcontinuous_b200 import _lib, build


def _header_functions():
    src = open(os.path.join(ROOT, "include", "ss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ss_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_header_symbols():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = _header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), "libss_b200.so does not export %s" % n


def test_binding_table_covers_header():
    assert sorted(_lib.SIGNATURES) == _header_functions()


def test_create_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.ss_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value
    assert b"CUDA" in lib.ss_last_error(None) or b"device" in lib.ss_last_error(None)
