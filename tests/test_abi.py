"""CPU tests of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and exports every symbol that
include/rabi_b200.h declares.  No compute calls are made (there is no GPU here)."""
import ctypes
import os
import re

import pytest

from rabi_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "rabi_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(rabi_[a-z0-9_]+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/rabi_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


def test_build_targets_sm_100a(lib):
    assert b"sm_100a" in lib.rabi_build_info()
    assert "compute_100a" in " ".join(_lib.NVCC_FLAGS) and "-lineinfo" in _lib.NVCC_FLAGS


def test_create_validates_arguments_and_has_no_cpu_fallback(lib):
    import torch

    h = ctypes.c_void_p()
    hidden = (ctypes.c_int * 2)(3, 3)
    rc = lib.rabi_maf_create(ctypes.byref(h), 0, 3, 4, 10, 2, hidden, -7.0, 5.0)
    assert rc != 0 and b"Hidden dimension must not be less than input dimension" in lib.rabi_last_error(None)
    if not torch.cuda.is_available():
        hidden = (ctypes.c_int * 2)(16, 16)
        rc = lib.rabi_maf_create(ctypes.byref(h), 0, 3, 4, 10, 2, hidden, -7.0, 5.0)
        assert rc != 0 and b"no CPU fallback" in lib.rabi_last_error(None)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "rabi_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
