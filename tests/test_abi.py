"""The C-ABI library loads and exports every symbol include/himut_b200.h declares; the
ctypes layouts match the compiled structs.  No compute calls: CPU only."""
import ctypes as C
import os
import re

import pytest

from himut_b200 import abi, lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "himut_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hm_[a-z_]+)\s*\(", text)))


def test_exports_every_declared_symbol():
    L = lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), "libhimut_b200.so does not export %s" % n
    assert sorted(lib.EXPORTS) == names


def test_struct_layouts():
    L = lib.load()
    assert L.hm_abi_version() == 1
    assert L.hm_abi_sizeof(0) == C.sizeof(abi.hm_read_batch)
    assert L.hm_abi_sizeof(1) == C.sizeof(abi.hm_chunk)
    assert L.hm_abi_sizeof(2) == C.sizeof(abi.hm_params)
    assert L.hm_abi_sizeof(3) == C.sizeof(abi.hm_site_record) == abi.SITE_DTYPE.itemsize


def test_no_cpu_fallback():
    """without a CUDA device the product refuses to run instead of falling back"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(lib.HimutError) as e:
        lib.Context(0)
    assert e.value.code == abi.HM_ERR_NO_DEVICE


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "himut_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".c", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("the CPU oracle (oracle/himut_oracle.c), which is test", ""), f
