"""CPU: the C-ABI library loads and exports every symbol include/radad_flat.h declares; without a GPU the
product fails loudly (no CPU fallback) instead of computing anything."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, Cfg

HEADER = os.path.join(ROOT, "include", "radad_flat.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rdb_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_bound_and_exported(pkg):
    from importlib import import_module
    cabi = import_module(pkg.__name__ + "._cabi")
    names = _declared()
    assert len(names) >= 25
    assert sorted(cabi.SIGNATURES) == names, "ctypes SIGNATURES must mirror include/radad_flat.h exactly"
    lib = ctypes.CDLL(cabi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in radad_flat.h but not exported by libradad_flat.so"
    assert cabi.load().rdb_abi_version() == cabi.ABI_VERSION


def test_header_cites_reference_for_every_entry_point():
    src = open(HEADER).read()
    assert src.count("vector_database.py:") >= 10 and "pipeline.py:503" in src


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        return False


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_fails_loudly(pkg, tmp_path):
    lib = pkg.load_native()
    h = ctypes.c_void_p()
    rc = lib.rdb_create(8, 0, 0, -1, 0, ctypes.byref(h))
    assert rc != 0 and not h
    assert b"no CPU fallback" in lib.rdb_last_error(None)
    with pytest.raises(RuntimeError):
        pkg.FlatIndex(8)
    with pytest.raises(RuntimeError):
        pkg.VectorDatabase(Cfg(tmp_path / "x"))      # the reference would silently fall back to CPU faiss


def test_product_never_imports_oracle():
    pk = os.path.join(ROOT, "radad-retrievalaugmenteddeepfakeaudiodetection_b200")
    for dirpath, _, files in os.walk(pk):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "flat_oracle" not in txt and "import oracle" not in txt and "from oracle" not in txt, f
