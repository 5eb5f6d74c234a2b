"""CPU-side checks of the drop-in boundary: libdeft4cu.so loads without a GPU and exports every symbol that
include/deft4cu.h declares; compute calls fail loudly (no CPU fallback)."""
import os
import re

import pytest

from conftest import ROOT


def _declared_symbols():
    with open(os.path.join(ROOT, "include", "deft4cu.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(deft4cu_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_are_exported_and_bound():
    import __graft_entry__
    __graft_entry__.build()
    from deft4j_b200 import _native
    L = _native.load()
    declared = _declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(L, name), name
    # the ctypes prototypes cover exactly the header
    assert sorted(_native.SYMBOLS) == declared


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from deft4j_b200 import _native, DeflateStream, optimise_batch, Deft
    with pytest.raises(_native.Deft4cuError):
        DeflateStream().parse(b"\x03\x00")
    with pytest.raises(_native.Deft4cuError):
        optimise_batch([b"\x03\x00"])
    with pytest.raises(_native.Deft4cuError):
        Deft.optimiseDeflateStream(b"\x03\x00")
    # the native file front-ends sit on the same device entry: no device, no result
    import workloads as W
    from deft4j_b200.container import optimise_png_files, optimise_zip_files
    with pytest.raises(_native.Deft4cuError):
        optimise_png_files(W.c3_png_files(1))
    with pytest.raises(_native.Deft4cuError):
        optimise_zip_files([W.c4_zip_archive(2)])


def test_product_does_not_import_the_oracle():
    """Nothing under deft4j_b200/ may reference oracle/ (the oracle is test infrastructure)."""
    pkg = os.path.join(ROOT, "deft4j_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".sh")):
                with open(os.path.join(dirpath, fn)) as f:
                    text = f.read()
                assert "oracle_lib" not in text and "libdeft_oracle" not in text and "deft_oracle.h" not in text, fn
