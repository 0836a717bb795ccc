"""CPU: the C-ABI library loads, exports every symbol include/hr_b200.h declares, the ctypes table
covers the header, and compute entry points fail loudly without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re
import subprocess

import numpy as np
import pytest

import intool_rag_b200  # noqa: F401
from intool_rag_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hr_b200.h")


def header_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(hr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_header_symbol():
    lib = _lib.lib()
    syms = header_symbols()
    assert len(syms) >= 30
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in hr_b200.h but not exported by libhr_b200.so"


def test_ctypes_table_matches_header():
    assert sorted(_lib.SYMBOLS) == header_symbols()


def test_exports_are_plain_c_abi():
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    for s in header_symbols():
        assert s in exported, f"{s} is not an unmangled exported function"


def test_library_is_sm100a_with_tcgen05_and_tma():
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    sass = subprocess.run(["cuobjdump", "-sass", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "UTCHMMA" in sass or "UTCQMMA" in sass or re.search(r"UTC\w*MMA", sass), "no tcgen05.mma in SASS"
    assert "UTMALDG" in sass, "no TMA loads in SASS"
    assert "LDTM" in sass, "no tcgen05.ld in SASS"


def test_library_was_built_from_the_sources_in_the_tree():
    """*.so is git-ignored: the hash of the sources compiled into the binary must match the tree, so a stale
    library cannot pass for the committed code (build() rebuilds on a mismatch)."""
    assert _lib.built_source_hash() == _lib.tree_source_hash()


def test_version_and_launch_counter():
    assert _lib.lib().hr_version() >= 200
    assert _lib.launch_count() >= 0


@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    from intool_rag_b200 import faiss, bm25
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        faiss.IndexFlatL2(8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        faiss.read_index("/nonexistent")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        bm25.BM25Index.from_docs([[0, 1]], 4)
    h = C.c_void_p()
    rc = _lib.lib().hr_index_create(8, 1, 0, 0, C.byref(h))
    assert rc == -2 and not h.value and "CPU fallback" in _lib.last_error() or "cudaGetDeviceCount" in _lib.last_error()


def test_argument_validation_needs_no_gpu():
    h = C.c_void_p()
    L = _lib.lib()
    assert L.hr_index_create(0, 1, 0, 0, C.byref(h)) == -1
    assert L.hr_index_create(8, 7, 0, 0, C.byref(h)) == -1
    assert L.hr_index_create(8, 1, 9, 0, C.byref(h)) == -1
    assert L.hr_index_search(None, None, 1, 1, None, None, 0, None) == -1
    assert "null index" in _lib.last_error()
    assert L.hr_index_ntotal(None) == -1
    assert L.hr_index_destroy(None) == 0
