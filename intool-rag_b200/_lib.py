"""ctypes binding of libhr_b200.so (include/hr_b200.h).  Fails loudly: no fallback of any kind."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libhr_b200.so")

HR_OK = 0
METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
STORAGE_F32 = 0
STORAGE_BF16 = 1
STORAGE_F32_SHADOW16 = 2
MODE_AUTO = 0
MODE_EXACT_SIMT = 1
FUSE_WEIGHTED = 0
FUSE_RRF = 1
IDF_LUCENE = 0
IDF_OKAPI = 1
MAX_K = 2048


class ScanStats(C.Structure):
    _fields_ = [("launches", C.c_int64), ("flagged", C.c_int64), ("overflow", C.c_int64),
                ("scan_ms", C.c_float), ("total_ms", C.c_float), ("mode_used", C.c_int),
                ("list_len", C.c_int), ("grid", C.c_int), ("deeper", C.c_int)]


# every symbol include/hr_b200.h declares: name -> (restype, argtypes)
_p = C.c_void_p
SYMBOLS = {
    "hr_last_error": (C.c_char_p, []),
    "hr_version": (C.c_int, []),
    "hr_source_hash": (C.c_char_p, []),
    "hr_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "hr_launch_count": (C.c_int64, []),
    "hr_set_option": (C.c_int, [C.c_char_p, C.c_int]),
    "hr_index_create": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(_p)]),
    "hr_index_destroy": (C.c_int, [_p]),
    "hr_index_reserve": (C.c_int, [_p, C.c_int64]),
    "hr_index_add": (C.c_int, [_p, _p, C.c_int64, C.c_int, _p]),
    "hr_index_reset": (C.c_int, [_p]),
    "hr_index_ntotal": (C.c_int64, [_p]),
    "hr_index_d": (C.c_int, [_p]),
    "hr_index_metric": (C.c_int, [_p]),
    "hr_index_storage": (C.c_int, [_p]),
    "hr_index_set_id_base": (C.c_int, [_p, C.c_int64]),
    "hr_index_set_mode": (C.c_int, [_p, C.c_int]),
    "hr_index_reconstruct": (C.c_int, [_p, C.c_int64, C.c_int64, _p]),
    "hr_index_search": (C.c_int, [_p, _p, C.c_int64, C.c_int, _p, _p, C.c_int, _p]),
    "hr_index_last_stats": (C.c_int, [_p, C.POINTER(ScanStats)]),
    "hr_index_debug_dump": (C.c_int, [_p, C.c_int64, _p, _p, _p, _p, _p]),
    "hr_index_save": (C.c_int, [_p, C.c_char_p]),
    "hr_index_load": (C.c_int, [C.c_char_p, C.c_int, C.c_int, C.POINTER(_p)]),
    "hr_bm25_create": (C.c_int, [_p, _p, _p, _p, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_int,
                                 C.c_int64, C.c_double, _p, C.c_int, C.c_int, _p, C.POINTER(_p)]),
    "hr_bm25_create_from_tokens": (C.c_int, [_p, _p, C.c_int64, C.c_int64, C.c_int64, C.c_float, C.c_float, C.c_int,
                                             C.c_int64, C.c_double, _p, C.c_int, C.c_int, _p, C.POINTER(_p)]),
    "hr_bm25_save": (C.c_int, [_p, C.c_char_p]),
    "hr_bm25_load": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_p)]),
    "hr_bm25_destroy": (C.c_int, [_p]),
    "hr_bm25_ndocs": (C.c_int64, [_p]),
    "hr_bm25_vocab": (C.c_int64, [_p]),
    "hr_bm25_nnz": (C.c_int64, [_p]),
    "hr_bm25_set_id_base": (C.c_int, [_p, C.c_int64]),
    "hr_bm25_search": (C.c_int, [_p, _p, _p, C.c_int64, C.c_int64, C.c_int, _p, _p, C.c_int, _p, C.POINTER(C.c_int64)]),
    "hr_merge_topk": (C.c_int, [_p, _p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, _p, _p, C.c_int, _p]),
    "hr_fuse": (C.c_int, [_p, _p, _p, _p, _p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                          C.c_float, _p, _p, C.c_int, _p]),
    "hr_rank_pages": (C.c_int, [_p, _p, C.c_int64, C.c_int, C.c_int, _p, C.c_int64, C.c_int64, C.c_int, _p, _p, _p,
                                C.c_int, _p]),
    "hr_candidates": (C.c_int, [_p, _p, _p, _p, _p, C.c_int64, C.c_int64, C.c_int, _p, _p, _p, _p, _p]),
    "hr_merge_fuse_lists": (C.c_int, [_p, _p, _p, _p, _p, C.c_int, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_float, _p, _p, _p]),
    "hr_comm_unique_id": (C.c_int, [_p]),
    "hr_comm_init": (C.c_int, [_p, C.c_int, C.c_int, C.c_int, C.POINTER(_p)]),
    "hr_comm_destroy": (C.c_int, [_p]),
    "hr_comm_rank": (C.c_int, [_p]),
    "hr_comm_world": (C.c_int, [_p]),
    "hr_retrieve_sharded": (C.c_int, [_p, _p, _p, _p, _p, _p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_float, C.c_float, _p, _p, C.c_int, _p]),
    "hr_retrieve": (C.c_int, [_p, _p, _p, _p, _p, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float,
                              _p, _p, C.c_int, _p]),
}

_lib = None


def lib() -> C.CDLL:
    """Load libhr_b200.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C intool-rag_b200/csrc`.  intool-rag_b200 has no CPU fallback.")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(l, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


SOURCE_FILES = ["hr_api.cu", "common.cuh", "dense_exact.cuh", "dense_scan_tc.cuh", "bm25.cuh", "bm25_sweep.cuh",
                "bm25_build.cuh", "fuse.cuh", os.path.join("..", "..", "include", "hr_b200.h")]


def tree_source_hash() -> str:
    """sha256 over the library's sources in the Makefile's order (csrc/Makefile: HASH), first 16 hex digits."""
    import hashlib
    h = hashlib.sha256()
    for f in SOURCE_FILES:
        with open(os.path.join(_HERE, "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


def built_source_hash() -> str:
    return lib().hr_source_hash().decode()


def last_error() -> str:
    return lib().hr_last_error().decode("utf-8", "replace")


def check(rc: int) -> None:
    if rc != HR_OK:
        raise RuntimeError(f"hr_b200 error {rc}: {last_error()}")


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().hr_device_count(C.byref(n))
    return n.value if rc == HR_OK else 0


def require_gpu() -> None:
    if device_count() <= 0:
        raise RuntimeError("intool-rag_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback: "
                           + last_error())


def launch_count() -> int:
    return int(lib().hr_launch_count())


def set_option(name: str, value: int) -> None:
    """Tuning / diagnostic knob of the library (include/hr_b200.h: hr_set_option)."""
    check(lib().hr_set_option(name.encode(), int(value)))


def current_stream_ptr(device: int | None = None) -> int:
    """Raw cudaStream_t of torch's current stream (0 = legacy default stream if torch is absent)."""
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_stream(device).cuda_stream)
    except Exception:
        pass
    return 0
