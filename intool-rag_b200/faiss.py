"""faiss-shaped module over the B200 flat index (drop-in for the reference's `import faiss`).

Covers exactly the surface the reference touches (SURVEY.md §8b surface 1):
``IndexFlatL2(d)`` (/root/reference/rag/storage/faiss_index.py:123), ``.add`` (:124),
``.search`` (:83, rag/agent/search_engine.py:45), ``.d`` / ``.ntotal`` (:58-59,97,103,126),
``read_index`` (:54) and ``write_index`` (:133) — plus ``IndexFlatIP``, ``reset``, ``reconstruct``.
Install with ``sys.modules["faiss"] = intool_rag_b200.faiss``.

All arithmetic runs in libhr_b200.so on the GPU; there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _lib
from .config import default_device

METRIC_INNER_PRODUCT = _lib.METRIC_INNER_PRODUCT
METRIC_L2 = _lib.METRIC_L2
_STORAGE = {"f32": _lib.STORAGE_F32, "fp32": _lib.STORAGE_F32, "float32": _lib.STORAGE_F32,
            "bf16": _lib.STORAGE_BF16, "bfloat16": _lib.STORAGE_BF16,
            "f32+bf16": _lib.STORAGE_F32_SHADOW16, "f32_shadow16": _lib.STORAGE_F32_SHADOW16}
_STORAGE_NAME = {_lib.STORAGE_F32: "f32", _lib.STORAGE_BF16: "bf16", _lib.STORAGE_F32_SHADOW16: "f32+bf16"}


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


class Index:
    """Base of the flat indexes (faiss.Index surface used by the reference)."""

    is_trained = True

    def __init__(self, handle: int, device: int):
        self._h = C.c_void_p(handle)
        self.device = device

    # -- attributes faiss exposes ------------------------------------------------------------
    @property
    def d(self) -> int:
        return int(_lib.lib().hr_index_d(self._h))

    @property
    def ntotal(self) -> int:
        return int(_lib.lib().hr_index_ntotal(self._h))

    @property
    def metric_type(self) -> int:
        return int(_lib.lib().hr_index_metric(self._h))

    @property
    def storage(self) -> str:
        return _STORAGE_NAME[int(_lib.lib().hr_index_storage(self._h))]

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().hr_index_destroy(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    # -- add -------------------------------------------------------------------------------------
    def reserve(self, n_rows: int) -> None:
        _lib.check(_lib.lib().hr_index_reserve(self._h, int(n_rows)))

    def add(self, x) -> None:
        """x: float32 [n, d] (numpy / anything array-like, or a torch CUDA tensor)."""
        if _is_torch_cuda(x):
            return self.add_device(x)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise AssertionError("add: x must be a 2-D array [n, d]")
        if x.shape[1] != self.d:
            raise AssertionError(f"add: x.shape[1] == {x.shape[1]} != index.d == {self.d}")
        _lib.check(_lib.lib().hr_index_add(self._h, x.ctypes.data, x.shape[0], 0, None))

    def add_device(self, x) -> None:
        import torch
        if x.dim() != 2 or x.shape[1] != self.d:
            raise AssertionError(f"add: expected a [n, {self.d}] tensor, got {tuple(x.shape)}")
        if x.device.index != self.device:
            raise RuntimeError(f"add: tensor on cuda:{x.device.index}, index on cuda:{self.device}")
        x = x.to(torch.float32).contiguous()
        st = _lib.current_stream_ptr(self.device)
        _lib.check(_lib.lib().hr_index_add(self._h, x.data_ptr(), x.shape[0], 1, st))

    def reset(self) -> None:
        _lib.check(_lib.lib().hr_index_reset(self._h))

    # -- search ----------------------------------------------------------------------------------
    def search(self, x, k: int):
        """(D float32[nq,k], I int64[nq,k]); L2: squared distances ascending, IP: inner products
        descending, padding I=-1 / D=+-FLT_MAX.  numpy in -> numpy out; torch CUDA in -> torch CUDA out."""
        if _is_torch_cuda(x):
            return self.search_device(x, k)
        x = np.ascontiguousarray(x, dtype=np.float32)
        if x.ndim != 2:
            raise AssertionError("search: x must be a 2-D array [nq, d]")
        if x.shape[1] != self.d:
            raise AssertionError(f"search: x.shape[1] == {x.shape[1]} != index.d == {self.d}")
        k = int(k)
        if k <= 0:
            raise AssertionError("search: k must be > 0")
        nq = x.shape[0]
        D = np.empty((nq, k), dtype=np.float32)
        I = np.empty((nq, k), dtype=np.int64)
        _lib.check(_lib.lib().hr_index_search(self._h, x.ctypes.data, nq, k, D.ctypes.data, I.ctypes.data, 0,
                                              _lib.current_stream_ptr(self.device)))
        return D, I

    def search_device(self, x, k: int):
        import torch
        if x.dim() != 2 or x.shape[1] != self.d:
            raise AssertionError(f"search: expected a [nq, {self.d}] tensor, got {tuple(x.shape)}")
        if x.device.index != self.device:
            raise RuntimeError(f"search: tensor on cuda:{x.device.index}, index on cuda:{self.device}")
        x = x.to(torch.float32).contiguous()
        nq, k = x.shape[0], int(k)
        D = torch.empty((nq, k), dtype=torch.float32, device=x.device)
        I = torch.empty((nq, k), dtype=torch.int64, device=x.device)
        _lib.check(_lib.lib().hr_index_search(self._h, x.data_ptr(), nq, k, D.data_ptr(), I.data_ptr(), 1,
                                              _lib.current_stream_ptr(self.device)))
        return D, I

    # -- misc ------------------------------------------------------------------------------------
    def reconstruct(self, i: int) -> np.ndarray:
        return self.reconstruct_n(int(i), 1)[0]

    def reconstruct_n(self, i0: int = 0, n: int | None = None) -> np.ndarray:
        n = self.ntotal - i0 if n is None else int(n)
        out = np.empty((n, self.d), dtype=np.float32)
        _lib.check(_lib.lib().hr_index_reconstruct(self._h, int(i0), n, out.ctypes.data))
        return out

    def set_id_base(self, base: int) -> None:
        _lib.check(_lib.lib().hr_index_set_id_base(self._h, int(base)))

    def set_mode(self, mode: str) -> None:
        """'auto' (tensor-core filter + exact re-score + certified fallback) or 'exact' (SIMT scan)."""
        m = {"auto": _lib.MODE_AUTO, "exact": _lib.MODE_EXACT_SIMT}[mode]
        _lib.check(_lib.lib().hr_index_set_mode(self._h, m))

    def stats(self) -> dict:
        s = _lib.ScanStats()
        _lib.check(_lib.lib().hr_index_last_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in _lib.ScanStats._fields_}


class IndexFlat(Index):
    def __init__(self, d: int, metric: int = METRIC_L2, storage: str | None = None, device: int | None = None):
        _lib.require_gpu()
        dev = default_device() if device is None else int(device)
        storage = storage or os.getenv("HR_STORAGE", "f32")
        if storage not in _STORAGE:
            raise ValueError(f"unknown storage {storage!r}")
        h = C.c_void_p()
        _lib.check(_lib.lib().hr_index_create(int(d), int(metric), _STORAGE[storage], dev, C.byref(h)))
        super().__init__(h.value, dev)


class IndexFlatL2(IndexFlat):
    def __init__(self, d: int, **kw):
        super().__init__(d, METRIC_L2, **kw)


class IndexFlatIP(IndexFlat):
    def __init__(self, d: int, **kw):
        super().__init__(d, METRIC_INNER_PRODUCT, **kw)


def write_index(index: Index, path: str) -> None:
    """faiss.write_index for flat indexes: writes faiss's own "IxF2"/"IxFI" byte layout."""
    _lib.check(_lib.lib().hr_index_save(index._h, os.fsencode(str(path))))


def read_index(path: str, storage: str | None = None, device: int | None = None) -> Index:
    """faiss.read_index for flat index files (as written by faiss-cpu 1.7.4 or write_index above)."""
    _lib.require_gpu()
    dev = default_device() if device is None else int(device)
    storage = storage or os.getenv("HR_STORAGE", "f32")
    h = C.c_void_p()
    _lib.check(_lib.lib().hr_index_load(os.fsencode(str(path)), dev, _STORAGE[storage], C.byref(h)))
    return Index(h.value, dev)
