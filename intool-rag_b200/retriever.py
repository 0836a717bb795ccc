"""retrieve(query_embeddings, query_tokens, top_k) — the hybrid query hot path.

This is the surface the reference advertises at ``rag/query/retriever.py``
(/root/reference/README.md:90) but never shipped; it sits beside
``PageLevelRetriever.retrieve_chunks`` (/root/reference/rag/query/page_retriever.py:92-143), whose
``search_faiss_by_vector(..., limit=50)`` call it generalises to a batch with BM25 + fusion.

Defaults are the reference's dead config constants (rag/config.py:41-45): VECTOR_WEIGHT=0.7,
BM25_WEIGHT=0.3, RETRIEVAL_TOP_K=10, HYBRID_SEARCH_ENABLED=true; candidate depth per modality is
max(top_k, 50).  Dense + BM25 + fusion run as one C-ABI call (hr_retrieve) on one stream.
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .bm25 import BM25Index, query_csr
from .config import config
from .faiss import Index

_MODES = {"weighted": _lib.FUSE_WEIGHTED, "rrf": _lib.FUSE_RRF}


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def candidate_depth(top_k: int) -> int:
    return max(int(top_k), int(config.CANDIDATE_DEPTH))


class HybridRetriever:
    def __init__(self, index: Index, bm25: BM25Index | None = None, vector_weight: float | None = None,
                 bm25_weight: float | None = None, fusion: str | None = None, hybrid: bool | None = None):
        self.index = index
        self.bm25 = bm25
        self.vector_weight = config.VECTOR_WEIGHT if vector_weight is None else float(vector_weight)
        self.bm25_weight = config.BM25_WEIGHT if bm25_weight is None else float(bm25_weight)
        self.fusion = fusion or config.FUSION
        self.hybrid = config.HYBRID_SEARCH_ENABLED if hybrid is None else bool(hybrid)
        if self.fusion not in _MODES:
            raise ValueError(f"unknown fusion {self.fusion!r}")

    def retrieve(self, query_embeddings, query_tokens=None, top_k: int | None = None, k_c: int | None = None):
        """query_embeddings: float32 [nq, d] (numpy, or torch CUDA — stays on device);
        query_tokens: ragged list of token-id lists or CSR (indptr, terms); None = dense only.
        Returns (scores float32[nq, top_k] fused, ids int64[nq, top_k], -1 padded)."""
        top_k = config.RETRIEVAL_TOP_K if top_k is None else int(top_k)
        kc = candidate_depth(top_k) if k_c is None else int(k_c)
        use_bm = self.hybrid and self.bm25 is not None and query_tokens is not None
        bm_h = self.bm25._h if use_bm else None
        L = _lib.lib()
        if _is_torch_cuda(query_embeddings):
            import torch
            q = query_embeddings.to(torch.float32).contiguous()
            if q.dim() != 2 or q.shape[1] != self.index.d:
                raise AssertionError(f"retrieve: expected [nq, {self.index.d}] embeddings, got {tuple(q.shape)}")
            nq = q.shape[0]
            qi = qt = None
            if use_bm:
                ip, tm = query_csr(query_tokens)
                if not _is_torch_cuda(ip):
                    ip = torch.from_numpy(np.ascontiguousarray(ip)).to(q.device)
                    tm = torch.from_numpy(np.ascontiguousarray(tm)).to(q.device)
                qi, qt = ip.to(torch.int32).contiguous(), tm.to(torch.int32).contiguous()
                if qi.shape[0] != nq + 1:
                    raise AssertionError("retrieve: query_tokens and query_embeddings disagree on nq")
            S = torch.empty((nq, top_k), dtype=torch.float32, device=q.device)
            I = torch.empty((nq, top_k), dtype=torch.int64, device=q.device)
            _lib.check(L.hr_retrieve(self.index._h, bm_h, q.data_ptr(), qi.data_ptr() if use_bm else None,
                                     qt.data_ptr() if use_bm else None, nq, int(qt.numel()) if use_bm else 0,
                                     top_k, kc, _MODES[self.fusion],
                                     self.vector_weight, self.bm25_weight, S.data_ptr(), I.data_ptr(), 1,
                                     _lib.current_stream_ptr(self.index.device)))
            return S, I
        q = np.ascontiguousarray(query_embeddings, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.index.d:
            raise AssertionError(f"retrieve: expected [nq, {self.index.d}] embeddings, got {q.shape}")
        nq = q.shape[0]
        qi = qt = None
        if use_bm:
            qi, qt = query_csr(query_tokens)
            if len(qi) != nq + 1:
                raise AssertionError("retrieve: query_tokens and query_embeddings disagree on nq")
        S = np.empty((nq, top_k), dtype=np.float32)
        I = np.empty((nq, top_k), dtype=np.int64)
        _lib.check(L.hr_retrieve(self.index._h, bm_h, q.ctypes.data, qi.ctypes.data if use_bm else None,
                                 qt.ctypes.data if use_bm else None, nq, int(qt.size) if use_bm else 0, top_k, kc, _MODES[self.fusion],
                                 self.vector_weight, self.bm25_weight, S.ctypes.data, I.ctypes.data, 0,
                                 _lib.current_stream_ptr(self.index.device)))
        return S, I


_default: HybridRetriever | None = None


def set_default_retriever(r: HybridRetriever | None) -> None:
    global _default
    _default = r


def retrieve(query_embeddings, query_tokens=None, top_k: int | None = None):
    """Module-level surface named by BASELINE.json: retrieve(query_embeddings, query_tokens, top_k)."""
    if _default is None:
        raise RuntimeError("no retriever configured: call set_default_retriever(HybridRetriever(index, bm25))")
    return _default.retrieve(query_embeddings, query_tokens, top_k)
