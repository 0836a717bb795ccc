"""ORACLE — test infrastructure only.  CPU definition of hybrid fusion + final top-k.

PARITY STATUS: **unpinned** — the reference implements no fusion.  The only
reference-side pins are the dead constants VECTOR_WEIGHT=0.7, BM25_WEIGHT=0.3,
RETRIEVAL_TOP_K=10, HYBRID_SEARCH_ENABLED=true (/root/reference/rag/config.py:41-45) and the
dense score transform clamp(1 - d/2, 0, 1) (/root/reference/rag/storage/faiss_index.py:86-88),
which are the defaults here.  Definitions follow SURVEY.md Appendix B.

Inputs per query: a dense candidate list and a sparse (BM25) candidate list, each of depth
k_c, best first, ids -1 padded.
  weighted: dense01 = clamp(sim, 0, 1)  (sim = inner product, or 1 - d/2 for an L2 index)
            bm25n   = bm25 / (best BM25 score of this query)   (0 if no term matched)
            fused   = w_vec * dense01 + w_bm25 * bm25n, a doc absent from a list gets 0 there
  rrf:      fused   = sum over lists containing the doc of 1 / (60 + rank), rank 1-based
Output: top_k of the union by (fused desc, id asc), -1 / 0.0 padded.
Computed in fp64 here; the CUDA path computes in fp32 and must agree to 1e-5 relative.
"""
from __future__ import annotations

from typing import Tuple

import numpy as np

VECTOR_WEIGHT = 0.7
BM25_WEIGHT = 0.3
RRF_K = 60.0


def dense_similarity(D: np.ndarray, metric_l2: bool) -> np.ndarray:
    D = np.asarray(D, dtype=np.float64)
    return 1.0 - D / 2.0 if metric_l2 else D


def fuse(dense_sim: np.ndarray, dense_ids: np.ndarray, bm25: np.ndarray, bm25_ids: np.ndarray,
         top_k: int, mode: str = "weighted", w_vec: float = VECTOR_WEIGHT,
         w_bm25: float = BM25_WEIGHT, bm25_max=None) -> Tuple[np.ndarray, np.ndarray]:
    """dense_sim/bm25: [nq, kc] float; *_ids: [nq, kc] int64 (-1 = padding).
    bm25_max: optional [nq] global per-query normaliser (sharded case); default = best of list."""
    nq = dense_ids.shape[0]
    out_s = np.zeros((nq, top_k), dtype=np.float32)
    out_i = np.full((nq, top_k), -1, dtype=np.int64)
    for q in range(nq):
        fused = {}
        if mode == "weighted":
            for s, i in zip(dense_sim[q], dense_ids[q]):
                if i >= 0:
                    fused[int(i)] = fused.get(int(i), 0.0) + w_vec * min(1.0, max(0.0, float(s)))
            valid = bm25_ids[q] >= 0
            mx = float(bm25_max[q]) if bm25_max is not None else (
                float(np.max(bm25[q][valid])) if valid.any() else 0.0)
            for s, i in zip(bm25[q], bm25_ids[q]):
                if i >= 0 and mx > 0:
                    fused[int(i)] = fused.get(int(i), 0.0) + w_bm25 * (float(s) / mx)
                elif i >= 0:
                    fused.setdefault(int(i), 0.0)
        elif mode == "rrf":
            for r, i in enumerate(dense_ids[q]):
                if i >= 0:
                    fused[int(i)] = fused.get(int(i), 0.0) + 1.0 / (RRF_K + r + 1)
            for r, i in enumerate(bm25_ids[q]):
                if i >= 0:
                    fused[int(i)] = fused.get(int(i), 0.0) + 1.0 / (RRF_K + r + 1)
        else:
            raise ValueError(mode)
        if not fused:
            continue
        ids = np.fromiter(fused.keys(), dtype=np.int64)
        sc = np.fromiter(fused.values(), dtype=np.float64)
        sc32 = sc.astype(np.float32)
        order = np.lexsort((ids, -sc32))[:top_k]
        out_i[q, :len(order)] = ids[order]
        out_s[q, :len(order)] = sc32[order]
    return out_s, out_i


def merge_shards(S: np.ndarray, I: np.ndarray, k: int, largest: bool = True,
                 pad_score: float = 0.0) -> Tuple[np.ndarray, np.ndarray]:
    """k-way merge of per-shard candidate lists: S,I [nq, shards*kc] -> best k by the
    documented order (score best-first, id asc); ids < 0 are padding."""
    nq = S.shape[0]
    out_s = np.full((nq, k), pad_score, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for q in range(nq):
        valid = np.nonzero(I[q] >= 0)[0]
        key = -S[q, valid] if largest else S[q, valid]
        order = np.lexsort((I[q, valid], key))[:k]
        sel = valid[order]
        out_i[q, :len(sel)] = I[q, sel]
        out_s[q, :len(sel)] = S[q, sel]
    return out_s, out_i
