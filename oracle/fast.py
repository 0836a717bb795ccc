"""ORACLE — test infrastructure only.  The CPU baseline's FAST path: the same definitions as
oracle/{flat,bm25,fusion}.py, computed the way a CPU service would:

* dense: fp32 sgemm in database blocks + top-k — the algorithm faiss-cpu 1.7.4's IndexFlat uses for
  nq >= 20 (the reference calls it at /root/reference/rag/storage/faiss_index.py:83),
  here torch.mm (MKL/OpenBLAS, all host threads) + torch.topk with a running merge;
* BM25: one sparse product  (queries x terms, weights mult * idf)  @  (terms x docs, folded impacts)  with
  scipy.sparse, then torch.topk;
* fusion: oracle/fusion.py.

PARITY STATUS: unpinned by the reference (see oracle/flat.py); tests/test_oracle_golden.py checks this fast
path against the plain oracle.  Only ``tests/`` and ``bench.py``'s CPU arm import it.
"""
from __future__ import annotations

import time
from typing import Sequence, Tuple

import numpy as np

from . import flat, fusion


def dense_topk(x: np.ndarray, q: np.ndarray, k: int, metric_l2: bool, block: int = 65536) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k of fp32 scores (sgemm blocks).  Returns (D float32[nq,k], I int64[nq,k]), -1 / +-FLT_MAX padded.
    Ties: torch.topk's order (the plain oracle applies (score, id asc); exact fp32 ties are absent from the
    synthetic data this path is timed on)."""
    import torch
    xt = torch.from_numpy(np.ascontiguousarray(x, dtype=np.float32))
    qt = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32))
    nq, n = qt.shape[0], xt.shape[0]
    pad = float(flat.FLT_MAX) if metric_l2 else -float(flat.FLT_MAX)
    best_s = torch.full((nq, k), -float(flat.FLT_MAX))      # "larger is better" domain
    best_i = torch.full((nq, k), -1, dtype=torch.int64)
    qn = (qt * qt).sum(1, keepdim=True)
    for j0 in range(0, n, block):
        xb = xt[j0:j0 + block]
        s = qt @ xb.T
        if metric_l2:
            s = -torch.clamp(qn + (xb * xb).sum(1)[None, :] - 2.0 * s, min=0.0)
        kk = min(k, s.shape[1])
        v, i = torch.topk(s, kk, dim=1)
        cs = torch.cat([best_s, v], 1)
        ci = torch.cat([best_i, i + j0], 1)
        v2, sel = torch.topk(cs, k, dim=1)
        best_s, best_i = v2, torch.gather(ci, 1, sel)
    D = best_s.numpy().copy()
    I = best_i.numpy().copy()
    if metric_l2:
        D = -D
    D[I < 0] = pad
    return D.astype(np.float32), I


class FastBM25:
    """terms x docs matrix of folded impacts (scipy CSR, fp64) built from an oracle BM25Corpus."""

    def __init__(self, corpus):
        import scipy.sparse as sp
        self.N, self.V = corpus.N, corpus.V
        self.idf = np.asarray(corpus.idf, np.float64)
        self.W = sp.csr_matrix((np.asarray(corpus.impact, np.float64), corpus.post_doc.astype(np.int64), corpus.indptr),
                               shape=(self.V, self.N))

    def search(self, queries: Sequence[Sequence[int]], k: int) -> Tuple[np.ndarray, np.ndarray]:
        import scipy.sparse as sp
        import torch
        rows, cols, vals = [], [], []
        for qi, q in enumerate(queries):
            for t in q:
                t = int(t)
                if 0 <= t < self.V:
                    rows.append(qi)
                    cols.append(t)
                    vals.append(self.idf[t])          # duplicates sum: one idf per occurrence
        nq = len(queries)
        Q = sp.csr_matrix((np.asarray(vals, np.float64), (np.asarray(rows, np.int64), np.asarray(cols, np.int64))),
                          shape=(nq, self.V))
        A = (Q @ self.W).astype(np.float32)           # fp64 sums rounded once, like the oracle
        S = np.zeros((nq, k), np.float32)
        I = np.full((nq, k), -1, np.int64)
        dense = torch.from_numpy(A.toarray())
        kk = min(k, self.N)
        v, i = torch.topk(dense, kk, dim=1)
        v, i = v.numpy(), i.numpy()
        ok = v > 0
        S[:, :kk] = np.where(ok, v, 0.0)
        I[:, :kk] = np.where(ok, i, -1)
        return S, I


def retrieve(x: np.ndarray, metric_l2: bool, fast_bm25: FastBM25, q: np.ndarray, query_tokens, top_k: int,
             k_c: int = 50, timings: dict | None = None):
    """Hybrid retrieve on the CPU fast path; fills `timings` with the per-stage seconds."""
    t0 = time.perf_counter()
    D, I = dense_topk(x, q, k_c, metric_l2)
    t1 = time.perf_counter()
    S, J = fast_bm25.search(query_tokens, k_c)
    t2 = time.perf_counter()
    fs, fi = fusion.fuse(fusion.dense_similarity(D, metric_l2), I, S, J, top_k)
    t3 = time.perf_counter()
    if timings is not None:
        timings["dense_s"] = timings.get("dense_s", 0.0) + (t1 - t0)
        timings["bm25_s"] = timings.get("bm25_s", 0.0) + (t2 - t1)
        timings["fusion_s"] = timings.get("fusion_s", 0.0) + (t3 - t2)
    return fs, fi
