"""ORACLE — test infrastructure only.  End-to-end CPU restatement of the hybrid query path
``retrieve(query_embeddings, query_tokens, top_k)`` (surface advertised at
/root/reference/README.md:90, absent from the tree — SURVEY.md §0 F2): dense flat search
(/root/reference/rag/storage/faiss_index.py:81-89) + BM25 + fusion, SURVEY.md Appendix B.
"""
from __future__ import annotations

import numpy as np

from . import flat, fusion


def candidate_depth(top_k: int) -> int:
    """k_c = max(top_k, 50); 50 is the live path's top_chunks
    (/root/reference/rag/query/page_retriever.py:81)."""
    return max(int(top_k), 50)


def retrieve(index: "flat.IndexFlat", corpus, query_embeddings, query_tokens, top_k: int,
             mode: str = "weighted", w_vec: float = fusion.VECTOR_WEIGHT,
             w_bm25: float = fusion.BM25_WEIGHT, k_c: int | None = None,
             precision: str = "f32"):
    """Returns (scores float32[nq,top_k], ids int64[nq,top_k], parts dict)."""
    kc = candidate_depth(top_k) if k_c is None else int(k_c)
    D, I = index.search(np.asarray(query_embeddings, np.float32), kc, precision=precision)
    sim = fusion.dense_similarity(D, index.metric_type == flat.METRIC_L2)
    if corpus is None or query_tokens is None:
        S = np.zeros((len(I), kc), np.float32)
        J = np.full((len(I), kc), -1, np.int64)
    else:
        S, J = corpus.search(query_tokens, kc)
    fs, fi = fusion.fuse(sim, I, S, J, top_k, mode=mode, w_vec=w_vec, w_bm25=w_bm25)
    return fs, fi, {"dense_D": D, "dense_I": I, "bm25_S": S, "bm25_I": J}
