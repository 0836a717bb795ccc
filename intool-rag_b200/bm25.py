"""BM25 inverted index on the GPU (host-side wrapper over hr_bm25_* of include/hr_b200.h).

The reference advertises BM25 hybrid search (/root/reference/README.md:54-58, dead constants at
rag/config.py:43-45) but ships no scorer; the definition is SURVEY.md Appendix B:
k1=1.5, b=0.75, Lucene idf by default, duplicate query terms count per occurrence, tokenisation
``text.lower().split()`` (/root/reference/rag/agent/query_processor.py:26).
"""
from __future__ import annotations

import ctypes as C
from typing import List, Sequence, Tuple

import numpy as np

from . import _lib
from .config import config, default_device

_IDF = {"lucene": _lib.IDF_LUCENE, "okapi": _lib.IDF_OKAPI}


def _is_torch_cuda(x) -> bool:
    return type(x).__module__.startswith("torch") and getattr(x, "is_cuda", False)


def tokenize(text: str) -> List[str]:
    """The reference's only tokenisation idiom (rag/agent/query_processor.py:26)."""
    return text.lower().split()


class Vocabulary:
    """word -> id in first-seen order (service adapter; benchmarks use integer ids directly)."""

    def __init__(self):
        self.word_to_id = {}

    def encode(self, text: str, grow: bool = False) -> List[int]:
        out = []
        for w in tokenize(text):
            i = self.word_to_id.get(w)
            if i is None:
                if not grow:
                    continue  # out-of-vocabulary query words match nothing
                i = len(self.word_to_id)
                self.word_to_id[w] = i
            out.append(i)
        return out

    def __len__(self):
        return len(self.word_to_id)


MAX_DISTINCT_TERMS = 64   # per query, what one warp of the scoring kernel owns (include/hr_b200.h: hr_bm25_search)


def cap_query_terms(tokens: Sequence[int], max_distinct: int = MAX_DISTINCT_TERMS) -> List[int]:
    """Service adapter for long query texts: keep every occurrence of the first `max_distinct` distinct term ids
    (duplicates still count per occurrence), drop the rest.  The library itself rejects a query with more than 64
    distinct scorable terms (HR_ERR_INVALID) on every input path instead of truncating silently."""
    seen, out = set(), []
    for t in tokens:
        if t not in seen:
            if len(seen) >= max_distinct:
                continue
            seen.add(t)
        out.append(t)
    return out


def query_csr(queries) -> Tuple[np.ndarray, np.ndarray]:
    """ragged list of int lists -> (indptr int32[nq+1], terms int32[total]); a CSR pair passes through."""
    if isinstance(queries, tuple) and len(queries) == 2:
        ip, tm = queries
        if _is_torch_cuda(ip):
            return ip, tm
        return np.ascontiguousarray(ip, np.int32), np.ascontiguousarray(tm, np.int32)
    nq = len(queries)
    indptr = np.zeros(nq + 1, dtype=np.int32)
    for i, q in enumerate(queries):
        indptr[i + 1] = indptr[i] + len(q)
    terms = np.zeros(int(indptr[-1]), dtype=np.int32)
    for i, q in enumerate(queries):
        if len(q):
            terms[indptr[i]:indptr[i + 1]] = np.asarray(q, dtype=np.int32)
    return indptr, terms


def build_csr(term_ids: np.ndarray, doc_ids: np.ndarray, n_docs: int, vocab: int):
    """Host-side CSR-by-term build from flat (term, doc) token occurrences (ingest path, small corpora).
    Returns indptr int64[V+1], post_doc int32[nnz] ascending per term, post_tf int32[nnz]."""
    t = np.asarray(term_ids, dtype=np.int64)
    d = np.asarray(doc_ids, dtype=np.int64)
    if t.size and (t.min() < 0 or t.max() >= vocab):
        raise ValueError("token id outside [0, vocab)")
    key = t * max(int(n_docs), 1) + d
    uk, tf = np.unique(key, return_counts=True)
    pt = uk // max(int(n_docs), 1)
    indptr = np.zeros(vocab + 1, dtype=np.int64)
    np.add.at(indptr, pt + 1, 1)
    indptr = np.cumsum(indptr)
    return indptr, (uk % max(int(n_docs), 1)).astype(np.int32), tf.astype(np.int32)


class BM25Index:
    def __init__(self, handle: int, device: int):
        self._h = C.c_void_p(handle)
        self.device = device

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().hr_bm25_destroy(self._h)
                self._h = C.c_void_p(None)
        except Exception:
            pass

    # -- construction ------------------------------------------------------------------------------
    @classmethod
    def from_csr(cls, indptr, post_doc, post_tf, doc_len, vocab: int, k1: float | None = None,
                 b: float | None = None, idf: str | None = None, n_docs_global: int = 0,
                 avgdl_global: float = 0.0, df_global=None, device: int | None = None) -> "BM25Index":
        """CSR arrays as numpy (host) or torch CUDA tensors (device, zero-copy into the build)."""
        _lib.require_gpu()
        dev = default_device() if device is None else int(device)
        k1 = config.BM25_K1 if k1 is None else k1
        b = config.BM25_B if b is None else b
        idf = idf or config.BM25_IDF
        on_dev = _is_torch_cuda(indptr)
        keep = []
        if on_dev:
            import torch

            def ptr(a, dt):
                a = a.to(dt).contiguous()
                keep.append(a)
                return a.data_ptr()
            ip, pd = ptr(indptr, torch.int64), ptr(post_doc, torch.int32)
            pt, dl = ptr(post_tf, torch.int32), ptr(doc_len, torch.int32)
            dfg = ptr(df_global, torch.int64) if df_global is not None else None
            n_docs = int(doc_len.shape[0])
            st = _lib.current_stream_ptr(dev)
        else:
            def ptr(a, dt):
                a = np.ascontiguousarray(a, dtype=dt)
                keep.append(a)
                return a.ctypes.data
            ip, pd = ptr(indptr, np.int64), ptr(post_doc, np.int32)
            pt, dl = ptr(post_tf, np.int32), ptr(doc_len, np.int32)
            dfg = ptr(df_global, np.int64) if df_global is not None else None
            n_docs = int(len(doc_len))
            st = None
        if len(indptr) != vocab + 1:
            raise ValueError("indptr must have vocab + 1 entries")
        h = C.c_void_p()
        _lib.check(_lib.lib().hr_bm25_create(ip, pd, pt, dl, n_docs, int(vocab), float(k1), float(b), _IDF[idf],
                                             int(n_docs_global), float(avgdl_global), dfg, int(on_dev), dev, st,
                                             C.byref(h)))
        return cls(h.value, dev)

    @classmethod
    def from_tokens(cls, term_ids, doc_ids, n_docs: int, vocab: int, k1: float | None = None,
                    b: float | None = None, idf: str | None = None, n_docs_global: int = 0,
                    avgdl_global: float = 0.0, df_global=None, device: int | None = None) -> "BM25Index":
        """Flat token occurrences (term_ids[i] occurs in doc_ids[i]), numpy or torch CUDA int32: the CSR is
        built on the GPU (sort + run-length encode), nothing is materialised on the host."""
        _lib.require_gpu()
        dev = default_device() if device is None else int(device)
        k1 = config.BM25_K1 if k1 is None else k1
        b = config.BM25_B if b is None else b
        idf = idf or config.BM25_IDF
        on_dev = _is_torch_cuda(term_ids)
        keep = []
        if on_dev:
            import torch
            t = term_ids.to(torch.int32).contiguous()
            d = doc_ids.to(torch.int32).contiguous()
            keep += [t, d]
            tp, dp, n = t.data_ptr(), d.data_ptr(), int(t.numel())
            dfg = None
            if df_global is not None:
                g = df_global.to(torch.int64).contiguous()
                keep.append(g)
                dfg = g.data_ptr()
            st = _lib.current_stream_ptr(dev)
        else:
            t = np.ascontiguousarray(term_ids, dtype=np.int32)
            d = np.ascontiguousarray(doc_ids, dtype=np.int32)
            keep += [t, d]
            tp, dp, n = t.ctypes.data, d.ctypes.data, int(t.size)
            dfg = None
            if df_global is not None:
                g = np.ascontiguousarray(df_global, dtype=np.int64)
                keep.append(g)
                dfg = g.ctypes.data
            st = None
        if int(d.shape[0]) != n:
            raise ValueError("term_ids and doc_ids must have the same length")
        h = C.c_void_p()
        _lib.check(_lib.lib().hr_bm25_create_from_tokens(tp, dp, n, int(n_docs), int(vocab), float(k1), float(b),
                                                         _IDF[idf], int(n_docs_global), float(avgdl_global), dfg,
                                                         int(on_dev), dev, st, C.byref(h)))
        return cls(h.value, dev)

    @classmethod
    def from_docs(cls, docs: Sequence[Sequence[int]], vocab: int, **kw) -> "BM25Index":
        """docs: list of token-id lists (one per chunk, row i of the dense index = doc i)."""
        doc_len = np.array([len(d) for d in docs], dtype=np.int64)
        if len(docs) and doc_len.sum():
            t = np.concatenate([np.asarray(d, dtype=np.int32) for d in docs if len(d)])
            dd = np.repeat(np.arange(len(docs), dtype=np.int32), doc_len)
        else:
            t = np.zeros(0, np.int32)
            dd = np.zeros(0, np.int32)
        if t.size and (t.min() < 0 or t.max() >= vocab):
            raise ValueError("token id outside [0, vocab)")
        return cls.from_tokens(t, dd, len(docs), vocab, **kw)

    # -- persistence -------------------------------------------------------------------------------
    def save(self, path: str) -> None:
        """Write the index (postings, folded impacts, idf) next to the faiss file: "HRBM25" v1."""
        import os
        _lib.check(_lib.lib().hr_bm25_save(self._h, os.fsencode(str(path))))

    @classmethod
    def load(cls, path: str, device: int | None = None) -> "BM25Index":
        import os
        _lib.require_gpu()
        dev = default_device() if device is None else int(device)
        h = C.c_void_p()
        _lib.check(_lib.lib().hr_bm25_load(os.fsencode(str(path)), dev, C.byref(h)))
        return cls(h.value, dev)

    # -- attributes ------------------------------------------------------------------------------
    @property
    def ndocs(self) -> int:
        return int(_lib.lib().hr_bm25_ndocs(self._h))

    @property
    def vocab(self) -> int:
        return int(_lib.lib().hr_bm25_vocab(self._h))

    @property
    def nnz(self) -> int:
        return int(_lib.lib().hr_bm25_nnz(self._h))

    def set_id_base(self, base: int) -> None:
        _lib.check(_lib.lib().hr_bm25_set_id_base(self._h, int(base)))

    # -- search ----------------------------------------------------------------------------------
    def search(self, queries, k: int, return_postings: bool = False):
        """queries: ragged list of token-id lists, or CSR (indptr, terms) as numpy / torch CUDA.
        Returns (S float32[nq,k] descending, I int64[nq,k]); padding I=-1, S=0."""
        indptr, terms = query_csr(queries)
        k = int(k)
        touched = C.c_int64(0)
        if _is_torch_cuda(indptr):
            import torch
            indptr = indptr.to(torch.int32).contiguous()
            terms = terms.to(torch.int32).contiguous()
            nq = int(indptr.shape[0]) - 1
            S = torch.empty((nq, k), dtype=torch.float32, device=indptr.device)
            I = torch.empty((nq, k), dtype=torch.int64, device=indptr.device)
            _lib.check(_lib.lib().hr_bm25_search(self._h, indptr.data_ptr(), terms.data_ptr(), nq, int(terms.numel()), k,
                                                 S.data_ptr(),
                                                 I.data_ptr(), 1, _lib.current_stream_ptr(self.device),
                                                 C.byref(touched)))
        else:
            nq = len(indptr) - 1
            S = np.empty((nq, k), dtype=np.float32)
            I = np.empty((nq, k), dtype=np.int64)
            _lib.check(_lib.lib().hr_bm25_search(self._h, indptr.ctypes.data, terms.ctypes.data, nq, int(terms.size), k,
                                                 S.ctypes.data, I.ctypes.data, 0,
                                                 _lib.current_stream_ptr(self.device), C.byref(touched)))
        if return_postings:
            return S, I, int(touched.value)
        return S, I
