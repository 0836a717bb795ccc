"""ORACLE — test infrastructure only.  CPU definition of BM25 over integer token ids.

PARITY STATUS: **unpinned** — the reference has *no* BM25 code at all (README claim at
/root/reference/README.md:54-58, dead constants /root/reference/rag/config.py:43-45;
SURVEY.md §0 F1).  This file therefore *defines* the scorer the CUDA path must match
 and is pinned by hand-computed known answers in tests/golden/.

Definition
  tokens     : integer ids in [0, V); the service adapter tokenises with the reference's
               only idiom ``text.lower().split()`` (/root/reference/rag/agent/query_processor.py:26)
  tf(t,d)    : multiplicity of t in doc d;   df(t): number of docs containing t
  dl(d)      : number of tokens in d;        avgdl = mean(dl) in fp64
  idf(t)     : 'lucene'  ln((N - df + 0.5)/(df + 0.5) + 1)          (default, > 0)
               'okapi'   ln((N - df + 0.5)/(df + 0.5)), negatives floored to
                         0.25 * mean(idf) (rank_bm25.BM25Okapi behaviour)
  impact(t,d): tf (k1+1) / (tf + k1 (1 - b + b dl/avgdl)),  k1 = 1.5, b = 0.75
  score(q,d) : sum over query-term *occurrences* (duplicates count) of idf(t) impact(t,d);
               ids outside [0,V) are ignored.
  Accumulate in fp64, round to fp32 once.  Only docs with score > 0 are candidates.
  Order: (score desc, id asc).  Fewer than k candidates -> id -1, score 0.
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np

K1 = 1.5
B = 0.75


class BM25Corpus:
    """CSR-by-term inverted index built from docs given as lists of int token ids."""

    def __init__(self, docs: Sequence[Sequence[int]], vocab_size: int, k1: float = K1,
                 b: float = B, idf_variant: str = "lucene"):
        self.N = len(docs)
        self.V = int(vocab_size)
        self.k1, self.b, self.idf_variant = float(k1), float(b), idf_variant
        doc_len = np.array([len(d) for d in docs], dtype=np.int64)
        terms, dids = [], []
        for i, d in enumerate(docs):
            if len(d):
                t = np.asarray(d, dtype=np.int64)
                assert t.min() >= 0 and t.max() < self.V
                terms.append(t)
                dids.append(np.full(len(t), i, dtype=np.int64))
        if terms:
            t = np.concatenate(terms)
            dd = np.concatenate(dids)
        else:
            t = np.zeros(0, np.int64)
            dd = np.zeros(0, np.int64)
        self._from_pairs(t, dd, doc_len)

    @classmethod
    def from_token_matrix(cls, term_ids: np.ndarray, doc_ids: np.ndarray, doc_len: np.ndarray,
                          vocab_size: int, k1: float = K1, b: float = B,
                          idf_variant: str = "lucene") -> "BM25Corpus":
        """Build from flat (term, doc) occurrence pairs (what the synthetic generator emits)."""
        self = cls.__new__(cls)
        self.N = int(len(doc_len))
        self.V = int(vocab_size)
        self.k1, self.b, self.idf_variant = float(k1), float(b), idf_variant
        self._from_pairs(np.asarray(term_ids, np.int64), np.asarray(doc_ids, np.int64),
                         np.asarray(doc_len, np.int64))
        return self

    @classmethod
    def from_csr(cls, indptr, post_doc, post_tf, doc_len, vocab_size: int, k1: float = K1, b: float = B,
                 idf_variant: str = "lucene", n_docs_global: int = 0, avgdl_global: float = 0.0,
                 df_global=None) -> "BM25Corpus":
        """Wrap ready CSR arrays (optionally a row shard scored with corpus-wide statistics)."""
        self = cls.__new__(cls)
        self.N = int(len(doc_len))
        self.V = int(vocab_size)
        self.k1, self.b, self.idf_variant = float(k1), float(b), idf_variant
        self.indptr = np.asarray(indptr, np.int64)
        self.post_doc = np.asarray(post_doc, np.int32)
        self.post_tf = np.asarray(post_tf, np.int32)
        self.doc_len = np.asarray(doc_len, np.int32)
        self.avgdl = float(avgdl_global) if avgdl_global > 0 else (
            float(self.doc_len.astype(np.float64).mean()) if self.N else 0.0)
        self.df = np.diff(self.indptr).astype(np.int64) if df_global is None else np.asarray(df_global, np.int64)
        self.idf = idf_table(self.df, int(n_docs_global) if n_docs_global > 0 else self.N, idf_variant)
        dl = self.doc_len[self.post_doc].astype(np.float64)
        tf64 = self.post_tf.astype(np.float64)
        norm = self.k1 * (1.0 - self.b + self.b * dl / self.avgdl) if self.avgdl > 0 else self.k1
        self.impact = tf64 * (self.k1 + 1.0) / (tf64 + norm)
        return self

    def _from_pairs(self, t: np.ndarray, dd: np.ndarray, doc_len: np.ndarray) -> None:
        N, V = self.N, self.V
        self.doc_len = doc_len.astype(np.int32)
        self.avgdl = float(doc_len.astype(np.float64).mean()) if N else 0.0
        key = t * max(N, 1) + dd
        uk, tf = np.unique(key, return_counts=True)  # sorted by (term, doc)
        pt = uk // max(N, 1)
        pd = uk % max(N, 1)
        self.indptr = np.zeros(V + 1, dtype=np.int64)
        np.add.at(self.indptr, pt + 1, 1)
        self.indptr = np.cumsum(self.indptr)
        self.post_doc = pd.astype(np.int32)
        self.post_tf = tf.astype(np.int32)
        self.df = np.diff(self.indptr).astype(np.int64)
        self.idf = idf_table(self.df, N, self.idf_variant)  # fp64
        dl = self.doc_len[self.post_doc].astype(np.float64)
        tf64 = self.post_tf.astype(np.float64)
        norm = self.k1 * (1.0 - self.b + self.b * dl / self.avgdl) if self.avgdl > 0 else self.k1
        self.impact = tf64 * (self.k1 + 1.0) / (tf64 + norm)  # fp64 [nnz]

    # -- scoring ---------------------------------------------------------------
    def scores(self, query: Sequence[int]) -> np.ndarray:
        """fp64 score of every doc for one query (list of term ids, duplicates count)."""
        acc = np.zeros(self.N, dtype=np.float64)
        for t in query:
            t = int(t)
            if t < 0 or t >= self.V:
                continue
            a, b_ = self.indptr[t], self.indptr[t + 1]
            acc[self.post_doc[a:b_]] += self.idf[t] * self.impact[a:b_]
        return acc

    def search(self, queries: Sequence[Sequence[int]], k: int) -> Tuple[np.ndarray, np.ndarray]:
        nq = len(queries)
        S = np.zeros((nq, k), dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        for qi, q in enumerate(queries):
            acc = self.scores(q).astype(np.float32)
            cand = np.nonzero(acc > 0)[0]
            if cand.size == 0:
                continue
            kk = min(k, cand.size)
            if cand.size > 4 * kk:
                kth = -np.partition(-acc[cand], kk - 1)[kk - 1]
                cand = cand[acc[cand] >= kth]
            order = np.lexsort((cand, -acc[cand]))[:kk]
            sel = cand[order]
            I[qi, :kk] = sel
            S[qi, :kk] = acc[sel]
        return S, I


def idf_table(df: np.ndarray, N: int, variant: str = "lucene") -> np.ndarray:
    df = df.astype(np.float64)
    if variant == "lucene":
        return np.log((N - df + 0.5) / (df + 0.5) + 1.0)
    if variant == "okapi":
        raw = np.log((N - df + 0.5) / (df + 0.5))
        present = df > 0
        mean_idf = raw[present].mean() if present.any() else 0.0
        eps = 0.25 * mean_idf
        return np.where(raw < 0, eps, raw)
    raise ValueError(f"unknown idf variant {variant!r}")


def to_query_csr(queries: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    indptr = np.zeros(len(queries) + 1, dtype=np.int32)
    for i, q in enumerate(queries):
        indptr[i + 1] = indptr[i] + len(q)
    terms = np.zeros(int(indptr[-1]), dtype=np.int32)
    for i, q in enumerate(queries):
        terms[indptr[i]:indptr[i + 1]] = np.asarray(q, dtype=np.int32)
    return indptr, terms


def tokenize(text: str) -> List[str]:
    """The reference's only tokenisation idiom (/root/reference/rag/agent/query_processor.py:26)."""
    return text.lower().split()
