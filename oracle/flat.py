"""ORACLE — test infrastructure only.  CPU restatement of the flat (exhaustive) index.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU arm may import this
module; the product package (``intool-rag_b200/``) never does.

PARITY STATUS: **unpinned by the reference's own tests** (the reference ships none,
SURVEY.md §0 F5).  The arithmetic lives in the un-vendored third-party wheel
``faiss-cpu==1.7.4`` (/root/reference/rag/requirements.txt:25) which is not installed
here, so this file restates its *published* ``IndexFlat`` behaviour (SURVEY.md
Appendix A) and is anchored on the reference's own call sites:

* build:   /root/reference/rag/storage/faiss_index.py:121-124  (float32, IndexFlatL2, add)
* search:  /root/reference/rag/storage/faiss_index.py:81-89    (float32 (1,d) query,
           squared-L2 ascending, score = clamp(1 - dist/2, 0, 1))
* persist: /root/reference/rag/storage/faiss_index.py:54,133   (read_index / write_index)
* agent:   /root/reference/rag/agent/search_engine.py:45-51    (score = 1/(1+dist), drop idx<0)

What pins it instead: the known-answer vectors in tests/golden/ (hand-computed
closed-form cases) and the outputs of the reference's *unmodified* wrapper run on top of
this stand-in (tests/golden/make_golden.py).

Semantics restated (faiss 1.7.4 IndexFlat):
  * ``add`` appends raw fp32 rows, ids are positions 0..ntotal-1.
  * METRIC_L2 returns squared Euclidean distance ascending; METRIC_INNER_PRODUCT returns
    inner products descending.  Exact, exhaustive.
  * fewer than k rows -> label -1, distance +FLT_MAX (L2) / -FLT_MAX (IP).
  * documented tie rule (this build's total order): (better score first, then id asc).
"""
from __future__ import annotations

import struct
from typing import Tuple

import numpy as np

METRIC_INNER_PRODUCT = 0
METRIC_L2 = 1
FLT_MAX = np.float32(3.4028234663852886e38)


def topk_rows(scores: np.ndarray, k: int, largest: bool) -> Tuple[np.ndarray, np.ndarray]:
    """Exact top-k per row under the documented total order (score best-first, id asc).

    scores: [nq, n] float array.  Returns (vals [nq,k] same dtype, ids [nq,k] int64),
    padded with id -1 / +-FLT_MAX when n < k.
    """
    nq, n = scores.shape
    out_v = np.full((nq, k), -FLT_MAX if largest else FLT_MAX, dtype=scores.dtype)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    if n == 0 or k == 0:
        return out_v, out_i
    key = -scores if largest else scores  # ascending key == best first
    for r in range(nq):
        row = key[r]
        kk = min(k, n)
        nan = np.isnan(row)
        if nan.any():
            # a NaN score is never a candidate: faiss' heap test `C::cmp(heap_top, score)` is false for NaN, so a
            # query (or corpus row) with a NaN component yields padding (by recollection)
            cand = np.nonzero(~nan)[0]
            kk = min(k, cand.size)
            order = np.lexsort((cand, row[cand]))[:kk]
            sel = cand[order]
            out_i[r, :kk] = sel
            out_v[r, :kk] = scores[r, sel]
            continue
        if n > 4 * kk:
            # candidates: everything <= kth key (keeps all boundary ties), then exact sort
            kth = np.partition(row, kk - 1)[kk - 1]
            cand = np.nonzero(row <= kth)[0]
        else:
            cand = np.arange(n)
        order = np.lexsort((cand, row[cand]))[:kk]
        sel = cand[order]
        out_i[r, :kk] = sel
        out_v[r, :kk] = scores[r, sel]
    return out_v, out_i


class IndexFlat:
    """numpy restatement of faiss.IndexFlat{L2,IP} (see module docstring)."""

    def __init__(self, d: int, metric: int = METRIC_L2):
        self.d = int(d)
        self.metric_type = int(metric)
        self.is_trained = True
        self._x = np.zeros((0, self.d), dtype=np.float32)

    @property
    def ntotal(self) -> int:
        return int(self._x.shape[0])

    def add(self, x) -> None:
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, "add: x must be [n, d]"
        self._x = np.concatenate([self._x, x], axis=0)

    def reset(self) -> None:
        self._x = np.zeros((0, self.d), dtype=np.float32)

    def reconstruct(self, i: int) -> np.ndarray:
        return self._x[int(i)].copy()

    # -- score matrices ---------------------------------------------------
    def scores_f64(self, q: np.ndarray) -> np.ndarray:
        """fp64 "truth" of the metric for fp32 inputs: [nq, ntotal]."""
        q64 = np.asarray(q, dtype=np.float64)
        x64 = self._x.astype(np.float64)
        ip = q64 @ x64.T
        if self.metric_type == METRIC_INNER_PRODUCT:
            return ip
        qn = (q64 * q64).sum(1)[:, None]
        xn = (x64 * x64).sum(1)[None, :]
        return np.maximum(qn + xn - 2.0 * ip, 0.0)

    def scores_f32(self, q: np.ndarray, block: int = 1024) -> np.ndarray:
        """fp32 scores the way faiss's BLAS path forms them (nq >= 20): sgemm in
        database blocks of 1024 rows; L2 as |x|^2+|y|^2-2<x,y> clamped at 0."""
        q = np.ascontiguousarray(q, dtype=np.float32)
        n = self.ntotal
        out = np.empty((q.shape[0], n), dtype=np.float32)
        qn = (q * q).sum(1, dtype=np.float32)[:, None]
        for j0 in range(0, n, block):
            xb = self._x[j0:j0 + block]
            ip = q @ xb.T
            if self.metric_type == METRIC_INNER_PRODUCT:
                out[:, j0:j0 + block] = ip
            else:
                xn = (xb * xb).sum(1, dtype=np.float32)[None, :]
                out[:, j0:j0 + block] = np.maximum(qn + xn - np.float32(2.0) * ip, np.float32(0))
        return out

    # -- search -----------------------------------------------------------
    def search(self, x, k: int, precision: str = "f32"):
        """(D float32[nq,k], I int64[nq,k]).  precision 'f32' mirrors faiss; 'f64' is the
        truth variant used to *measure* tolerance budgets (scores still returned as fp32
        roundings of the fp64 value)."""
        x = np.ascontiguousarray(x, dtype=np.float32)
        assert x.ndim == 2 and x.shape[1] == self.d, "search: x must be [nq, d]"
        largest = self.metric_type == METRIC_INNER_PRODUCT
        nq = x.shape[0]
        D = np.full((nq, k), -FLT_MAX if largest else FLT_MAX, dtype=np.float32)
        I = np.full((nq, k), -1, dtype=np.int64)
        if self.ntotal == 0 or nq == 0:
            return D, I
        step = max(1, (1 << 27) // max(self.ntotal, 1))  # bound the score matrix
        for q0 in range(0, nq, step):
            xs = x[q0:q0 + step]
            s = self.scores_f64(xs) if precision == "f64" else self.scores_f32(xs)
            v, i = topk_rows(s, k, largest)
            D[q0:q0 + step] = v.astype(np.float32)
            I[q0:q0 + step] = i
        return D, I


def IndexFlatL2(d: int) -> IndexFlat:
    return IndexFlat(d, METRIC_L2)


def IndexFlatIP(d: int) -> IndexFlat:
    return IndexFlat(d, METRIC_INNER_PRODUCT)


# -- reference-side score transforms ------------------------------------------
def reference_score_from_l2(dist):
    """/root/reference/rag/storage/faiss_index.py:86-88: clamp(1 - dist/2, 0, 1),
    evaluated in Python float (fp64) on the fp32 distance."""
    s = 1.0 - (np.asarray(dist, dtype=np.float64) / 2.0)
    return np.clip(s, 0.0, 1.0)


def agent_score_from_l2(dist):
    """/root/reference/rag/agent/search_engine.py:50: 1 / (1 + dist)."""
    return 1.0 / (1.0 + np.asarray(dist, dtype=np.float64))


# -- faiss flat index file format ------------------
def write_index(index: IndexFlat, path: str) -> None:
    fourcc = b"IxFI" if index.metric_type == METRIC_INNER_PRODUCT else b"IxF2"
    with open(path, "wb") as f:
        f.write(fourcc)
        f.write(struct.pack("<i", index.d))
        f.write(struct.pack("<q", index.ntotal))
        f.write(struct.pack("<q", 1 << 20))
        f.write(struct.pack("<q", 1 << 20))
        f.write(struct.pack("<B", 1))
        f.write(struct.pack("<i", index.metric_type))
        f.write(struct.pack("<Q", index.ntotal * index.d))
        f.write(np.ascontiguousarray(index._x, dtype="<f4").tobytes())


def read_index(path: str) -> IndexFlat:
    with open(path, "rb") as f:
        fourcc = f.read(4)
        if fourcc not in (b"IxFI", b"IxF2", b"IxFl"):
            raise RuntimeError(f"unsupported index fourcc {fourcc!r}")
        (d,) = struct.unpack("<i", f.read(4))
        (ntotal,) = struct.unpack("<q", f.read(8))
        f.read(16)
        (is_trained,) = struct.unpack("<B", f.read(1))
        (metric,) = struct.unpack("<i", f.read(4))
        if metric > 1:
            f.read(4)
        (count,) = struct.unpack("<Q", f.read(8))
        if count != ntotal * d:
            raise RuntimeError("corrupt flat index: vector length mismatch")
        x = np.frombuffer(f.read(count * 4), dtype="<f4").reshape(ntotal, d)
    idx = IndexFlat(d, metric)
    idx.add(x)
    return idx
