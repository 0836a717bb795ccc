"""One logical corpus over every per-document file under STORAGE_DIR.

The reference writes one ``{doc_id}_faiss.index`` + ``{doc_id}_chunks.json`` per ingested document
(/root/reference/rag/ingest/ingestion_pipeline.py:88-94) and then searches only the FIRST index file
(/root/reference/rag/storage/faiss_index.py:162-167), re-parsing that document's whole chunk JSON on every
query (:175, /root/reference/rag/storage/file_storage.py:139-166).  Here:

* ``Corpus.load`` concatenates all flat index files (sorted by name) into ONE index in HBM.  Global row
  ``g`` = ``row0[doc] + local_row``; the (doc, local row) <-> global row table is two small arrays.  A rank
  of a row-sharded deployment loads only the rows [lo, hi) it owns, straight from the files' row region
  (``np.memmap`` at byte 45): no document is read twice.
* BM25: ingest writes, next to the single-document sidecar, the raw CSR of the document
  (``{doc_id}_bm25_csr.npz``: indptr, post_doc, post_tf, doc_len, local vocabulary).  The corpus index is
  their concatenation under a unified vocabulary (first-seen order over the files), built WITHOUT a sort:
  every document's lists are already (term, doc)-sorted and documents own disjoint ascending row ranges, so
  a global term's list is the concatenation of the per-document lists.  idf / avgdl come from the whole
  corpus (every file's df and doc_len), also on a rank that holds only a shard.
* ``ChunkStore``: per chunk a (file, byte offset, byte length) triple into the reference's own
  ``{doc_id}_chunks.json`` files plus the chunk's page number; a query reads and parses only its k chunks.
  Host memory: 20 bytes per chunk (200 MB for 10M chunks) instead of every chunk as a Python dict (>= 10 GB at
  10M chunks).
"""
from __future__ import annotations

import glob
import json
import mmap
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import faiss
from .sharded import shard_bounds

_FAISS_DATA_OFFSET = 45      # fourcc 4 | d 4 | ntotal 8 | 2 x 8 | trained 1 | metric 4 | count 8  (metric <= 1)


def read_flat_header(path: str) -> Tuple[int, int, int]:
    """(d, ntotal, metric) of a faiss flat index file ("IxF2" / "IxFI"), without reading the rows."""
    with open(path, "rb") as f:
        h = f.read(_FAISS_DATA_OFFSET)
    if len(h) < _FAISS_DATA_OFFSET or h[:4] not in (b"IxF2", b"IxFI", b"IxFl"):
        raise RuntimeError(f"not a faiss flat index (IxF2/IxFI): {path}")
    d = int.from_bytes(h[4:8], "little", signed=True)
    n = int.from_bytes(h[8:16], "little", signed=True)
    metric = int.from_bytes(h[33:37], "little", signed=True)
    count = int.from_bytes(h[37:45], "little")
    if d <= 0 or n < 0 or metric not in (0, 1) or count != n * d:
        raise RuntimeError(f"corrupt flat index header: {path}")
    if os.path.getsize(path) < _FAISS_DATA_OFFSET + 4 * n * d:
        raise RuntimeError(f"truncated index file: {path}")
    return d, n, metric


def doc_id_of(index_path: str) -> str:
    return os.path.basename(index_path)[: -len(".index")].replace("_faiss", "")


# ---------------------------------------------------------------------------------------------------
# chunk metadata: fetch k rows, not the corpus
# ---------------------------------------------------------------------------------------------------
def _chunk_spans(path: str, with_pages: bool = False):
    """Byte spans [[offset, length], ...] of the elements of the "chunks" array of a ``*_chunks.json`` file, in
    the order ``list({c["chunk_id"]: c for c in chunks}.values())`` would have (the reference's row -> chunk
    mapping, faiss_index.py:176-181): position of the FIRST occurrence of a chunk_id, content of the LAST.
    with_pages: also the chunks' "page" values (int32, 0 when absent: the wrapper's default, faiss_index.py:187)."""
    with open(path, "rb") as f:
        raw = f.read()
    text = raw.decode("utf-8")
    ascii_only = len(text) == len(raw)
    dec = json.JSONDecoder()
    key = text.find('"chunks"')
    if key < 0:
        return (np.zeros((0, 2), np.int64), np.zeros(0, np.int32)) if with_pages else np.zeros((0, 2), np.int64)
    i = text.index("[", key) + 1
    pages: List[int] = []
    spans: List[List[int]] = []
    pos_of: Dict[str, int] = {}
    n = len(text)
    char_pos, byte_pos = 0, 0        # a character index whose byte offset is known (non-ASCII files)
    while True:
        while i < n and text[i] in " \t\r\n,":
            i += 1
        if i >= n or text[i] == "]":
            break
        obj, end = dec.raw_decode(text, i)
        if ascii_only:
            off, ln = i, end - i
        else:
            off = byte_pos + len(text[char_pos:i].encode("utf-8"))
            ln = len(text[i:end].encode("utf-8"))
            char_pos, byte_pos = end, off + ln
        cid = obj.get("chunk_id") if isinstance(obj, dict) else None
        page = obj.get("page", 0) if isinstance(obj, dict) else 0
        page = int(page) if isinstance(page, (int, float)) else 0
        if cid is not None and cid in pos_of:
            spans[pos_of[cid]] = [off, ln]
            pages[pos_of[cid]] = page
        else:
            if cid is not None:
                pos_of[cid] = len(spans)
            spans.append([off, ln])
            pages.append(page)
        i = end
    out = np.asarray(spans, np.int64).reshape(-1, 2)
    return (out, np.asarray(pages, np.int32)) if with_pages else out


class ChunkStore:
    """row -> chunk dict, reading only the requested rows from the per-document JSON files."""

    def __init__(self, chunk_paths: Sequence[str], rows_per_doc: Sequence[int]):
        self.paths = [str(p) for p in chunk_paths]
        files, offs, lens, pages = [], [], [], []
        self.row0 = np.zeros(len(self.paths) + 1, np.int64)
        for fi, (p, n_rows) in enumerate(zip(self.paths, rows_per_doc)):
            spans, pg = _chunk_spans(p, True) if os.path.exists(p) else (np.zeros((0, 2), np.int64), np.zeros(0, np.int32))
            n = int(n_rows)
            pr = np.zeros(n, np.int32)
            pr[:min(n, len(pg))] = pg[:n]
            pages.append(pr)
            o = np.full(n, -1, np.int64)           # rows without a chunk (shorter JSON) resolve to None
            ln = np.zeros(n, np.int32)
            m = min(n, len(spans))
            o[:m], ln[:m] = spans[:m, 0], spans[:m, 1]
            files.append(np.full(n, fi, np.int32))
            offs.append(o)
            lens.append(ln)
            self.row0[fi + 1] = self.row0[fi] + n
        self.file_of = np.concatenate(files) if files else np.zeros(0, np.int32)
        self.offset = np.concatenate(offs) if offs else np.zeros(0, np.int64)
        self.length = np.concatenate(lens) if lens else np.zeros(0, np.int32)
        self.page_of = np.concatenate(pages) if pages else np.zeros(0, np.int32)   # row -> page (device page ranking)
        self._maps: Dict[int, mmap.mmap] = {}

    def __len__(self) -> int:
        return int(self.file_of.shape[0])

    @property
    def table_bytes(self) -> int:
        return int(self.file_of.nbytes + self.offset.nbytes + self.length.nbytes + self.page_of.nbytes)

    def _map(self, fi: int) -> mmap.mmap:
        m = self._maps.get(fi)
        if m is None:
            if len(self._maps) >= 256:               # bounded number of open mappings (thousands of documents)
                old = next(iter(self._maps))
                self._maps.pop(old).close()
            with open(self.paths[fi], "rb") as f:
                m = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
            self._maps[fi] = m
        return m

    def get(self, row: int) -> Optional[dict]:
        row = int(row)
        if row < 0 or row >= len(self) or self.offset[row] < 0:
            return None
        o, ln = int(self.offset[row]), int(self.length[row])
        return json.loads(self._map(int(self.file_of[row]))[o:o + ln].decode("utf-8"))

    def close(self) -> None:
        for m in self._maps.values():
            m.close()
        self._maps.clear()


# ---------------------------------------------------------------------------------------------------
# BM25 over many documents
# ---------------------------------------------------------------------------------------------------
def csr_sidecar_path(storage_dir: str, doc_id: str) -> str:
    return os.path.join(storage_dir, f"{doc_id}_bm25_csr.npz")


def save_doc_csr(path: str, indptr, post_doc, post_tf, doc_len, words: Sequence[str]) -> None:
    np.savez(path, indptr=np.asarray(indptr, np.int64), post_doc=np.asarray(post_doc, np.int32),
             post_tf=np.asarray(post_tf, np.int32), doc_len=np.asarray(doc_len, np.int32),
             words=np.asarray(list(words), dtype=object))


def merge_doc_csrs(parts: Sequence[dict], row0: Sequence[int], lo: int, hi: int):
    """parts[f] = {indptr, post_doc, post_tf, doc_len, words} of document f whose rows are
    [row0[f], row0[f] + len(doc_len)).  Returns the CSR of the rows [lo, hi) under the unified vocabulary, with
    the corpus-wide statistics: (indptr, post_doc (local to lo), post_tf, doc_len[lo:hi], words, df_global,
    n_docs_global, avgdl_global).  No sort: a global term's list is the concatenation of the documents' lists."""
    word_id: Dict[str, int] = {}
    maps = []
    for p in parts:
        m = np.empty(len(p["words"]), np.int64)
        for t, w in enumerate(p["words"]):
            g = word_id.get(w)
            if g is None:
                g = len(word_id)
                word_id[w] = g
            m[t] = g
        maps.append(m)
    V = max(len(word_id), 1)
    df_g = np.zeros(V, np.int64)
    n_g, len_g = 0, 0
    for p, m in zip(parts, maps):
        np.add.at(df_g, m, np.diff(p["indptr"]))
        n_g += len(p["doc_len"])
        len_g += int(np.asarray(p["doc_len"], np.int64).sum())
    # postings of the shard, per part restricted to [lo, hi)
    sel = []
    df_s = np.zeros(V, np.int64)
    for p, m, r0 in zip(parts, maps, row0):
        n = len(p["doc_len"])
        if r0 + n <= lo or r0 >= hi or n == 0:
            sel.append(None)
            continue
        terms = np.repeat(np.arange(len(p["words"]), dtype=np.int64), np.diff(p["indptr"]))
        docs = p["post_doc"].astype(np.int64) + r0
        keep = (docs >= lo) & (docs < hi)
        terms, docs, tf = terms[keep], docs[keep] - lo, p["post_tf"][keep]
        per_term = np.bincount(terms, minlength=len(p["words"]))
        np.add.at(df_s, m, per_term)
        sel.append((terms, docs, tf, per_term))
    indptr = np.zeros(V + 1, np.int64)
    indptr[1:] = np.cumsum(df_s)
    nnz = int(indptr[-1])
    post_doc = np.empty(nnz, np.int32)
    post_tf = np.empty(nnz, np.int32)
    fill = indptr[:-1].copy()
    for s, m in zip(sel, maps):
        if s is None:
            continue
        terms, docs, tf, per_term = s
        first = np.cumsum(per_term) - per_term
        pos = fill[m[terms]] + (np.arange(terms.size, dtype=np.int64) - first[terms])
        post_doc[pos] = docs
        post_tf[pos] = tf
        np.add.at(fill, m, per_term)
    doc_len = np.concatenate([np.asarray(p["doc_len"], np.int32) for p in parts]) if parts else np.zeros(0, np.int32)
    words = [None] * len(word_id)
    for w, g in word_id.items():
        words[g] = w
    return (indptr, post_doc, post_tf, doc_len[lo:hi], words, df_g, n_g, (len_g / n_g if n_g else 0.0))


# ---------------------------------------------------------------------------------------------------
@dataclass
class DocEntry:
    doc_id: str
    index_path: str
    n_rows: int
    row0: int


class Corpus:
    """Every ``*_faiss.index`` under `storage_dir` as one index (+ one BM25 index when every document has its
    CSR sidecar), optionally only the row shard of `rank` / `world`."""

    def __init__(self, storage_dir: str, rank: int = 0, world: int = 1, storage: Optional[str] = None,
                 device: Optional[int] = None, with_bm25: bool = True):
        from .bm25 import BM25Index, Vocabulary
        self.storage_dir = str(storage_dir)
        paths = sorted(glob.glob(os.path.join(self.storage_dir, "*_faiss.index")))
        self.docs: List[DocEntry] = []
        d0 = metric0 = None
        r = 0
        for p in paths:
            d, n, metric = read_flat_header(p)
            if d0 is None:
                d0, metric0 = d, metric
            elif (d, metric) != (d0, metric0):
                raise RuntimeError(f"{p}: d={d} metric={metric} differs from the corpus ({d0}, {metric0})")
            self.docs.append(DocEntry(doc_id_of(p), p, n, r))
            r += n
        self.ntotal_global = r
        self.d = d0 or 0
        self.metric = metric0 if metric0 is not None else faiss.METRIC_L2
        self.rank, self.world = int(rank), int(world)
        self.lo, self.hi = shard_bounds(self.ntotal_global, world, rank)
        self.row0 = np.array([e.row0 for e in self.docs] + [self.ntotal_global], np.int64)
        self.signature = tuple((e.index_path, os.path.getmtime(e.index_path), e.n_rows) for e in self.docs)
        self.index = None
        self.bm25 = None
        self.vocab = None
        if not self.docs:
            return
        cls = faiss.IndexFlatIP if self.metric == faiss.METRIC_INNER_PRODUCT else faiss.IndexFlatL2
        kw = {} if device is None else {"device": device}
        self.index = cls(self.d, storage=storage, **kw)
        self.index.reserve(max(self.hi - self.lo, 1))
        for e in self.docs:                                   # only the rows this rank owns are read
            a, b = max(self.lo, e.row0), min(self.hi, e.row0 + e.n_rows)
            if a >= b:
                continue
            rows = np.memmap(e.index_path, dtype="<f4", mode="r", offset=_FAISS_DATA_OFFSET,
                             shape=(e.n_rows, self.d))
            self.index.add(np.ascontiguousarray(rows[a - e.row0:b - e.row0]))
            del rows
        self.index.set_id_base(self.lo)
        self.chunks = ChunkStore([os.path.join(self.storage_dir, f"{e.doc_id}_chunks.json") for e in self.docs],
                                 [e.n_rows for e in self.docs])
        if with_bm25:
            parts = []
            for e in self.docs:
                p = csr_sidecar_path(self.storage_dir, e.doc_id)
                if not os.path.exists(p):
                    parts = None
                    break
                z = np.load(p, allow_pickle=True)
                if len(z["doc_len"]) != e.n_rows:
                    raise RuntimeError(f"{p}: {len(z['doc_len'])} docs, index has {e.n_rows} rows")
                parts.append({k: z[k] for k in ("indptr", "post_doc", "post_tf", "doc_len")} | {"words": list(z["words"])})
            if parts is not None:
                indptr, pd, tf, dl, words, df_g, n_g, avgdl_g = merge_doc_csrs(parts, self.row0[:-1], self.lo, self.hi)
                kw = {} if device is None else {"device": device}
                self.bm25 = BM25Index.from_csr(indptr, pd, tf, dl, len(indptr) - 1, n_docs_global=n_g,
                                               avgdl_global=avgdl_g, df_global=df_g, **kw)
                self.bm25.set_id_base(self.lo)
                self.vocab = Vocabulary()
                self.vocab.word_to_id = {w: i for i, w in enumerate(words)}

    # -- global row <-> (document, local row) ------------------------------------------------------------
    def locate(self, row: int) -> Tuple[str, int]:
        f = int(np.searchsorted(self.row0, int(row), side="right")) - 1
        if f < 0 or f >= len(self.docs):
            raise IndexError(row)
        return self.docs[f].doc_id, int(row) - int(self.row0[f])

    def global_row(self, doc_id: str, local_row: int) -> int:
        for e in self.docs:
            if e.doc_id == doc_id:
                if not 0 <= local_row < e.n_rows:
                    raise IndexError(local_row)
                return e.row0 + int(local_row)
        raise KeyError(doc_id)

    def hit_dict(self, row: int, score: float) -> Optional[dict]:
        c = self.chunks.get(row)
        if c is None:
            return None
        meta = c.get("metadata", {})
        return {"chunk_id": c.get("chunk_id", f"unknown_{row}"), "text": c.get("text", ""), "score": float(score),
                "page": c.get("page", 0), "chapter": meta.get("chapter"), "section": meta.get("section"),
                "subsection": meta.get("subsection"), "title": meta.get("title"),
                "source_filename": meta.get("source_filename")}
