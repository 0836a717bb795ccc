"""Host-side mirror of the reference's storage wrapper, backed by the B200 flat index.

Same function names, arguments and error behaviour as /root/reference/rag/storage/faiss_index.py
(FAISSIndexReader :26-103, create_faiss_index :106-128, save_faiss_index :131-134,
search_faiss_by_vector :137-199, initialize_storage :202-228) so the service code calls it
unchanged.  Written against this package's :mod:`.faiss`; nothing is copied from the reference.

On top of the mirror, the hybrid service path the reference advertises (README.md:54-58) but never
shipped: `build_bm25_sidecar` (ingest: chunk texts -> BM25 index file + vocabulary next to the faiss
file) and `search_hybrid_by_vector` (query vector + query text -> fused hits, same dict shape as
`search_faiss_by_vector`).

Deliberate, documented deviations:
  * ONE logical corpus: the search functions and `initialize_storage` work on the concatenation of every
    ``*_faiss.index`` under STORAGE_DIR (sorted by file name; :mod:`.corpus`), not only on the first file the
    glob returns (faiss_index.py:162-167).  ``STORAGE_MODE=first`` restores the reference's behaviour.  With a
    single document both modes return the same hits;
  * hits are enriched by reading only the k chunks they name from ``{doc_id}_chunks.json`` (byte-offset table,
    20 bytes of host memory per chunk) instead of parsing the whole file into dicts;
  * padded hits (id -1, when limit > ntotal) are dropped instead of silently mapping to the LAST
    chunk through Python's negative indexing (faiss_index.py:180-181);
  * the index cache is keyed by (path, mtime) so a re-ingested document is seen without restart;
  * the chunk JSON is parsed once per (path, mtime) instead of once per query (faiss_index.py:175).
"""
from __future__ import annotations

import glob
import json
import logging
import os
from typing import Dict, List, Optional, Tuple

import numpy as np

from . import faiss
from .config import config

logger = logging.getLogger("intool_rag_b200.storage")

_CORPUS_CACHE: Dict[str, tuple] = {}      # storage_dir -> (signature, Corpus)
_INDEX_CACHE: Dict[Tuple[str, float], "faiss.Index"] = {}
_BM25_CACHE: Dict[Tuple[str, float], tuple] = {}
_CHUNK_CACHE: Dict[Tuple[str, float], List[dict]] = {}


def _mtime(path: str) -> float:
    try:
        return os.path.getmtime(path)
    except OSError:
        return -1.0


def storage_mode() -> str:
    return os.environ.get("STORAGE_MODE", "corpus").lower()


def get_corpus(storage_dir: Optional[str] = None, rank: int = 0, world: int = 1):
    """The logical corpus over `storage_dir` (default STORAGE_DIR), loaded into HBM once and rebuilt when an
    index file, a BM25 CSR sidecar or the shard (rank, world) changes."""
    from .corpus import Corpus, csr_sidecar_path, doc_id_of
    sd = str(storage_dir or os.environ.get("STORAGE_DIR", config.STORAGE_DIR))
    files = sorted(glob.glob(os.path.join(sd, "*_faiss.index")))
    sig = (rank, world, os.environ.get("HR_STORAGE", "f32"), config.HYBRID_SEARCH_ENABLED) + tuple(
        (p, _mtime(p), _mtime(csr_sidecar_path(sd, doc_id_of(p)))) for p in files)
    hit = _CORPUS_CACHE.get(sd)
    if hit is not None and hit[0] == sig:
        return hit[1]
    corpus = Corpus(sd, rank=rank, world=world, with_bm25=config.HYBRID_SEARCH_ENABLED)
    _CORPUS_CACHE[sd] = (sig, corpus)
    if corpus.docs:
        logger.info("Loaded corpus: %d documents, %d rows (this shard: rows [%d, %d)), d=%d, chunk table %d bytes",
                    len(corpus.docs), corpus.ntotal_global, corpus.lo, corpus.hi, corpus.d, corpus.chunks.table_bytes)
    return corpus


def _reference_scores(D) -> List[float]:
    """clamp(1 - squared_L2 / 2, 0, 1) in Python floats on the fp32 distance (faiss_index.py:86-88)."""
    return [max(0.0, min(1.0, 1.0 - (float(d) / 2.0))) for d in D]


class FAISSIndexReader:
    """Read-only index wrapper with caching (reference: faiss_index.py:26-103)."""

    def __init__(self, index_path: str):
        self.index_path = str(index_path)
        self.index = None
        self._load_index()

    def _load_index(self) -> None:
        key = (self.index_path, _mtime(self.index_path))
        hit = _INDEX_CACHE.get(key)
        if hit is not None:
            self.index = hit
            return
        try:
            self.index = faiss.read_index(self.index_path)
        except Exception as e:  # same wrapping as the reference (:60-61)
            raise RuntimeError(f"Failed to load FAISS index: {e}")
        for old in [k for k in _INDEX_CACHE if k[0] == self.index_path]:
            del _INDEX_CACHE[old]
        _INDEX_CACHE[key] = self.index
        logger.info("Loaded FAISS index: %s (d=%d, %d vectors)", self.index_path, self.index.d, self.index.ntotal)

    def search(self, query_embedding: List[float], top_k: int = 10) -> List[Tuple[int, float]]:
        """[(row id, score)], score = clamp(1 - squared_L2 / 2, 0, 1) — the reference transform
        (:86-88), evaluated in Python floats on the fp32 distance exactly as the reference does."""
        if self.index is None:
            raise RuntimeError("Index not loaded")
        q = np.array([query_embedding], dtype=np.float32)
        D, I = self.index.search(q, top_k)
        out = []
        for idx, dist in zip(I[0], D[0]):
            score = 1.0 - (float(dist) / 2.0)
            score = max(0.0, min(1.0, score))
            out.append((int(idx), float(score)))
        return out

    def get_dimension(self) -> int:
        if self.index is None:
            raise RuntimeError("Index not loaded")
        return self.index.d

    def get_size(self) -> int:
        if self.index is None:
            raise RuntimeError("Index not loaded")
        return self.index.ntotal


def create_faiss_index(embeddings) -> "faiss.Index":
    """Build an IndexFlatL2 from a list of vectors / ndarray (reference: faiss_index.py:106-128)."""
    x = np.array(embeddings, dtype=np.float32)
    if x.ndim != 2:
        raise ValueError("embeddings must be a non-empty [n, d] array")
    index = faiss.IndexFlatL2(x.shape[1])
    index.add(x)
    logger.info("Created FAISS index: %d vectors, dim=%d", index.ntotal, index.d)
    return index


def save_faiss_index(index: "faiss.Index", path: str) -> None:
    faiss.write_index(index, path)
    logger.info("Saved FAISS index to %s", path)


def _load_chunk_list(storage_dir: str, doc_id: str) -> List[dict]:
    path = os.path.join(storage_dir, f"{doc_id}_chunks.json")
    key = (path, _mtime(path))
    hit = _CHUNK_CACHE.get(key)
    if hit is not None:
        return hit
    if not os.path.exists(path):
        raise FileNotFoundError(f"Chunks not found: {path}")
    with open(path, "r", encoding="utf-8") as f:
        data = json.load(f)
    by_id = {}
    for c in data.get("chunks", []):  # the reference keys by chunk_id, then lists dict values
        by_id[c["chunk_id"]] = c
    chunks = list(by_id.values())
    for old in [k for k in _CHUNK_CACHE if k[0] == path]:
        del _CHUNK_CACHE[old]
    _CHUNK_CACHE[key] = chunks
    return chunks


async def search_faiss_by_vector(query_vector: List[float], limit: int = 50,
                                 project: Optional[str] = None) -> List[dict]:
    """Search the first ``*_faiss.index`` under STORAGE_DIR and enrich hits with chunk metadata
    (reference: faiss_index.py:137-199; `project` is accepted and ignored there too)."""
    storage_dir = os.environ.get("STORAGE_DIR", config.STORAGE_DIR)
    index_files = sorted(glob.glob(os.path.join(str(storage_dir), "*_faiss.index")))
    if not index_files:
        logger.warning("No FAISS indices found")
        return []
    if storage_mode() != "first":
        corpus = get_corpus(str(storage_dir))
        D, I = corpus.index.search(np.array([query_vector], dtype=np.float32), int(limit))
        out = []
        for row, score in zip(I[0], _reference_scores(D[0])):
            h = corpus.hit_dict(int(row), score) if row >= 0 else None
            if h is not None:
                out.append(h)
        logger.info("FAISS search returned %d results", len(out))
        return out
    index_path = index_files[0]
    reader = FAISSIndexReader(index_path)
    hits = reader.search(query_vector, top_k=limit)
    doc_id = os.path.basename(index_path)[: -len(".index")].replace("_faiss", "")
    chunks = _load_chunk_list(str(storage_dir), doc_id)
    out = []
    for row, score in hits:
        if 0 <= row < len(chunks):
            c = chunks[row]
            meta = c.get("metadata", {})
            out.append({
                "chunk_id": c.get("chunk_id", f"unknown_{row}"),
                "text": c.get("text", ""),
                "score": score,
                "page": c.get("page", 0),
                "chapter": meta.get("chapter"),
                "section": meta.get("section"),
                "subsection": meta.get("subsection"),
                "title": meta.get("title"),
                "source_filename": meta.get("source_filename"),
            })
    logger.info("FAISS search returned %d results", len(out))
    return out


async def initialize_storage() -> None:
    """Pre-load every ``*_faiss.index`` under STORAGE_DIR into HBM (reference: faiss_index.py:202-228)."""
    storage_dir = os.environ.get("STORAGE_DIR", config.STORAGE_DIR)
    if not os.path.isdir(str(storage_dir)):
        logger.warning("Storage directory not found: %s", storage_dir)
        return
    if storage_mode() != "first":
        try:
            corpus = get_corpus(str(storage_dir))
            logger.info("Initialized storage: %d documents (%d rows) concatenated into one index in HBM",
                        len(corpus.docs), corpus.ntotal_global)
        except Exception as e:
            logger.error("Storage initialization failed: %s", e)
        return
    count = 0
    for path in sorted(glob.glob(os.path.join(str(storage_dir), "*_faiss.index"))):
        try:
            FAISSIndexReader(path)
            count += 1
        except Exception as e:
            logger.error("Failed to pre-load index %s: %s", path, e)
    logger.info("Initialized storage: loaded %d indices into HBM", count)


# ---------------------------------------------------------------------------------------------------
# hybrid service path (BM25 sidecar + fused search)
# ---------------------------------------------------------------------------------------------------
def _sidecar_paths(storage_dir: str, doc_id: str) -> Tuple[str, str]:
    return (os.path.join(storage_dir, f"{doc_id}_bm25.hrb"), os.path.join(storage_dir, f"{doc_id}_vocab.json"))


def build_bm25_sidecar(doc_id: str, chunk_texts: List[str], storage_dir: Optional[str] = None):
    """Ingest side: tokenise every chunk with the reference's idiom (``text.lower().split()``,
    rag/agent/query_processor.py:26), build the BM25 index on the GPU (row i of the faiss index == doc i)
    and write ``{doc_id}_bm25.hrb`` + ``{doc_id}_vocab.json`` next to ``{doc_id}_faiss.index``.
    Returns (BM25Index, Vocabulary)."""
    from .bm25 import BM25Index, Vocabulary
    storage_dir = str(storage_dir or os.environ.get("STORAGE_DIR", config.STORAGE_DIR))
    vocab = Vocabulary()
    docs = [vocab.encode(t, grow=True) for t in chunk_texts]
    index = BM25Index.from_docs(docs, max(len(vocab), 1))
    bm_path, vocab_path = _sidecar_paths(storage_dir, doc_id)
    index.save(bm_path)
    # the document's raw CSR: what the multi-document corpus is merged from (corpus-wide idf / avgdl need tf and
    # doc_len, not the impacts folded with this document's own statistics)
    from .bm25 import build_csr
    from .corpus import csr_sidecar_path, save_doc_csr
    doc_len = np.array([len(d) for d in docs], dtype=np.int32)
    if doc_len.sum():
        t = np.concatenate([np.asarray(d, dtype=np.int32) for d in docs if len(d)])
        dd = np.repeat(np.arange(len(docs), dtype=np.int32), doc_len)
    else:
        t, dd = np.zeros(0, np.int32), np.zeros(0, np.int32)
    indptr, pd, tf = build_csr(t, dd, len(docs), max(len(vocab), 1))
    save_doc_csr(csr_sidecar_path(storage_dir, doc_id), indptr, pd, tf, doc_len, list(vocab.word_to_id.keys()))
    with open(vocab_path, "w", encoding="utf-8") as f:
        json.dump({"n_docs": len(docs), "words": list(vocab.word_to_id.keys())}, f, ensure_ascii=False)
    logger.info("Saved BM25 sidecar: %s (%d docs, %d terms)", bm_path, len(docs), len(vocab))
    return index, vocab


def _load_bm25_sidecar(storage_dir: str, doc_id: str):
    from .bm25 import BM25Index, Vocabulary
    bm_path, vocab_path = _sidecar_paths(storage_dir, doc_id)
    if not (os.path.exists(bm_path) and os.path.exists(vocab_path)):
        return None
    key = (bm_path, _mtime(bm_path))
    hit = _BM25_CACHE.get(key)
    if hit is not None:
        return hit
    with open(vocab_path, "r", encoding="utf-8") as f:
        words = json.load(f)["words"]
    vocab = Vocabulary()
    vocab.word_to_id = {w: i for i, w in enumerate(words)}
    pair = (BM25Index.load(bm_path), vocab)
    for old in [k for k in _BM25_CACHE if k[0] == bm_path]:
        del _BM25_CACHE[old]
    _BM25_CACHE[key] = pair
    return pair


async def search_hybrid_by_vector(query_vector: List[float], query_text: str, limit: int = 10,
                                  project: Optional[str] = None) -> List[dict]:
    """Hybrid variant of `search_faiss_by_vector`: dense + BM25 + weighted fusion (weights and defaults from
    rag/config.py:41-45) over the first ``*_faiss.index`` and its BM25 sidecar; without a sidecar it is the
    dense search.  `score` is the fused score; hits carry the same chunk metadata keys."""
    from .retriever import HybridRetriever
    storage_dir = str(os.environ.get("STORAGE_DIR", config.STORAGE_DIR))
    index_files = sorted(glob.glob(os.path.join(storage_dir, "*_faiss.index")))
    if not index_files:
        logger.warning("No FAISS indices found")
        return []
    q = np.array([query_vector], dtype=np.float32)
    limit = int(limit)
    if storage_mode() != "first":
        from .bm25 import cap_query_terms
        corpus = get_corpus(storage_dir)
        tokens = None
        if corpus.bm25 is not None:
            tokens = [cap_query_terms(corpus.vocab.encode(query_text or ""))]
        scores, ids = HybridRetriever(corpus.index, corpus.bm25).retrieve(
            q, tokens, top_k=limit, k_c=min(128, max(limit, config.CANDIDATE_DEPTH)))
        out = []
        for row, score in zip(ids[0], scores[0]):
            h = corpus.hit_dict(int(row), float(score)) if row >= 0 else None
            if h is not None:
                out.append(h)
        logger.info("Hybrid search returned %d results", len(out))
        return out
    index_path = index_files[0]
    reader = FAISSIndexReader(index_path)
    doc_id = os.path.basename(index_path)[: -len(".index")].replace("_faiss", "")
    side = _load_bm25_sidecar(storage_dir, doc_id) if config.HYBRID_SEARCH_ENABLED else None
    if side is None:
        engine, tokens = HybridRetriever(reader.index, None), None
    else:
        from .bm25 import cap_query_terms
        # a long query text keeps its first 64 distinct in-vocabulary words (the library rejects more)
        engine, tokens = HybridRetriever(reader.index, side[0]), [cap_query_terms(side[1].encode(query_text or ""))]
    scores, ids = engine.retrieve(q, tokens, top_k=limit, k_c=min(128, max(limit, config.CANDIDATE_DEPTH)))
    chunks = _load_chunk_list(storage_dir, doc_id)
    out = []
    for row, score in zip(ids[0], scores[0]):
        row = int(row)
        if 0 <= row < len(chunks):
            c = chunks[row]
            meta = c.get("metadata", {})
            out.append({
                "chunk_id": c.get("chunk_id", f"unknown_{row}"),
                "text": c.get("text", ""),
                "score": float(score),
                "page": c.get("page", 0),
                "chapter": meta.get("chapter"),
                "section": meta.get("section"),
                "subsection": meta.get("subsection"),
                "title": meta.get("title"),
                "source_filename": meta.get("source_filename"),
            })
    logger.info("Hybrid search returned %d results", len(out))
    return out
