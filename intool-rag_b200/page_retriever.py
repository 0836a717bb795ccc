"""Page-level retrieval on top of the hybrid engine: the caller side of the hot path.

Mirrors the interface of /root/reference/rag/query/page_retriever.py (RetrievedChunk :26, PageRanking :36,
PageLevelRetriever :78 with retrieve_chunks :92, group_chunks_by_page :145, rank_pages :166, select_top_pages
:215, retrieve_and_rank_pages :238, module function :271) so the service code keeps its call sites; written
against this package's storage mirror, nothing is copied from the reference.

Differences, all additive:
  * the query embedding comes from an `embed` callable handed to the retriever (the reference reaches for its
    provider singleton, rag/llm/embeddings/factory.py:10; embedding models are out of scope here);
  * with `hybrid=True` (default when the corpus has a BM25 index) the chunks come from
    `search_hybrid_by_vector`, i.e. the query text takes part in the ranking;
  * `rank_pages_batch` ranks the pages of a whole batch of hit lists on the device (hr_rank_pages): same
    numbers as rank_pages (double precision, the reference's summation order), no host round-trip per query.
"""
from __future__ import annotations

import inspect
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional

from . import _lib, storage

PAGE_CHUNK_BOOST = 0.05      # per chunk on the page ...
PAGE_BOOST_CAP = 0.15        # ... up to this much (rag/query/page_retriever.py:193)


@dataclass
class RetrievedChunk:
    chunk_id: str
    text: str
    score: float
    page: int
    metadata: Dict[str, Any] = field(default_factory=dict)


@dataclass
class PageRanking:
    page: int
    score: float
    chunks: List[RetrievedChunk]
    metadata: Dict[str, Any] = field(default_factory=dict)

    def get_context_text(self) -> str:
        head = [f"{label} {self.metadata[key]}" if label else str(self.metadata[key])
                for key, label in (("chapter", "Chapter"), ("section", "Section"), ("title", "")) if self.metadata.get(key)]
        parts = ([f"[{' | '.join(head)}]", ""] if head else [])
        for c in self.chunks:
            parts += [c.text, ""]
        return "\n".join(parts).strip()

    def to_citation(self) -> Dict[str, Any]:
        m = self.metadata
        return {"page": self.page, "chapter": m.get("chapter"), "section": m.get("section"),
                "subsection": m.get("subsection"), "title": m.get("title"), "source_file": m.get("source_filename"),
                "relevance_score": round(self.score, 3)}


def page_score(scores: List[float]) -> float:
    """mean of the chunk scores + a coverage boost of 0.05 per chunk, capped at 0.15."""
    return sum(scores) / len(scores) + min(len(scores) * PAGE_CHUNK_BOOST, PAGE_BOOST_CAP)


class PageLevelRetriever:
    def __init__(self, top_chunks: int = 50, top_pages: int = 5, embed: Optional[Callable] = None,
                 hybrid: Optional[bool] = None):
        self.top_chunks, self.top_pages, self.embed, self.hybrid = int(top_chunks), int(top_pages), embed, hybrid

    async def _embedding(self, query: str) -> List[float]:
        if self.embed is None:
            raise RuntimeError("PageLevelRetriever needs an `embed` callable (query text -> vector)")
        v = self.embed(query)
        if inspect.isawaitable(v):
            v = await v
        return [float(x) for x in v]

    async def retrieve_chunks(self, query: str, project: Optional[str] = None) -> List[RetrievedChunk]:
        vec = await self._embedding(query)
        if self.hybrid is False:
            hits = await storage.search_faiss_by_vector(vec, limit=self.top_chunks, project=project)
        else:
            hits = await storage.search_hybrid_by_vector(vec, query, limit=self.top_chunks, project=project)
        keys = ("chapter", "section", "subsection", "title", "source_filename", "doc_id")
        return [RetrievedChunk(chunk_id=h.get("chunk_id", "unknown"), text=h.get("text", ""), score=h.get("score", 0),
                               page=h.get("page", 0), metadata={k: h.get(k) for k in keys}) for h in hits]

    def group_chunks_by_page(self, chunks: List[RetrievedChunk]) -> Dict[int, List[RetrievedChunk]]:
        pages: Dict[int, List[RetrievedChunk]] = {}
        for c in chunks:
            pages.setdefault(c.page, []).append(c)
        return pages

    def rank_pages(self, chunks_by_page: Dict[int, List[RetrievedChunk]]) -> List[PageRanking]:
        ranked = [PageRanking(page=p, score=page_score([c.score for c in cs]), chunks=cs, metadata=cs[0].metadata)
                  for p, cs in chunks_by_page.items()]
        ranked.sort(key=lambda r: r.score, reverse=True)       # stable: ties keep the order of first appearance
        return ranked

    def select_top_pages(self, rankings: List[PageRanking], max_pages: Optional[int] = None) -> List[PageRanking]:
        return rankings[:(max_pages or self.top_pages)]

    async def retrieve_and_rank_pages(self, query: str, project: Optional[str] = None,
                                      max_pages: Optional[int] = None) -> List[PageRanking]:
        chunks = await self.retrieve_chunks(query, project)
        if not chunks:
            return []
        return self.select_top_pages(self.rank_pages(self.group_chunks_by_page(chunks)), max_pages)


async def retrieve_and_rank_pages(query: str, project: Optional[str] = None, top_pages: int = 5,
                                  embed: Optional[Callable] = None) -> List[PageRanking]:
    return await PageLevelRetriever(top_pages=top_pages, embed=embed).retrieve_and_rank_pages(query, project, top_pages)


def rank_pages_batch(scores, ids, page_of_row, top_pages: int = 5, id_base: int = 0, l2_distances: bool = False):
    """Page ranking of a batch of hit lists on the device.  scores / ids: [nq, k] torch CUDA (as returned by
    retrieve() or Index.search()); page_of_row: int32 [n_rows] torch CUDA (row -> page, ChunkStore.page_of).
    l2_distances: `scores` are squared L2 distances of an IndexFlatL2 and the hit score is the wrapper's
    clamp(1 - d/2, 0, 1).  Returns (pages int32[nq, top_pages] (-1 padded), page scores float64, chunk counts)."""
    import torch
    s = scores.to(torch.float32).contiguous()
    i = ids.to(torch.int64).contiguous()
    pg = page_of_row.to(torch.int32).contiguous()
    nq, k = s.shape
    dev = s.device
    oP = torch.empty((nq, top_pages), dtype=torch.int32, device=dev)
    oS = torch.empty((nq, top_pages), dtype=torch.float64, device=dev)
    oC = torch.empty((nq, top_pages), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().hr_rank_pages(s.data_ptr(), i.data_ptr(), nq, k, 1 if l2_distances else 0, pg.data_ptr(),
                                        int(pg.numel()), int(id_base), int(top_pages), oP.data_ptr(), oS.data_ptr(),
                                        oC.data_ptr(), dev.index or 0, _lib.current_stream_ptr(dev.index or 0)))
    return oP, oS, oC
