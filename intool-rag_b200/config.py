"""Retrieval defaults.  Names, env variables and default values mirror the reference's dead
constants (/root/reference/rag/config.py:30,41-45) so the same environment configures both."""
from __future__ import annotations

import os


class Config:
    VECTOR_DIMENSION = int(os.getenv("VECTOR_DIMENSION", "1024"))
    RETRIEVAL_TOP_K = int(os.getenv("RETRIEVAL_TOP_K", "10"))
    RETRIEVAL_MIN_SCORE = float(os.getenv("RETRIEVAL_MIN_SCORE", "0.3"))
    HYBRID_SEARCH_ENABLED = os.getenv("HYBRID_SEARCH_ENABLED", "true").lower() == "true"
    BM25_WEIGHT = float(os.getenv("BM25_WEIGHT", "0.3"))
    VECTOR_WEIGHT = float(os.getenv("VECTOR_WEIGHT", "0.7"))
    # new knobs (no reference counterpart)
    FUSION = os.getenv("FUSION", "weighted")          # 'weighted' | 'rrf'
    CANDIDATE_DEPTH = int(os.getenv("CANDIDATE_DEPTH", "50"))  # live top_chunks, page_retriever.py:81
    BM25_K1 = float(os.getenv("BM25_K1", "1.5"))
    BM25_B = float(os.getenv("BM25_B", "0.75"))
    BM25_IDF = os.getenv("BM25_IDF", "lucene")        # 'lucene' | 'okapi'
    STORAGE_DIR = os.getenv("STORAGE_DIR", "./storages")


config = Config()


def default_device() -> int:
    """One process per GPU: HR_DEVICE, else LOCAL_RANK, else torch's current device, else 0."""
    for var in ("HR_DEVICE", "LOCAL_RANK"):
        v = os.getenv(var)
        if v is not None and v != "":
            return int(v)
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0
