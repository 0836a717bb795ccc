"""intool-rag_b200 — B200-native hybrid retrieval backend for intool-rag's query hot path.

Python here is host-side glue over the C ABI of ``csrc/libhr_b200.so`` (include/hr_b200.h):

* :mod:`.faiss`     — faiss-shaped module (IndexFlatL2 / IndexFlatIP / read_index / write_index),
                      drop-in for the calls in /root/reference/rag/storage/faiss_index.py.
* :mod:`.bm25`      — BM25 inverted index (the reference advertises BM25 but has none).
* :mod:`.retriever` — ``retrieve(query_embeddings, query_tokens, top_k)`` (rag/query/retriever.py).
* :mod:`.storage`   — mirror of the reference's storage wrapper functions on top of :mod:`.faiss`.
* :mod:`.sharded`   — row-sharded multi-GPU retrieval (torch.distributed all-gather + merge).

There is no CPU fallback: every compute call raises RuntimeError when the CUDA library or a
B200 is missing.
"""
from . import _lib  # noqa: F401

__all__ = ["faiss", "bm25", "retriever", "storage", "sharded", "synth", "config"]
__version__ = "0.1.0"
