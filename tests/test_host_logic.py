"""CPU: host-side logic of the package (no GPU compute): query CSR packing, CSR build, sharding
arithmetic, vocabulary/tokenisation, config defaults, storage helpers."""
import json
import os

import numpy as np
import pytest

import intool_rag_b200  # noqa: F401
from intool_rag_b200 import bm25 as pbm25
from intool_rag_b200 import config as pconfig
from intool_rag_b200 import sharded, synth, storage, retriever
from oracle import bm25 as obm25


def test_query_csr_ragged_and_passthrough():
    ip, tm = pbm25.query_csr([[3, 1, 3], [], [7]])
    assert ip.tolist() == [0, 3, 3, 4] and tm.tolist() == [3, 1, 3, 7]
    assert ip.dtype == np.int32 and tm.dtype == np.int32
    ip2, tm2 = pbm25.query_csr((ip, tm))
    assert ip2.tolist() == ip.tolist() and tm2.tolist() == tm.tolist()
    oip, otm = obm25.to_query_csr([[3, 1, 3], [], [7]])
    assert oip.tolist() == ip.tolist() and otm.tolist() == tm.tolist()


def test_build_csr_matches_oracle():
    t, dd, dl = synth.sparse_corpus_np(500, 64, mean_len=30.0)
    indptr, pd, tf = pbm25.build_csr(t, dd, 500, 64)
    c = obm25.BM25Corpus.from_token_matrix(t, dd, dl, 64)
    assert np.array_equal(indptr, c.indptr) and np.array_equal(pd, c.post_doc) and np.array_equal(tf, c.post_tf)
    for v in range(64):  # ascending doc ids inside every posting list
        seg = pd[indptr[v]:indptr[v + 1]]
        assert (np.diff(seg) > 0).all()
    assert tf.sum() == len(t)
    with pytest.raises(ValueError):
        pbm25.build_csr(np.array([64]), np.array([0]), 1, 64)


def test_shard_bounds_partition():
    for n in (0, 1, 7, 100, 10_000_001):
        for w in (1, 2, 3, 8):
            cuts = [sharded.shard_bounds(n, w, r) for r in range(w)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in cuts]
            assert max(sizes) - min(sizes) <= 1


def test_vocabulary_and_tokenize():
    v = pbm25.Vocabulary()
    assert v.encode("The quick  brown\tFox the", grow=True) == [0, 1, 2, 3, 0]
    assert v.encode("fox jumps THE") == [3, 0]          # OOV dropped, lower-cased
    assert pbm25.tokenize("A b\nC") == obm25.tokenize("A b\nC") == ["a", "b", "c"]


def test_config_defaults_are_the_reference_constants():
    c = pconfig.Config
    assert (c.VECTOR_WEIGHT, c.BM25_WEIGHT, c.RETRIEVAL_TOP_K, c.VECTOR_DIMENSION) == (0.7, 0.3, 10, 1024)
    assert c.HYBRID_SEARCH_ENABLED is True and c.CANDIDATE_DEPTH == 50
    assert retriever.candidate_depth(10) == 50 and retriever.candidate_depth(100) == 100


def test_synth_generators_are_seeded_and_shaped():
    x = synth.dense_corpus_np(200, 48)
    assert x.dtype == np.float32 and np.allclose(np.linalg.norm(x, axis=1), 1, atol=1e-5)
    assert np.array_equal(x, synth.dense_corpus_np(200, 48))
    q = synth.dense_queries_np(x, 10)
    assert q.shape == (10, 48) and np.allclose(np.linalg.norm(q, axis=1), 1, atol=1e-5)
    t, dd, dl = synth.sparse_corpus_np(300, 1000)
    assert dl.min() >= 16 and dl.max() <= 512 and len(t) == dl.sum() and t.max() < 1000
    counts = np.bincount(t, minlength=1000)
    assert counts[0] > counts[10] > counts[500]            # Zipf head
    qs = synth.sparse_queries_np(50, 1000)
    assert all(3 <= len(a) <= 12 and len(set(a)) == len(a) and min(a) >= synth.STOP_RANKS for a in qs)


def test_storage_chunk_cache_and_missing(tmp_path):
    p = tmp_path / "docX_chunks.json"
    p.write_text(json.dumps({"total": 2, "chunks": [{"chunk_id": "a", "text": "A", "page": 1},
                                                   {"chunk_id": "b", "text": "B", "page": 2}]}))
    ch = storage._load_chunk_list(str(tmp_path), "docX")
    assert [c["chunk_id"] for c in ch] == ["a", "b"]
    assert storage._load_chunk_list(str(tmp_path), "docX") is ch   # parsed once per (path, mtime)
    with pytest.raises(FileNotFoundError):
        storage._load_chunk_list(str(tmp_path), "nope")


def test_retrieve_without_default_raises():
    retriever.set_default_retriever(None)
    with pytest.raises(RuntimeError, match="no retriever configured"):
        retriever.retrieve(np.zeros((1, 4), np.float32), [[1]], 3)
