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


def test_chunk_spans_follow_the_reference_row_mapping(tmp_path):
    """row -> chunk = list({c["chunk_id"]: c ...}.values())[row] (/root/reference/rag/storage/faiss_index.py:175-181):
    first position, last content of a repeated chunk_id; non-ASCII text (the reference saves with
    ensure_ascii=False, file_storage.py:131-132) must not shift the byte offsets."""
    from intool_rag_b200 import corpus
    chunks = [{"chunk_id": f"c{i}", "page": i, "text": f"héllo wörld {i} ✓ \"quoted\" ]", "chunk_index": i} for i in range(6)]
    chunks.append({"chunk_id": "c1", "page": 99, "text": "dup wins", "chunk_index": 1})
    for kw in ({"indent": 2, "ensure_ascii": False}, {"ensure_ascii": True}, {"separators": (",", ":"), "ensure_ascii": False}):
        p = tmp_path / "a_chunks.json"
        p.write_text(json.dumps({"total": 6, "chunks": chunks}, **kw), encoding="utf-8")
        want = list({c["chunk_id"]: c for c in chunks}.values())
        store = corpus.ChunkStore([str(p)], [len(want) + 2])
        assert [store.get(i) for i in range(len(want))] == want
        assert store.get(len(want)) is None and store.get(-1) is None and store.get(10 ** 6) is None
        assert store.table_bytes == (len(want) + 2) * 20
        store.close()


def test_merge_doc_csrs_equals_one_build_over_the_concatenation():
    from intool_rag_b200 import corpus
    rng = np.random.default_rng(0)
    parts, row0, docs_all, r = [], [], [], 0
    for f in range(4):
        n = int(rng.integers(3, 9))
        words = [f"w{j}" for j in rng.permutation(14)[:8]]
        docs = [[int(x) for x in rng.integers(0, 8, size=int(rng.integers(1, 9)))] for _ in range(n)]
        dl = np.array([len(x) for x in docs], np.int32)
        t = np.concatenate([np.array(x, np.int32) for x in docs])
        ip, pd, tf = pbm25.build_csr(t, np.repeat(np.arange(n, dtype=np.int32), dl), n, 8)
        parts.append(dict(indptr=ip, post_doc=pd, post_tf=tf, doc_len=dl, words=words))
        row0.append(r)
        r += n
        docs_all += [[words[x] for x in d] for d in docs]
    for lo, hi in ((0, r), (4, r - 3), (r - 1, r), (2, 2)):
        ip, pd, tf, dl, words, df_g, n_g, avg = corpus.merge_doc_csrs(parts, row0, lo, hi)
        wid = {w: i for i, w in enumerate(words)}
        t = np.concatenate([np.array([wid[w] for w in d], np.int32) for d in docs_all])
        dlen = np.array([len(d) for d in docs_all])
        dd = np.repeat(np.arange(r, dtype=np.int32), dlen)
        full = pbm25.build_csr(t, dd, r, len(words))
        m = (dd >= lo) & (dd < hi)
        ip2, pd2, tf2 = pbm25.build_csr(t[m], dd[m] - lo, max(hi - lo, 1), len(words))
        assert np.array_equal(ip, ip2) and np.array_equal(pd, pd2) and np.array_equal(tf, tf2)
        assert np.array_equal(df_g, np.diff(full[0])) and n_g == r and avg == pytest.approx(dlen.mean(), rel=1e-15)
        assert np.array_equal(dl, dlen[lo:hi])


def test_flat_header_reader(tmp_path):
    from intool_rag_b200 import corpus
    from oracle import flat
    ix = flat.IndexFlatL2(6)
    ix.add(np.arange(30, dtype=np.float32).reshape(5, 6))
    p = tmp_path / "d_faiss.index"
    flat.write_index(ix, str(p))
    assert corpus.read_flat_header(str(p)) == (6, 5, 1) and corpus.doc_id_of(str(p)) == "d"
    rows = np.memmap(str(p), dtype="<f4", mode="r", offset=45, shape=(5, 6))
    assert np.array_equal(np.asarray(rows), ix._x)
    p.write_bytes(p.read_bytes()[:60])
    with pytest.raises(RuntimeError, match="truncated"):
        corpus.read_flat_header(str(p))
