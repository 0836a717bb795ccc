"""GPU parity of the ingest-side pieces: the device CSR build from token occurrences equals
the host builder bit for bit (integer work), and the BM25 index file round-trips."""
import numpy as np
import pytest

import intool_rag_b200  # noqa: F401
from intool_rag_b200 import bm25 as pbm25
from intool_rag_b200 import synth
from oracle import bm25 as obm25

pytestmark = pytest.mark.gpu


def _same_search(a, b, qs, k=20):
    Sa, Ia = a.search(qs, k)
    Sb, Ib = b.search(qs, k)
    assert np.array_equal(Ia, Ib) and np.array_equal(Sa, Sb)
    return Sa, Ia


def test_device_csr_build_equals_host_build(gpu):
    import torch
    n, V = 50000, 4000
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=40.0)
    rng = np.random.default_rng(3)
    perm = rng.permutation(t.size)                      # token order must not matter
    qs = synth.sparse_queries_np(64, V, stop=8)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    host = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V)
    dev_np = pbm25.BM25Index.from_tokens(t[perm].astype(np.int32), dd[perm].astype(np.int32), n, V)
    dev_t = pbm25.BM25Index.from_tokens(torch.from_numpy(t[perm].astype(np.int32)).cuda(),
                                        torch.from_numpy(dd[perm].astype(np.int32)).cuda(), n, V)
    assert host.nnz == dev_np.nnz == dev_t.nnz == len(pd)
    S, I = _same_search(host, dev_np, qs)
    _same_search(host, dev_t, qs)
    o = obm25.BM25Corpus.from_token_matrix(t, dd, dl, V)
    Sr, Ir = o.search(qs, 20)
    np.testing.assert_allclose(S, Sr, rtol=1e-5, atol=1e-7)
    # docs as ragged token lists (the service adapter's path) go through the same device build
    docs = [[] for _ in range(200)]
    for term, doc in zip(t[dd < 200], dd[dd < 200]):
        docs[int(doc)].append(int(term))
    a = pbm25.BM25Index.from_docs(docs, V)
    ip2, pd2, tf2 = pbm25.build_csr(t[dd < 200], dd[dd < 200], 200, V)
    b = pbm25.BM25Index.from_csr(ip2, pd2, tf2, dl[:200], V)
    _same_search(a, b, qs[:16], 10)


def test_device_csr_build_rejects_bad_ids_and_handles_empty(gpu):
    with pytest.raises(RuntimeError, match="outside"):
        pbm25.BM25Index.from_tokens(np.array([0, 7], np.int32), np.array([0, 1], np.int32), 2, 5)
    with pytest.raises(RuntimeError, match="outside"):
        pbm25.BM25Index.from_tokens(np.array([0, 1], np.int32), np.array([0, 2], np.int32), 2, 5)
    e = pbm25.BM25Index.from_tokens(np.zeros(0, np.int32), np.zeros(0, np.int32), 3, 5)
    assert e.nnz == 0 and e.ndocs == 3
    S, I = e.search([[1, 2]], 4)
    assert (I == -1).all() and (S == 0).all()


def test_bm25_index_file_roundtrip(gpu, tmp_path):
    n, V = 30000, 2500
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=32.0)
    qs = synth.sparse_queries_np(40, V, stop=8)
    a = pbm25.BM25Index.from_tokens(t.astype(np.int32), dd.astype(np.int32), n, V, idf="okapi")
    a.set_id_base(1000)
    path = tmp_path / "corpus_bm25.hrb"
    a.save(str(path))
    b = pbm25.BM25Index.load(str(path))
    assert (b.ndocs, b.vocab, b.nnz) == (a.ndocs, a.vocab, a.nnz)
    S, I = _same_search(a, b, qs)
    assert I[I >= 0].min() >= 1000                      # id_base travels with the file
    bad = tmp_path / "junk.hrb"
    bad.write_bytes(b"not an index")
    with pytest.raises(RuntimeError, match="HRBM25"):
        pbm25.BM25Index.load(str(bad))
    trunc = tmp_path / "trunc.hrb"
    trunc.write_bytes(path.read_bytes()[:-100])
    with pytest.raises(RuntimeError, match="truncated"):
        pbm25.BM25Index.load(str(trunc))
