"""GPU: the robustness rules of the C ABI (include/hr_b200.h): per-handle serialisation, long queries, malformed
CSR / sidecar files, duplicate ids and NaN in the merges, the device-driven certificate fallback beyond its
capacity, and hr_retrieve_sharded with a world of one (no NCCL, no torch.distributed)."""
import ctypes as C
import threading

import numpy as np
import pytest

import intool_rag_b200  # noqa: F401
from intool_rag_b200 import _lib, synth
from intool_rag_b200 import bm25 as pbm25
from intool_rag_b200 import faiss as hf
from intool_rag_b200.retriever import HybridRetriever
from oracle import bm25 as obm25
from oracle import flat

pytestmark = pytest.mark.gpu


def _small_hybrid(n=20000, d=64, V=600, seed=0):
    x = synth.dense_corpus_np(n, d, seed=synth.DENSE_SEED + seed)
    t, dd, dl = synth.sparse_corpus_np(n, V, seed=synth.SPARSE_SEED + seed, mean_len=24.0)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    ix = hf.IndexFlatIP(d)
    ix.add(x)
    bm = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V)
    return x, (t, dd, dl), ix, bm


def test_two_threads_share_one_handle(gpu):
    """faiss-cpu's IndexFlat.search is thread-safe; ctypes drops the GIL, so two threads really overlap here.
    The per-handle mutex serialises them: every answer equals the single-threaded one."""
    x, _, ix, bm = _small_hybrid()
    V = bm.vocab
    eng = HybridRetriever(ix, bm)
    batches = []
    for s in range(8):
        q = synth.dense_queries_np(x, 24 + s, seed=100 + s)
        qs = synth.sparse_queries_np(24 + s, V, seed=200 + s, stop=8)
        batches.append((q, qs, eng.retrieve(q, qs, 10), ix.search(q, 7)))
    errors = []

    def worker(order):
        try:
            for _ in range(6):
                for i in order:
                    q, qs, (S0, I0), (D0, J0) = batches[i]
                    S, I = eng.retrieve(q, qs, 10)
                    D, J = ix.search(q, 7)
                    if not (np.array_equal(I, I0) and np.array_equal(S, S0) and np.array_equal(J, J0)
                            and np.array_equal(D, D0)):
                        errors.append(i)
        except Exception as e:   # noqa: BLE001
            errors.append(repr(e))

    th = [threading.Thread(target=worker, args=(o,)) for o in ([0, 1, 2, 3, 4, 5, 6, 7], [7, 5, 3, 1, 6, 4, 2, 0],
                                                                 [3, 3, 0, 7, 1, 1, 5, 2])]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors[:5]


def test_long_queries_same_rule_on_host_and_device_input(gpu):
    """Any number of raw terms is fine (duplicates fold into multiplicities); more than 64 DISTINCT scorable
    terms is HR_ERR_INVALID for numpy and for torch CUDA queries alike (round 1 truncated the device path)."""
    import torch
    n, V = 5000, 400
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=20.0)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    bm = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V)
    o = obm25.BM25Corpus.from_token_matrix(t, dd, dl, V)
    rng = np.random.default_rng(5)
    long_dup = [int(v) for v in rng.integers(20, 50, size=300)]          # 300 raw terms, 30 distinct
    sixty_four = [int(v) for v in rng.permutation(np.arange(10, 200))[:64]]
    qs = [long_dup, sixty_four, [3, 3, 3]]
    Sr, Ir = o.search(qs, 20)
    S, I = bm.search(qs, 20)
    np.testing.assert_allclose(S, Sr, rtol=1e-5, atol=1e-7)
    ip, tm = pbm25.query_csr(qs)
    Sd, Id = bm.search((torch.from_numpy(ip).cuda(), torch.from_numpy(tm).cuda()), 20)
    assert np.array_equal(Id.cpu().numpy(), I) and np.array_equal(Sd.cpu().numpy(), S)
    too_many = [list(range(10, 110))]                                      # 100 distinct in-vocabulary terms
    with pytest.raises(RuntimeError, match="64 distinct"):
        bm.search(too_many, 10)
    ip, tm = pbm25.query_csr(too_many)
    with pytest.raises(RuntimeError, match="64 distinct"):
        bm.search((torch.from_numpy(ip).cuda(), torch.from_numpy(tm).cuda()), 10)
    # the handle stays usable, and the service adapter's cap keeps the first 64 distinct words
    S2, I2 = bm.search(qs, 20)
    assert np.array_equal(I2, I)
    capped = pbm25.cap_query_terms(too_many[0] + [10, 11])
    assert len(set(capped)) == 64 and capped[-2:] == [10, 11]
    Sc, Ic = bm.search([capped], 10)
    Sor, Ior = o.search([capped], 10)
    np.testing.assert_allclose(Sc, Sor, rtol=1e-5, atol=1e-7)


def test_malformed_csr_and_corrupt_sidecar_are_rejected(gpu, tmp_path):
    n, V = 300, 20
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=12.0)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    good = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V)
    bad_order = pd.copy()
    a = int(indptr[3])
    bad_order[a], bad_order[a + 1] = bad_order[a + 1], bad_order[a]        # descending pair inside a list
    with pytest.raises(RuntimeError, match="bad CSR"):
        pbm25.BM25Index.from_csr(indptr, bad_order, tf, dl, V)
    bad_range = pd.copy()
    bad_range[int(indptr[5]) + 2] = n + 7                                   # doc id outside [0, n)
    with pytest.raises(RuntimeError, match="bad CSR"):
        pbm25.BM25Index.from_csr(indptr, bad_range, tf, dl, V)
    bad_tf = tf.copy()
    bad_tf[0] = 0
    with pytest.raises(RuntimeError, match="bad CSR"):
        pbm25.BM25Index.from_csr(indptr, pd, bad_tf, dl, V)
    bad_ip = indptr.copy()
    bad_ip[4], bad_ip[5] = bad_ip[5], bad_ip[4] - 1
    with pytest.raises(RuntimeError, match="non-decreasing"):
        pbm25.BM25Index.from_csr(bad_ip, pd, tf, dl, V)
    # sidecar: a flipped doc id in the posting region must fail the load-time check
    path = tmp_path / "x_bm25.hrb"
    good.save(str(path))
    raw = bytearray(path.read_bytes())
    hdr = 8 + 5 * 8 + (V + 1) * 8 + V * 4
    first = int.from_bytes(raw[hdr:hdr + 4], "little")
    second = int.from_bytes(raw[hdr + 4:hdr + 8], "little")
    raw[hdr:hdr + 4] = second.to_bytes(4, "little")
    raw[hdr + 4:hdr + 8] = first.to_bytes(4, "little")
    bad = tmp_path / "y_bm25.hrb"
    bad.write_bytes(bytes(raw))
    with pytest.raises(RuntimeError, match="corrupt BM25 index file"):
        pbm25.BM25Index.load(str(bad))
    assert pbm25.BM25Index.load(str(path)).nnz == good.nnz


def test_merge_topk_with_duplicate_ids_and_nan_writes_every_slot(gpu):
    """Overlapping shards (id_base not set) deliver equal (score, id) pairs, and a NaN has no place in the
    order: every output slot must still be written (round 1 left slots uninitialised)."""
    import torch
    nq, L, k = 4, 24, 24
    S = np.tile(np.linspace(1.0, 0.1, 12, dtype=np.float32), (nq, 2))      # every (score, id) appears twice
    I = np.tile(np.arange(12, dtype=np.int64), (nq, 2))
    S[1, 5] = np.nan
    Sd, Id = torch.from_numpy(S).cuda(), torch.from_numpy(I).cuda()
    oS = torch.full((nq, k), -777.0, device="cuda")
    oI = torch.full((nq, k), -777, dtype=torch.int64, device="cuda")
    _lib.check(_lib.lib().hr_merge_topk(Sd.data_ptr(), Id.data_ptr(), nq, L, k, 1, -5.0, oS.data_ptr(), oI.data_ptr(),
                                        0, _lib.current_stream_ptr(0)))
    torch.cuda.synchronize()
    oS, oI = oS.cpu().numpy(), oI.cpu().numpy()
    assert not (oI == -777).any() and not (oS == -777.0).any() and not np.isnan(oS).any()
    assert oI[0].tolist() == [i // 2 for i in range(24)]
    assert (np.diff(oS, axis=1) <= 0).all()
    assert oS[1, -1] == -5.0 and oI[1, -1] == 5                            # the NaN entry ranks like padding


def test_more_fallback_queries_than_the_device_side_capacity(gpu):
    """All rows identical: every query's certificate fails (massive ties).  The device-driven fallback handles 64
    queries per search, the rest is finished after the synchronisation: answers equal the exact scan."""
    import torch
    n, d, nq, k = 6000, 32, 150, 10
    row = np.ones((1, d), np.float32) / np.sqrt(d)
    x = np.repeat(row, n, axis=0)
    x[1234] *= 1.5
    q = np.repeat(row, nq, axis=0)
    ix = hf.IndexFlatIP(d)
    ix.add(x)
    D, I = ix.search(q, k)
    st = ix.stats()
    assert st["flagged"] > 64, st
    assert (I[:, 0] == 1234).all() and (I[:, 1:] == np.arange(0, k - 1)).all()   # (score desc, id asc)
    ix.set_mode("exact")
    De, Ie = ix.search(q, k)
    assert np.array_equal(I, Ie) and np.array_equal(D, De)
    ix.set_mode("auto")
    # same through the hybrid call and with device tensors
    S, J = HybridRetriever(ix, None).retrieve(torch.from_numpy(q).cuda(), None, 5)
    assert (J.cpu().numpy()[:, 0] == 1234).all()


def test_retrieve_sharded_world_of_one_through_the_c_abi(gpu):
    """hr_comm_init(world = 1) + hr_retrieve_sharded need neither NCCL nor torch.distributed and must equal
    hr_retrieve; a null comm is an error, not a crash."""
    x, _, ix, bm = _small_hybrid(n=12000, seed=3)
    nq = 17
    q = synth.dense_queries_np(x, nq, seed=9)
    qs = synth.sparse_queries_np(nq, bm.vocab, seed=10, stop=8)
    S0, I0 = HybridRetriever(ix, bm).retrieve(q, qs, 10)
    L = _lib.lib()
    comm = C.c_void_p()
    _lib.check(L.hr_comm_init(None, 0, 1, 0, C.byref(comm)))
    assert L.hr_comm_world(comm) == 1 and L.hr_comm_rank(comm) == 0
    qi, qt = pbm25.query_csr(qs)
    S = np.empty((nq, 10), np.float32)
    I = np.empty((nq, 10), np.int64)
    _lib.check(L.hr_retrieve_sharded(comm, ix._h, bm._h, q.ctypes.data, qi.ctypes.data, qt.ctypes.data, nq, int(qt.size),
                                     10, 50, 0, 0.7, 0.3, S.ctypes.data, I.ctypes.data, 0, None))
    assert np.array_equal(I, I0) and np.array_equal(S, S0)
    assert L.hr_retrieve_sharded(None, ix._h, bm._h, q.ctypes.data, qi.ctypes.data, qt.ctypes.data, nq, int(qt.size),
                                 10, 50, 0, 0.7, 0.3, S.ctypes.data, I.ctypes.data, 0, None) == -1
    assert "null comm" in _lib.last_error()
    assert L.hr_comm_init(None, 0, 2, 0, C.byref(C.c_void_p())) == -1     # world > 1 needs a unique id
    _lib.check(L.hr_comm_destroy(comm))


def test_unsupported_dimension_is_rejected_at_create(gpu):
    with pytest.raises(RuntimeError, match="6400"):
        hf.IndexFlatL2(7000)
    assert hf.IndexFlatL2(6400).d == 6400


def test_k_above_128_runs_the_exact_scan(gpu):
    x = synth.dense_corpus_np(5000, 48)
    q = synth.dense_queries_np(x, 5)
    ix = hf.IndexFlatIP(48)
    ix.add(x)
    D, I = ix.search(q, 300)
    assert ix.stats()["mode_used"] == _lib.MODE_EXACT_SIMT
    o = flat.IndexFlatIP(48)
    o.add(x)
    Dr, Ir = o.search(q, 300)
    assert (I == Ir).mean() > 0.999
    np.testing.assert_allclose(D, Dr, atol=3e-6)


@pytest.mark.parametrize("storage", ["f32", "f32+bf16", "bf16"])
@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_non_finite_inputs_follow_the_nan_rule(gpu, storage, metric):
    """NaN scores are never candidates (oracle/flat.py: faiss' heap test is false for NaN): a NaN query gets padding,
    a NaN corpus row is never returned, in the tensor-core path and in the exhaustive one alike; nothing hangs."""
    rng = np.random.default_rng(9)
    n, d = 30000, 96
    x = rng.standard_normal((n, d)).astype(np.float32)
    if storage == "bf16":
        import torch
        x = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    q = rng.standard_normal((6, d)).astype(np.float32)
    o = (flat.IndexFlatIP if metric == "ip" else flat.IndexFlatL2)(d)
    o.add(x)
    _, I0 = o.search(q, 10)
    victim = int(I0[0, 0])
    x[victim, 7] = np.nan
    q[1, 2] = np.nan
    q[2, :] = 0.0
    o = (flat.IndexFlatIP if metric == "ip" else flat.IndexFlatL2)(d)
    o.add(x)
    Do, Io = o.search(q, 10, precision="f64")
    ix = (hf.IndexFlatIP if metric == "ip" else hf.IndexFlatL2)(d, storage=storage)
    ix.add(x)
    for mode in ("auto", "exact"):
        ix.set_mode(mode)
        D, I = ix.search(q, 10)
        assert (I[1] == -1).all(), mode
        assert victim not in I
        for r in (0, 3, 4, 5):      # generic rows: ids equal the fp64 oracle (gaps are far above fp32 rounding)
            assert np.array_equal(I[r], Io[r]), (mode, r)
        if metric == "ip":           # the zero query ties every finite row at 0: smallest ids first, NaN row skipped
            want = [i for i in range(12) if i != victim][:10]
            assert list(I[2]) == want, mode


def test_hybrid_batches_beyond_one_scan_launch(gpu):
    """More queries than one scan launch holds (2048): the dense part runs in chunks with a synchronisation in
    between, BM25 and the fusion take the whole batch; host and device inputs give the same answer."""
    import torch
    from oracle import hybrid
    x, (t, dd, dl), ix, bm = _small_hybrid(n=30000, d=64, V=800, seed=3)
    eng = HybridRetriever(ix, bm)
    oi = flat.IndexFlatIP(64)
    oi.add(x)
    oc = obm25.BM25Corpus.from_token_matrix(t, dd, dl, 800)
    nq = 2500
    q = synth.dense_queries_np(x, nq, seed=77)
    qs = synth.sparse_queries_np(nq, 800, seed=78, stop=8)
    S, I = eng.retrieve(q, qs, 10)
    So, Io, _ = hybrid.retrieve(oi, oc, q, qs, 10, precision="f64")
    np.testing.assert_allclose(S, So, atol=2e-6)
    diff = I != Io
    assert diff.mean() < 1e-3
    # a differing id sits inside a near-tie of fused scores (fp32 here, fp64 in the oracle)
    assert np.all(np.abs(S[diff] - So[diff]) < 2e-6)
    S2, I2 = eng.retrieve(torch.from_numpy(q).cuda(), pbm25.query_csr(qs), 10)
    assert np.array_equal(I2.cpu().numpy(), I) and np.array_equal(S2.cpu().numpy(), S)
