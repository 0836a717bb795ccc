"""GPU parity: BM25 scoring, hybrid fusion and the merges vs the oracle (1e-5 relative)."""
import json
import os

import numpy as np
import pytest

import intool_rag_b200  # noqa: F401
from intool_rag_b200 import _lib, synth
from intool_rag_b200 import bm25 as pbm25
from intool_rag_b200 import faiss as hf
from intool_rag_b200.retriever import HybridRetriever
from oracle import bm25 as obm25
from oracle import flat, fusion, hybrid

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _assert_ranked(S, I, S_ref, I_ref, what):
    assert S.shape == S_ref.shape
    for q in range(S.shape[0]):
        assert (I[q] >= 0).sum() == (I_ref[q] >= 0).sum(), f"{what} q{q}: candidate count"
        np.testing.assert_allclose(S[q], S_ref[q], rtol=RTOL, atol=1e-7, err_msg=f"{what} q{q}")
        for j in range(S.shape[1]):
            if I[q, j] != I_ref[q, j]:
                # only allowed inside a run of (near-)equal reference scores
                close = np.abs(S_ref[q] - S_ref[q, j]) <= RTOL * max(abs(S_ref[q, j]), 1e-6) * 2
                pos = np.nonzero(I_ref[q] == I[q, j])[0]
                assert (pos.size and close[pos[0]]) or close[-1], f"{what} q{q} rank {j}: {I[q, j]} vs {I_ref[q, j]}"


def test_bm25_kat(gpu, golden_dir):
    k = json.load(open(os.path.join(golden_dir, "kat_bm25.json")))
    ix = pbm25.BM25Index.from_docs(k["docs"], k["vocab"], k1=k["k1"], b=k["b"])
    assert (ix.ndocs, ix.vocab) == (3, 5)
    S, I = ix.search(k["queries"], 3)
    want = np.array(k["scores"])
    for q in range(3):
        order = [i for i in np.lexsort((np.arange(3), -want[q])) if want[q][i] > 0]
        assert I[q, :len(order)].tolist() == order and (I[q, len(order):] == -1).all()
        np.testing.assert_allclose(S[q, :len(order)], want[q][order], rtol=RTOL)
        assert (S[q, len(order):] == 0).all()


@pytest.mark.parametrize("idf", ["lucene", "okapi"])
def test_bm25_synthetic_vs_oracle(gpu, idf):
    n, V, nq = 60000, 3000, 200            # > 16384 docs: several shared-memory windows per query
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=48.0)
    qs = synth.sparse_queries_np(nq, V, stop=16)
    qs[3] = []                             # empty query
    qs[4] = [V + 5, -2]                    # only out-of-vocabulary ids
    qs[5] = qs[5] + qs[5][:2]              # duplicated terms count twice
    qs[6] = [0, 1, 2]                      # the most frequent terms: posting lists ~ every doc
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    ix = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V, idf=idf)
    assert ix.nnz == len(pd)
    o = obm25.BM25Corpus.from_token_matrix(t, dd, dl, V, idf_variant=idf)
    for k in (10, 50):
        S, I, touched = ix.search(qs, k, return_postings=True)
        Sr, Ir = o.search(qs, k)
        _assert_ranked(S, I, Sr, Ir, f"bm25/{idf}/k{k}")
        want_touched = sum(int(o.df[t_]) for q in qs for t_ in set(q) if 0 <= t_ < V)
        assert touched == want_touched
    assert (I[3] == -1).all() and (I[4] == -1).all()


def test_bm25_device_csr_and_single_query_groups(gpu):
    """CSR handed over as CUDA tensors; nq=1 runs many range groups per query (the latency mode)."""
    import torch
    n, V = 200000, 5000
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=32.0)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    ix = pbm25.BM25Index.from_csr(torch.from_numpy(indptr).cuda(), torch.from_numpy(pd).cuda(),
                                  torch.from_numpy(tf).cuda(), torch.from_numpy(dl).cuda(), V)
    o = obm25.BM25Corpus.from_token_matrix(t, dd, dl, V)
    qs = synth.sparse_queries_np(3, V, stop=8)
    for q in qs:
        S, I = ix.search([q], 50)
        Sr, Ir = o.search([q], 50)
        _assert_ranked(S, I, Sr, Ir, "bm25/nq1")
    ip, tm = pbm25.query_csr(qs)
    S, I = ix.search((torch.from_numpy(ip).cuda(), torch.from_numpy(tm).cuda()), 20)
    Sr, Ir = o.search(qs, 20)
    _assert_ranked(S.cpu().numpy(), I.cpu().numpy(), Sr, Ir, "bm25/devcsr")


def test_bm25_ties_resolve_by_id(gpu):
    docs = [[1, 2]] * 40 + [[3]] * 5
    ix = pbm25.BM25Index.from_docs(docs, 8)
    S, I = ix.search([[1]], 10)
    assert I[0].tolist() == list(range(10)) and np.all(S[0] == S[0, 0])


def test_fusion_kat(gpu, golden_dir):
    import torch
    k = json.load(open(os.path.join(golden_dir, "kat_fusion.json")))
    dS = torch.tensor(k["dense_sim"], dtype=torch.float32).cuda()
    dI = torch.tensor(k["dense_ids"], dtype=torch.int64).cuda()
    bS = torch.tensor(k["bm25"], dtype=torch.float32).cuda()
    bI = torch.tensor(k["bm25_ids"], dtype=torch.int64).cuda()
    for mode, code in (("weighted", 0), ("rrf", 1)):
        oS = torch.empty((1, 9), dtype=torch.float32).cuda()
        oI = torch.empty((1, 9), dtype=torch.int64).cuda()
        _lib.check(_lib.lib().hr_fuse(dS.data_ptr(), dI.data_ptr(), bS.data_ptr(), bI.data_ptr(), None, 1, 5, 9,
                                      0, code, k["w_vec"], k["w_bm25"], oS.data_ptr(), oI.data_ptr(), 0,
                                      _lib.current_stream_ptr(0)))
        torch.cuda.synchronize()
        assert oI[0, :7].tolist() == [t[0] for t in k[mode]] and oI[0, 7:].tolist() == [-1, -1]
        np.testing.assert_allclose(oS[0, :7].cpu().numpy(), [t[1] for t in k[mode]], rtol=RTOL)


@pytest.mark.parametrize("metric", ["ip", "l2"])
@pytest.mark.parametrize("mode", ["weighted", "rrf"])
def test_hybrid_retrieve_vs_golden(gpu, golden_dir, metric, mode):
    z = np.load(os.path.join(golden_dir, "synthetic_small.npz"))
    n, d, V, nq = int(z["n"]), int(z["d"]), int(z["V"]), int(z["nq"])
    x = synth.dense_corpus_np(n, d)
    x[100] = x[7]
    q = synth.dense_queries_np(x, nq)
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=40.0)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    ix = hf.IndexFlatL2(d) if metric == "l2" else hf.IndexFlatIP(d)
    ix.add(x)
    bm = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V)
    r = HybridRetriever(ix, bm, fusion=mode)
    S, I = r.retrieve(q, (z["q_indptr"], z["q_terms"]), 10)
    _assert_ranked(S, I, z[f"{metric}_{mode}_S"], z[f"{metric}_{mode}_I"], f"hybrid/{metric}/{mode}")
    # torch CUDA in -> torch CUDA out, identical numbers
    import torch
    S2, I2 = r.retrieve(torch.from_numpy(q).cuda(), (z["q_indptr"], z["q_terms"]), 10)
    assert np.array_equal(I2.cpu().numpy(), I) and np.array_equal(S2.cpu().numpy(), S)
    # dense-only retrieve (no tokens) = dense ranking with weighted scores
    S3, I3 = r.retrieve(q, None, 10)
    D, Id = ix.search(q, 10)
    if mode == "weighted":
        assert np.array_equal(I3, Id)


def test_c0_hybrid_config(gpu):
    """BASELINE config C0 at reduced nq (oracle BM25 is a Python loop): 100k x 1024 + 30k-term
    vocabulary, hybrid top-10."""
    n, d, V, nq = 100_000, 1024, 30_000, 48
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, nq)
    t, dd, dl = synth.sparse_corpus_np(n, V)
    qs = synth.sparse_queries_np(nq, V)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    ix = hf.IndexFlatIP(d)
    ix.add(x)
    bm = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V)
    S, I = HybridRetriever(ix, bm).retrieve(q, qs, 10)
    oi = flat.IndexFlatIP(d)
    oi.add(x)
    oc = obm25.BM25Corpus.from_token_matrix(t, dd, dl, V)
    Sr, Ir, _ = hybrid.retrieve(oi, oc, q, qs, 10)
    _assert_ranked(S, I, Sr, Ir, "C0 hybrid")


def test_two_shards_on_one_gpu_equal_single_index(gpu):
    """Row sharding emulated on one device: shard-local search with id_base + global BM25 statistics,
    candidates concatenated in rank order (what the all-gather produces), hr_merge_topk + hr_fuse ==
    the unsharded hr_retrieve, bit for bit."""
    import torch
    from intool_rag_b200.sharded import ShardedRetriever, shard_bounds
    n, d, V, nq = 30000, 64, 800, 40
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, nq)
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=30.0)
    qs = synth.sparse_queries_np(nq, V, stop=8)
    indptr, pd, tf = pbm25.build_csr(t, dd, n, V)
    full_ix = hf.IndexFlatIP(d)
    full_ix.add(x)
    full_bm = pbm25.BM25Index.from_csr(indptr, pd, tf, dl, V)
    S, I = HybridRetriever(full_ix, full_bm).retrieve(q, qs, 10)
    df = np.diff(indptr)
    outs, blocks = [], []
    for rank in range(2):
        lo, hi = shard_bounds(n, 2, rank)
        ix = hf.IndexFlatIP(d)
        ix.add(x[lo:hi])
        ix.set_id_base(lo)
        m = (dd >= lo) & (dd < hi)
        ip_l, pd_l, tf_l = pbm25.build_csr(t[m], dd[m] - lo, hi - lo, V)
        bm = pbm25.BM25Index.from_csr(ip_l, pd_l, tf_l, dl[lo:hi], V, n_docs_global=n,
                                      avgdl_global=float(dl.astype(np.float64).mean()), df_global=df)
        bm.set_id_base(lo)
        qd = torch.from_numpy(q).cuda()
        D, Id = ix.search(qd, 50)
        Sb, Ib = bm.search(pbm25.query_csr(qs), 50)
        outs.append((D, Id, torch.from_numpy(Sb).cuda(), torch.from_numpy(Ib).cuda()))
        # the packed block hr_candidates writes on this "rank" (what the single all-gather carries)
        nn = nq * 50
        blk = torch.empty(24 * nn, dtype=torch.uint8, device="cuda")
        Dv, Sv, Iv, Jv = ShardedRetriever._views(blk, nn)
        qi_np, qt_np = pbm25.query_csr(qs)
        qi_d, qt_d = torch.from_numpy(qi_np).cuda(), torch.from_numpy(qt_np).cuda()
        _lib.check(_lib.lib().hr_candidates(ix._h, bm._h, qd.data_ptr(), qi_d.data_ptr(), qt_d.data_ptr(), nq,
                                            int(qt_d.numel()), 50, Dv.data_ptr(), Iv.data_ptr(), Sv.data_ptr(),
                                            Jv.data_ptr(), _lib.current_stream_ptr(0)))
        assert torch.equal(Dv.view(nq, 50), D) and torch.equal(Iv.view(nq, 50), Id)
        assert torch.equal(Jv.view(nq, 50), outs[-1][3]) and torch.equal(Sv.view(nq, 50), outs[-1][2])
        blocks.append(blk)
    Dg = torch.cat([outs[0][0], outs[1][0]], 1).contiguous()
    Ig = torch.cat([outs[0][1], outs[1][1]], 1).contiguous()
    Sg = torch.cat([outs[0][2], outs[1][2]], 1).contiguous()
    Jg = torch.cat([outs[0][3], outs[1][3]], 1).contiguous()
    L = _lib.lib()
    st = _lib.current_stream_ptr(0)
    Dm, Im = torch.empty((nq, 50), device="cuda"), torch.empty((nq, 50), dtype=torch.int64, device="cuda")
    Sm, Jm = torch.empty((nq, 50), device="cuda"), torch.empty((nq, 50), dtype=torch.int64, device="cuda")
    _lib.check(L.hr_merge_topk(Dg.data_ptr(), Ig.data_ptr(), nq, 100, 50, 1, -3.4e38, Dm.data_ptr(), Im.data_ptr(), 0, st))
    _lib.check(L.hr_merge_topk(Sg.data_ptr(), Jg.data_ptr(), nq, 100, 50, 1, 0.0, Sm.data_ptr(), Jm.data_ptr(), 0, st))
    oS, oI = torch.empty((nq, 10), device="cuda"), torch.empty((nq, 10), dtype=torch.int64, device="cuda")
    _lib.check(L.hr_fuse(Dm.data_ptr(), Im.data_ptr(), Sm.data_ptr(), Jm.data_ptr(), None, nq, 50, 10, 0, 0, 0.7, 0.3,
                         oS.data_ptr(), oI.data_ptr(), 0, st))
    torch.cuda.synchronize()
    assert np.array_equal(oI.cpu().numpy(), I)
    np.testing.assert_allclose(oS.cpu().numpy(), S, rtol=2e-6)
    # merged dense candidates equal the unsharded dense search exactly
    Df, If = full_ix.search(q, 50)
    assert np.array_equal(Im.cpu().numpy(), If) and np.array_equal(Dm.cpu().numpy(), Df)
    # merge + fusion straight from the gathered per-rank blocks (the N>1 path): same answer
    gathered = torch.cat(blocks)
    nn = nq * 50
    Dv, Sv, Iv, Jv = ShardedRetriever._views(gathered, nn)
    gS, gI = torch.empty((nq, 10), device="cuda"), torch.empty((nq, 10), dtype=torch.int64, device="cuda")
    _lib.check(L.hr_merge_fuse_lists(full_ix._h, Dv.data_ptr(), Iv.data_ptr(), Sv.data_ptr(), Jv.data_ptr(), 2, 24 * nn,
                                     nq, 50, 10, 0, 0.7, 0.3, gS.data_ptr(), gI.data_ptr(), st))
    torch.cuda.synchronize()
    assert torch.equal(gI, oI) and torch.equal(gS, oS)
    # single-rank ShardedRetriever (no process group) is the same call path the N>1 bench uses
    S1, I1 = ShardedRetriever(full_ix, full_bm).retrieve(torch.from_numpy(q).cuda(), qs, 10)
    assert np.array_equal(I1.cpu().numpy(), I)


def test_merge_topk_vs_oracle(gpu):
    import torch
    rng = np.random.default_rng(1)
    nq, L, k = 33, 96, 20
    S = rng.standard_normal((nq, L)).astype(np.float32)
    S[:, 10] = S[:, 3]                                   # exact ties
    I = np.stack([rng.permutation(10000)[:L] for _ in range(nq)]).astype(np.int64)
    I[:, 90:] = -1
    for largest in (1, 0):
        Sr, Ir = fusion.merge_shards(S, I, k, largest=bool(largest), pad_score=7.0)
        Sd, Id = torch.from_numpy(S).cuda(), torch.from_numpy(I).cuda()
        oS, oI = torch.empty((nq, k), device="cuda"), torch.empty((nq, k), dtype=torch.int64, device="cuda")
        _lib.check(_lib.lib().hr_merge_topk(Sd.data_ptr(), Id.data_ptr(), nq, L, k, largest, 7.0, oS.data_ptr(),
                                            oI.data_ptr(), 0, _lib.current_stream_ptr(0)))
        torch.cuda.synchronize()
        assert np.array_equal(oI.cpu().numpy(), Ir) and np.array_equal(oS.cpu().numpy(), Sr)
