"""CPU, world_size 2 over gloo: the multi-rank host logic (candidate all-gather layout, global BM25
statistics, shard bounds).  Per-rank compute is injected from the oracle so the plumbing is tested
without a GPU; the merged answer must equal the single-shard answer exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import intool_rag_b200  # noqa: F401
    from intool_rag_b200 import sharded, synth
    from oracle import flat, bm25, fusion
    n, d, V, nq, kc = 900, 32, 120, 11, 20
    x = synth.dense_corpus_np(n, d)
    x[500] = x[40]
    q = synth.dense_queries_np(x, nq)
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=24.0)
    qt = synth.sparse_queries_np(nq, V, stop=4)
    lo, hi = sharded.shard_bounds(n, world, rank)
    # local dense shard -> global ids
    ix = flat.IndexFlatIP(d)
    ix.add(x[lo:hi])
    D, I = ix.search(q, kc)
    I = np.where(I >= 0, I + lo, -1)
    # local sparse shard with GLOBAL statistics
    m = (dd >= lo) & (dd < hi)
    local = bm25.BM25Corpus.from_token_matrix(t[m], dd[m] - lo, dl[lo:hi], V)
    df, N, avgdl = sharded.global_bm25_stats(torch.from_numpy(local.df.copy()), hi - lo, int(dl[lo:hi].sum()))
    local.idf = bm25.idf_table(df.numpy(), N)
    dlf = local.doc_len[local.post_doc].astype(np.float64)
    tf = local.post_tf.astype(np.float64)
    local.impact = tf * (local.k1 + 1) / (tf + local.k1 * (1 - local.b + local.b * dlf / avgdl))
    S, J = local.search(qt, kc)
    J = np.where(J >= 0, J + lo, -1)
    Dg, Ig = sharded.gather_candidates(torch.from_numpy(D), torch.from_numpy(I))
    Sg, Jg = sharded.gather_candidates(torch.from_numpy(S), torch.from_numpy(J))
    assert Dg.shape == (nq, world * kc)
    # the packed single all-gather the GPU path uses (D | S | I | J per rank, 24 bytes per candidate):
    # rank r's block must hold exactly rank r's four lists
    nn = nq * kc
    loc = torch.empty(24 * nn, dtype=torch.uint8)
    for view, arr in zip(sharded.ShardedRetriever._views(loc, nn), (D, S, I, J)):
        view.copy_(torch.from_numpy(np.ascontiguousarray(arr)).reshape(-1))
    gathered = torch.empty(world * 24 * nn, dtype=torch.uint8)
    dist.all_gather_into_tensor(gathered, loc)
    for r in range(world):
        Dr, Sr, Ir, Jr = sharded.ShardedRetriever._views(gathered[r * 24 * nn:(r + 1) * 24 * nn], nn)
        assert torch.equal(Dr.view(nq, kc), Dg[:, r * kc:(r + 1) * kc]) and torch.equal(Ir.view(nq, kc), Ig[:, r * kc:(r + 1) * kc])
        assert torch.equal(Sr.view(nq, kc), Sg[:, r * kc:(r + 1) * kc]) and torch.equal(Jr.view(nq, kc), Jg[:, r * kc:(r + 1) * kc])
    Dm, Im = fusion.merge_shards(Dg.numpy(), Ig.numpy(), kc, largest=True, pad_score=-flat.FLT_MAX)
    Sm, Jm = fusion.merge_shards(Sg.numpy(), Jg.numpy(), kc, largest=True, pad_score=0.0)
    fs, fi = fusion.fuse(Dm, Im, Sm, Jm, 10)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), Dm=Dm, Im=Im, Sm=Sm, Jm=Jm, fs=fs, fi=fi,
             N=N, avgdl=avgdl, df=df.numpy())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gather_merge_equals_single_shard(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    import intool_rag_b200  # noqa: F401
    from intool_rag_b200 import synth
    from oracle import flat, bm25, fusion
    n, d, V, nq, kc = 900, 32, 120, 11, 20
    x = synth.dense_corpus_np(n, d)
    x[500] = x[40]
    q = synth.dense_queries_np(x, nq)
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=24.0)
    qt = synth.sparse_queries_np(nq, V, stop=4)
    ix = flat.IndexFlatIP(d)
    ix.add(x)
    D, I = ix.search(q, kc)
    c = bm25.BM25Corpus.from_token_matrix(t, dd, dl, V)
    S, J = c.search(qt, kc)
    fs, fi = fusion.fuse(D, I, S, J, 10)
    r0 = np.load(tmp_path / "r0.npz")
    r1 = np.load(tmp_path / "r1.npz")
    for key in ("Dm", "Im", "Sm", "Jm", "fs", "fi"):
        assert np.array_equal(r0[key], r1[key]), f"ranks disagree on {key}"
    assert int(r0["N"]) == n and float(r0["avgdl"]) == pytest.approx(c.avgdl, rel=1e-15)
    assert np.array_equal(r0["df"], c.df)
    assert np.array_equal(r0["Im"], I) and np.array_equal(r0["Dm"], D)
    assert np.array_equal(r0["Jm"], J)
    np.testing.assert_allclose(r0["Sm"], S, rtol=2e-7)
    assert np.array_equal(r0["fi"], fi)
    np.testing.assert_allclose(r0["fs"], fs, rtol=1e-6)
