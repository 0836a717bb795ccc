#!/usr/bin/env python
"""Randomised parity soak (run on the GPU box): dense auto path vs the exhaustive exact scan (bit-exact), BM25 vs the
oracle (1e-5 relative), hybrid retrieve vs the oracle, over random shapes, metrics, storages, duplicates, ragged and
empty queries.  python tools/fuzz_parity.py [seconds] [seed]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import bm25 as pbm25, faiss as hf, synth  # noqa: E402
from intool_rag_b200.retriever import HybridRetriever  # noqa: E402
from oracle import bm25 as obm25, flat, hybrid  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
t_end = time.time() + budget
n_dense = n_bm = n_hy = 0
stats = {"flagged": 0, "deeper": 0}
while time.time() < t_end:
    # ---------------- dense: auto == exact, bit for bit ----------------
    n = int(rng.choice([1, 7, 255, 257, 1000, 5000, 40000, 150000]))
    d = int(rng.choice([1, 2, 31, 32, 64, 100, 256, 384]))
    nq = int(rng.choice([1, 2, 33, 128, 129, 300]))
    k = int(rng.choice([1, 5, 10, 50, 64, 100, 128]))
    metric = str(rng.choice(["ip", "l2"]))
    storage = str(rng.choice(["f32", "f32+bf16", "bf16"]))
    x = rng.standard_normal((n, d)).astype(np.float32)
    if rng.random() < 0.7:
        x /= np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-6)
    if n > 50 and rng.random() < 0.5:      # duplicates and near-duplicates
        src = int(rng.integers(0, n))
        m = int(rng.integers(2, min(n, 400)))
        where = rng.choice(n, size=m, replace=False)
        x[where] = x[src] + (0 if rng.random() < 0.5 else 1e-4) * rng.standard_normal((m, d)).astype(np.float32)
    q = rng.standard_normal((nq, d)).astype(np.float32)
    q[: nq // 2] = x[rng.integers(0, n, size=nq // 2)] + 0.05 * q[: nq // 2]
    ix = (hf.IndexFlatIP if metric == "ip" else hf.IndexFlatL2)(d, storage=storage)
    half = n // 2
    ix.add(x[:half])
    ix.add(x[half:])
    D, I = ix.search(q, k)
    st = ix.stats()
    stats["flagged"] += st["flagged"]
    stats["deeper"] += st["deeper"]
    ix.set_mode("exact")
    De, Ie = ix.search(q, k)
    assert np.array_equal(I, Ie) and np.array_equal(D, De), ("dense", n, d, nq, k, metric, storage)
    n_dense += 1
    # ---------------- BM25 vs oracle ----------------
    nd = int(rng.choice([1, 50, 3000, 20000, 70000]))
    V = int(rng.choice([5, 200, 3000]))
    t, dd, dl = synth.sparse_corpus_np(nd, V, seed=int(rng.integers(1 << 30)), mean_len=float(rng.choice([8.0, 40.0])))
    nqs = int(rng.choice([1, 3, 40, 700]))    # 700: the running-list (lock) mode of the sweep kernel
    qs = synth.sparse_queries_np(nqs, V, seed=int(rng.integers(1 << 30)), stop=min(4, V - 1))
    if nqs > 2:
        qs[0] = []
        qs[1] = qs[1] + qs[1] + [V + 3, -1]
    kk = int(rng.choice([1, 10, 50, 128]))
    bm = pbm25.BM25Index.from_tokens(t.astype(np.int32), dd.astype(np.int32), nd, V)
    oc = obm25.BM25Corpus.from_token_matrix(t, dd, dl, V)
    S, J = bm.search(qs, kk)
    Sr, Jr = oc.search(qs, kk)
    np.testing.assert_allclose(S, Sr, rtol=1e-5, atol=1e-7, err_msg=str(("bm25", nd, V, nqs, kk)))
    assert ((J >= 0) == (Jr >= 0)).all()
    n_bm += 1
    # ---------------- hybrid vs oracle (small) ----------------
    if nd >= 50 and rng.random() < 0.5:
        dh = 32
        xh = synth.dense_corpus_np(nd, dh, seed=int(rng.integers(1 << 30)))
        qh = synth.dense_queries_np(xh, nqs, seed=int(rng.integers(1 << 30)))
        ixh = hf.IndexFlatIP(dh, storage=storage)
        ixh.add(xh)
        oi = flat.IndexFlatIP(dh)
        oi.add(xh if storage != "bf16" else __import__("torch").from_numpy(xh).to(__import__("torch").bfloat16).float().numpy())
        Sh, Ih = HybridRetriever(ixh, bm).retrieve(qh, qs, 10)
        Sor, Ior, _ = hybrid.retrieve(oi, oc, qh, qs, 10, precision="f64")
        np.testing.assert_allclose(Sh, Sor, rtol=3e-5, atol=2e-6, err_msg=str(("hybrid", nd, V, nqs)))
        n_hy += 1
print(f"fuzz ok: {n_dense} dense cases (auto == exact bit for bit; {stats}), {n_bm} BM25 cases, {n_hy} hybrid cases", flush=True)
