#!/usr/bin/env python
"""Latency pieces on one GPU (run on the GPU box): BM25 search, dense search and hybrid retrieve with CUDA events
for growing batches, then the batch-1 hybrid retrieve() host in / host out by wall clock.
python tools/c4_check.py [rows]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import bm25 as pbm25, faiss as hf, synth  # noqa: E402
from intool_rag_b200.retriever import HybridRetriever  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
dev = torch.device("cuda", 0)
d, V = 1024, 30_000
ix = hf.IndexFlatIP(d, storage="f32+bf16")
planted = synth.dense_corpus_into(ix, rows, d, dev, keep_rows=4096)
indptr, post_doc, post_tf, doc_len = synth.sparse_corpus_csr_torch(rows, V, dev)
bm = pbm25.BM25Index.from_csr(indptr, post_doc, post_tf, doc_len, V, device=0)
del indptr, post_doc, post_tf, doc_len
torch.cuda.empty_cache()
eng = HybridRetriever(ix, bm)
qs = synth.sparse_queries_np(1024, V)
q_all = synth.dense_queries_torch(planted, 1024, d, dev)
q_host = q_all.cpu().numpy()


def ev(fn, reps=100):
    for i in range(5):
        fn(i)
    ms = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.percentile(ms, 50)), float(np.percentile(ms, 99))


for nq in (1, 2, 8, 32, 128, 256, 512, 1024):
    csr, qd = [], []
    for i in range(8):
        o = (i * nq) % 1024 if nq < 1024 else 0
        a, b = pbm25.query_csr(qs[o:o + nq])
        csr.append((torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)))
        qd.append(q_all[o:o + nq].contiguous())
    b50 = ev(lambda i: bm.search(csr[i % 8], 50))
    d50 = ev(lambda i: ix.search(qd[i % 8], 50))
    scan = ix.stats()["scan_ms"]
    h = ev(lambda i: eng.retrieve(qd[i % 8], csr[i % 8], 10))
    print(f"nq={nq:5d}  bm25 p50/p99 {b50[0]:7.3f}/{b50[1]:7.3f} ms   dense {d50[0]:7.3f}/{d50[1]:7.3f} (scan kernel {scan:6.3f})   "
          f"hybrid {h[0]:7.3f}/{h[1]:7.3f}", flush=True)
t = []
for i in range(300):
    t0 = time.perf_counter()
    eng.retrieve(q_host[i:i + 1], [qs[i]], 10)
    t.append((time.perf_counter() - t0) * 1e3)
print(f"batch-1 hybrid retrieve host in/out wall: p50 {np.percentile(t[20:], 50):.3f} p99 {np.percentile(t[20:], 99):.3f} ms")
