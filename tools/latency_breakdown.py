#!/usr/bin/env python
"""Latency mode (C4) breakdown on one GPU: batch-1 dense search, BM25 search and hybrid retrieve(), device-resident
inputs (CUDA events) and host in / host out (perf_counter).  Run on the GPU box."""
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

dev = torch.device("cuda", 0)
n, d, V = 10_000_000, 1024, 30_000
ix = hf.IndexFlatIP(d, storage=os.getenv("C4_STORAGE", "f32+bf16"))
planted = synth.dense_corpus_into(ix, n, d, dev, keep_rows=4096)
indptr, post_doc, post_tf, doc_len = synth.sparse_corpus_csr_torch(n, V, dev)
bm = pbm25.BM25Index.from_csr(indptr, post_doc, post_tf, doc_len, V, device=0)
del indptr, post_doc, post_tf, doc_len
torch.cuda.empty_cache()
eng = HybridRetriever(ix, bm)
qs = synth.sparse_queries_np(200, V)
q_all = synth.dense_queries_torch(planted, 200, d, dev)
q_host = q_all.cpu().numpy()


def ev(fn, reps=200):
    for i in range(10):
        fn(i)
    ms = []
    for i in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(i % 200)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.percentile(ms, 50)), float(np.percentile(ms, 99))


def wall(fn, reps=200):
    for i in range(10):
        fn(i)
    ms = []
    for i in range(reps):
        t0 = time.perf_counter()
        fn(i % 200)
        ms.append((time.perf_counter() - t0) * 1e3)
    return float(np.percentile(ms, 50)), float(np.percentile(ms, 99))


csr = [pbm25.query_csr([q]) for q in qs]
csr_dev = [(torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)) for a, b in csr]
print("dense search  k=50, device io  p50/p99 ms:", ev(lambda i: ix.search(q_all[i:i + 1], 50)))
print("bm25  search  k=50, device io  p50/p99 ms:", ev(lambda i: bm.search(csr_dev[i], 50)))
print("hybrid retrieve top-10, device io       :", ev(lambda i: eng.retrieve(q_all[i:i + 1], csr_dev[i], 10)))
print("hybrid retrieve top-10, host in/out wall:", wall(lambda i: eng.retrieve(q_host[i:i + 1], [qs[i]], 10)))
print("dense search  k=50, host in/out wall    :", wall(lambda i: ix.search(q_host[i:i + 1], 50)))
