#!/usr/bin/env python
"""Why does the BM25 search take ~2 ms longer inside the hybrid step than back to back?  (run on the GPU box)
Times the BM25 search (CUDA events around it alone) after: another BM25 search, an L2 flush, the dense search,
the dense search + an idle gap.  python tools/after_scan.py [rows] [nq]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import bm25 as pbm25, faiss as hf, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
nq = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
dev = torch.device("cuda", 0)
d, V = 1024, 30_000
ix = hf.IndexFlatIP(d, storage="f32+bf16")
planted = synth.dense_corpus_into(ix, rows, d, dev, keep_rows=4096)
indptr, post_doc, post_tf, doc_len = synth.sparse_corpus_csr_torch(rows, V, dev)
bm = pbm25.BM25Index.from_csr(indptr, post_doc, post_tf, doc_len, V, device=0)
del indptr, post_doc, post_tf, doc_len
torch.cuda.empty_cache()
qs = synth.sparse_queries_np(nq, V)
a, b = pbm25.query_csr(qs)
csr = (torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev))
q = synth.dense_queries_torch(planted, nq, d, dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)


def timed_bm25(before, reps=20):
    ms = []
    for _ in range(reps):
        before()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        bm.search(csr, 50)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms)), float(np.min(ms)), float(np.max(ms))


def dense():
    ix.search(q, 50)


def dense_then_idle(ms):
    def f():
        ix.search(q, 50)
        torch.cuda.synchronize()
        time.sleep(ms / 1e3)
    return f


for _ in range(3):
    bm.search(csr, 50)
    dense()
torch.cuda.synchronize()
print("after another BM25 search      : %.3f ms (min %.3f max %.3f)" % timed_bm25(lambda: bm.search(csr, 50)))
print("after an L2 flush (512 MB fill): %.3f ms (min %.3f max %.3f)" % timed_bm25(lambda: flush.fill_(1)))
print("after the dense search         : %.3f ms (min %.3f max %.3f)" % timed_bm25(dense))
for gap in (1, 3, 10, 30):
    print("after dense + %2d ms idle       : %.3f ms (min %.3f max %.3f)" % ((gap,) + timed_bm25(dense_then_idle(gap))))
print("after dense search x3          : %.3f ms (min %.3f max %.3f)" % timed_bm25(lambda: (dense(), dense(), dense())))
