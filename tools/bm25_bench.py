#!/usr/bin/env python
"""BM25 kernel timing at bench scale (run on the GPU box): python tools/bm25_bench.py [rows] [vocab] [nq] [k]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import bm25 as pbm25, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 30_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
k = int(sys.argv[4]) if len(sys.argv) > 4 else 50
dev = torch.device("cuda", 0)
t0 = time.time()
indptr, post_doc, post_tf, doc_len = synth.sparse_corpus_csr_torch(rows, vocab, dev)
bm = pbm25.BM25Index.from_csr(indptr, post_doc, post_tf, doc_len, vocab, device=0)
del indptr, post_doc, post_tf, doc_len
torch.cuda.empty_cache()
qi, qt = synth.sparse_queries_csr(nq, vocab)
qi, qt = torch.from_numpy(qi).to(dev), torch.from_numpy(qt).to(dev)
print(f"setup {time.time() - t0:.1f}s nnz {bm.nnz}", flush=True)
for _ in range(3):
    S, I, touched = bm.search((qi, qt), k, return_postings=True)
ms = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    bm.search((qi, qt), k)
    e1.record()
    torch.cuda.synchronize()
    ms.append(e0.elapsed_time(e1))
m = sorted(ms)[len(ms) // 2]
print(f"bm25 search rows={rows} V={vocab} nq={nq} k={k}: {m:.3f} ms (all {['%.2f' % x for x in ms]}), "
      f"postings {touched}, {touched * 8 / m / 1e6:.1f} GB/s algorithmic", flush=True)
print("checksum", float(S.sum()), int(I.sum()))
