#!/usr/bin/env python
"""BM25 kernel variants at bench scale, one corpus build (run on the GPU box):

    python tools/bm25_tune.py [rows] [vocab] [nq] [k] [variants]

variants: comma list of wide:spans:batch (e.g. 0:0:3,1:0:3,0:32:3).  Every variant's (S, I) is compared bit for bit
with the first one: every configuration sums a doc's terms in plan order, so the answers must be identical.  (Round 2
used this tool to compare the flat-sweep kernel with the round-1 slot kernel, since removed: bit-identical at C1.)"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import _lib, bm25 as pbm25, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
vocab = int(sys.argv[2]) if len(sys.argv) > 2 else 30_000
nq = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
k = int(sys.argv[4]) if len(sys.argv) > 4 else 50
variants = sys.argv[5] if len(sys.argv) > 5 else "0:0:3,0:0:2,1:0:3"
dev = torch.device("cuda", 0)
t0 = time.time()
indptr, post_doc, post_tf, doc_len = synth.sparse_corpus_csr_torch(rows, vocab, dev)
bm = pbm25.BM25Index.from_csr(indptr, post_doc, post_tf, doc_len, vocab, device=0)
del indptr, post_doc, post_tf, doc_len
torch.cuda.empty_cache()
qi, qt = synth.sparse_queries_csr(nq, vocab)
qi, qt = torch.from_numpy(qi).to(dev), torch.from_numpy(qt).to(dev)
print(f"setup {time.time() - t0:.1f}s nnz {bm.nnz}", flush=True)
ref = None
for v in variants.split(","):
    wide, spans, batch = (int(x) for x in (v.split(":") + ["0", "3"])[:3])
    _lib.set_option("bm25_batch", batch)
    _lib.set_option("bm25_wide", wide)
    _lib.set_option("bm25_spans", spans)
    for _ in range(3):
        S, I, touched = bm.search((qi, qt), k, return_postings=True)
    ms = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        S, I = bm.search((qi, qt), k)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    m = sorted(ms)[len(ms) // 2]
    same = ""
    if ref is None:
        ref = (S.clone(), I.clone())
    else:
        same = f" equal_to_first: S {bool(torch.equal(S, ref[0]))} I {bool(torch.equal(I, ref[1]))}"
        if not torch.equal(I, ref[1]):
            bad = (I != ref[1]).any(dim=1).nonzero().flatten()[:5].tolist()
            same += f" first differing queries {bad}"
    print(f"variant wide={wide} spans={spans} batch={batch}: {m:.3f} ms (min {min(ms):.3f}) postings {touched} "
          f"{touched * 8 / m / 1e6:.1f} GB/s algorithmic{same}", flush=True)
