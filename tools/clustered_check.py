#!/usr/bin/env python
"""Certificate behaviour on CLUSTERED embeddings (real corpora are not uniform on the sphere): how many queries
the filter certifies vs sends to the exact scan, per storage mode.  Run on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import faiss as hf  # noqa: E402

dev = torch.device("cuda", 0)
n, d, nq, ncl = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 1024, 256, 2000
g = torch.Generator(device=dev)
g.manual_seed(7)
centers = torch.nn.functional.normalize(torch.randn((ncl, d), generator=g, device=dev), dim=1)
for noise in (0.5, 1.0, 2.0):
    for storage in ("f32", "f32+bf16", "bf16"):
        ix = hf.IndexFlatIP(d, storage=storage)
        gg = torch.Generator(device=dev)
        gg.manual_seed(11)
        for r0 in range(0, n, 1 << 18):
            nr = min(1 << 18, n - r0)
            cid = torch.randint(0, ncl, (nr,), generator=gg, device=dev)
            x = centers[cid] + noise / 32 * torch.randn((nr, d), generator=gg, device=dev)
            ix.add(torch.nn.functional.normalize(x, dim=1))
        cid = torch.randint(0, ncl, (nq,), generator=gg, device=dev)
        q = torch.nn.functional.normalize(centers[cid] + noise / 32 * torch.randn((nq, d), generator=gg, device=dev), dim=1)
        res = {}
        for k in (10, 50):
            D, I = ix.search(q, k)
            st = ix.stats()
            res[k] = (int(st["flagged"]), round(float(st["total_ms"]), 2), round(float((D[:, 0] - D[:, k - 1]).median()), 4))
        print(f"noise={noise} storage={storage}: k -> (flagged of {nq}, total ms, median top1-topk gap) {res}", flush=True)
        del ix
        torch.cuda.empty_cache()
