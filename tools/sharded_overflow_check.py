#!/usr/bin/env python
"""Two ranks, one of them with more certificate failures than the device-driven fallback handles (64 per search):
exercises the trailer / flag / repeat-the-exchange path of hr_retrieve_sharded, which the benchmarks never take.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29530 tools/sharded_overflow_check.py
Rank 0 holds random unit rows, rank 1 holds 5000 IDENTICAL rows (massive ties: every query's certificate fails there);
150 queries equal to that row.  The sharded answer must equal the answer with both ranks in exhaustive exact mode."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import faiss as hf, synth  # noqa: E402
from intool_rag_b200.sharded import ShardedRetriever  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
d, n, nq, k = 32, 5000, 150, 10
row = (np.ones((1, d), np.float32) / np.sqrt(d)).astype(np.float32)
x = synth.dense_corpus_np(n, d, seed=99) if rank == 0 else np.repeat(row, n, axis=0)
ix = hf.IndexFlatIP(d, device=local)
ix.add(x)
ix.set_id_base(rank * n)
q = torch.from_numpy(np.repeat(row, nq, axis=0)).cuda()
eng = ShardedRetriever(ix, None)
S, I = eng.retrieve(q, None, k)
st = ix.stats()
ix.set_mode("exact")
Se, Ie = eng.retrieve(q, None, k)
ix.set_mode("auto")
Sh, Ih = eng.retrieve(q.cpu().numpy(), None, k)          # host in / host out takes the same path
ok = bool(torch.equal(I, Ie) and torch.equal(S, Se) and np.array_equal(Ih, I.cpu().numpy()))
want = torch.arange(n, n + k, device=I.device).expand(nq, k)   # the identical rows of rank 1, ids ascending
ok = ok and bool(torch.equal(I, want))
flags = torch.tensor([st["flagged"], int(ok)], device=I.device)
allf = [torch.zeros_like(flags) for _ in range(world)]
dist.all_gather(allf, flags)
if rank == 0:
    print("flagged per rank:", [int(f[0]) for f in allf], " answers equal exact mode on every rank:", all(int(f[1]) for f in allf))
    assert all(int(f[1]) for f in allf) and int(allf[1][0]) > 64 and int(allf[0][0]) <= 64
    print("sharded overflow check OK")
dist.barrier()
dist.destroy_process_group()
