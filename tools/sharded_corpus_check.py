#!/usr/bin/env python
"""File-backed corpus, row-sharded over real ranks (SURVEY.md 8f rank 2 + 8e): every rank loads only its row range of
the per-document index files, BM25 lists are merged under the corpus-wide statistics, hr_retrieve_sharded answers equal
the single-process answer over the whole corpus.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/sharded_corpus_check.py"""
import asyncio
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import corpus as pcorpus, storage  # noqa: E402
from intool_rag_b200.retriever import HybridRetriever  # noqa: E402
from intool_rag_b200.sharded import ShardedRetriever  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
os.environ["HR_DEVICE"] = str(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
# every rank writes the same documents into its own temp dir (same seed): stands for a shared STORAGE_DIR
sd = tempfile.mkdtemp(prefix=f"corpus_r{rank}_")
os.environ["STORAGE_DIR"] = sd
rng = np.random.default_rng(7)
d, words = 64, [f"w{i}" for i in range(300)]
xs, texts_all = [], []
for di in range(7):
    n = int(rng.integers(150, 600))
    x = rng.standard_normal((n, d)).astype(np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    doc_id = f"doc{di:02d}"
    storage.save_faiss_index(storage.create_faiss_index(x), os.path.join(sd, f"{doc_id}_faiss.index"))
    texts = [" ".join(rng.choice(words, size=int(rng.integers(5, 40)))) for _ in range(n)]
    chunks = [{"chunk_id": f"{doc_id}:{i}", "page": i // 3, "text": texts[i], "chunk_index": i} for i in range(n)]
    json.dump({"total": n, "chunks": chunks}, open(os.path.join(sd, f"{doc_id}_chunks.json"), "w"))
    storage.build_bm25_sidecar(doc_id, texts, storage_dir=sd)
    xs.append(x)
    texts_all += texts
X = np.concatenate(xs)
shard = pcorpus.Corpus(sd, rank=rank, world=world, device=local)
assert shard.index.ntotal == shard.hi - shard.lo and shard.ntotal_global == len(X)
eng = ShardedRetriever(shard.index, shard.bm25)
nq = 40
qrows = rng.integers(0, len(X), size=nq)
q = X[qrows] + 0.05 * rng.standard_normal((nq, d)).astype(np.float32)
q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
toks = [shard.vocab.encode(" ".join(texts_all[r].split()[:4]) + " w3 w17") for r in qrows]
S, I = eng.retrieve(q, toks, 10)
ok = True
if rank == 0:
    full = pcorpus.Corpus(sd, device=local)                       # the whole corpus on one GPU
    S1, I1 = HybridRetriever(full.index, full.bm25).retrieve(q, toks, 10)
    ok = bool(np.array_equal(I, I1) and np.array_equal(S, S1))
    hit = full.hit_dict(int(I[0, 0]), float(S[0, 0]))
    print("rows per rank:", shard.hi - shard.lo, "of", len(X), "| sharded == single:", ok, "| first hit:", hit["chunk_id"], hit["page"])
flag = torch.tensor([int(ok)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
assert int(flag.item()) == 1
if rank == 0:
    print("sharded corpus check OK")
dist.barrier()
dist.destroy_process_group()
