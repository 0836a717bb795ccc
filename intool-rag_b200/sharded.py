"""Row-sharded multi-GPU retrieval: one process per GPU, torch.distributed for the plumbing.

The corpus is split by row (document) across ranks; every rank scans its shard for the whole
query batch and emits its local top-k_c per modality with GLOBAL ids (row + shard base).  One
all-gather carries only those (world x k_c) candidates; every rank then merges and fuses, so all
ranks hold the same answer.  BM25 stays exact because idf / avgdl use corpus-wide statistics
(all-reduced at build time).  The reference is single-process; this is the B200
design of §8(e).
"""
from __future__ import annotations

import numpy as np

from . import _lib


def shard_bounds(n_rows: int, world: int, rank: int):
    """Contiguous row range [lo, hi) of `rank`; the first n_rows % world ranks get one extra row."""
    base, rem = divmod(int(n_rows), int(world))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def global_bm25_stats(df_local, n_docs_local: int, total_len_local: int, group=None):
    """All-reduce the corpus statistics BM25 needs: (df int64[V], N, avgdl).  Works under gloo (CPU
    tensors) and nccl (CUDA tensors)."""
    import torch
    import torch.distributed as dist
    df = df_local.clone() if hasattr(df_local, "clone") else torch.from_numpy(np.asarray(df_local, np.int64)).clone()
    scal = torch.tensor([int(n_docs_local), int(total_len_local)], dtype=torch.int64, device=df.device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(df, group=group)
        dist.all_reduce(scal, group=group)
    n, tot = int(scal[0].item()), int(scal[1].item())
    return df, n, (tot / n if n else 0.0)


def gather_candidates(scores, ids, group=None):
    """scores/ids: [nq, kc] of this rank -> ([nq, world*kc], [nq, world*kc]) in rank order."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return scores, ids
    world = dist.get_world_size(group)
    nq, kc = scores.shape
    s_all = torch.empty((world * nq, kc), dtype=scores.dtype, device=scores.device)
    i_all = torch.empty((world * nq, kc), dtype=ids.dtype, device=ids.device)
    dist.all_gather_into_tensor(s_all, scores.contiguous(), group=group)
    dist.all_gather_into_tensor(i_all, ids.contiguous(), group=group)
    s_all = s_all.view(world, nq, kc)
    i_all = i_all.view(world, nq, kc)
    return (s_all.permute(1, 0, 2).reshape(nq, world * kc).contiguous(),
            i_all.permute(1, 0, 2).reshape(nq, world * kc).contiguous())


class ShardedRetriever:
    """Hybrid retrieval over row shards.  `index` / `bm25` hold THIS rank's shard with id_base set
    to the shard's first global row and BM25 built with global statistics.

    One step = hr_candidates (local BM25 + dense top-k_c into one packed buffer: D | S | I | J),
    ONE all-gather of that buffer (24 * nq * k_c bytes per rank), hr_merge_fuse_lists reading the
    gathered per-rank blocks in place (rank order == id order, comparator (score, id))."""

    def __init__(self, index, bm25=None, vector_weight: float = 0.7, bm25_weight: float = 0.3,
                 fusion: str = "weighted", group=None):
        self.index, self.bm25, self.group = index, bm25, group
        self.vector_weight, self.bm25_weight, self.fusion = vector_weight, bm25_weight, fusion
        self._bufs = {}

    @staticmethod
    def _views(buf, n):
        """(D float32[n], S float32[n], I int64[n], J int64[n]) views of one packed 24n-byte block."""
        import torch
        f = buf[:8 * n].view(torch.float32)
        i = buf[8 * n:24 * n].view(torch.int64)
        return f[:n], f[n:2 * n], i[:n], i[n:2 * n]

    def retrieve(self, query_embeddings, query_tokens=None, top_k: int = 10, k_c: int | None = None):
        """query_embeddings: torch CUDA float32 [nq, d] (replicated on every rank).  Returns torch CUDA
        (scores [nq, top_k], ids [nq, top_k]) — identical on every rank."""
        import torch
        import torch.distributed as dist
        from .retriever import candidate_depth, _MODES
        kc = candidate_depth(top_k) if k_c is None else int(k_c)
        q = query_embeddings.to(torch.float32).contiguous()
        dev = q.device
        nq = q.shape[0]
        if q.dim() != 2 or q.shape[1] != self.index.d:
            raise AssertionError(f"retrieve: expected [nq, {self.index.d}] embeddings, got {tuple(q.shape)}")
        world = dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1
        n = nq * kc
        key = (nq, kc, world)
        if key not in self._bufs:
            local = torch.empty(24 * n, dtype=torch.uint8, device=dev)
            gathered = torch.empty(world * 24 * n, dtype=torch.uint8, device=dev) if world > 1 else local
            # raw pointers of the four lists of the local block and of rank 0's block of the gathered buffer
            self._bufs = {key: (local, gathered, [v.data_ptr() for v in self._views(local, n)],
                                [v.data_ptr() for v in self._views(gathered, n)])}
        local, gathered, lp, gp = self._bufs[key]
        use_bm = self.bm25 is not None and query_tokens is not None
        qi = qt = None
        if use_bm:
            from .bm25 import query_csr
            ip, tm = query_csr(query_tokens)
            if not hasattr(ip, "is_cuda"):
                ip = torch.from_numpy(np.ascontiguousarray(ip)).to(dev)
                tm = torch.from_numpy(np.ascontiguousarray(tm)).to(dev)
            qi, qt = ip.to(torch.int32).contiguous(), tm.to(torch.int32).contiguous()
            if qi.shape[0] != nq + 1:
                raise AssertionError("retrieve: query_tokens and query_embeddings disagree on nq")
        L = _lib.lib()
        st = _lib.current_stream_ptr(self.index.device)
        # outputs first: hr_candidates returns after a stream synchronisation, nothing but the all-gather and the
        # merge launch should stand between that point and the GPU's next kernel
        oS = torch.empty((nq, top_k), dtype=torch.float32, device=dev)
        oI = torch.empty((nq, top_k), dtype=torch.int64, device=dev)
        _lib.check(L.hr_candidates(self.index._h, self.bm25._h if use_bm else None, q.data_ptr(),
                                   qi.data_ptr() if use_bm else None, qt.data_ptr() if use_bm else None, nq,
                                   int(qt.numel()) if use_bm else 0, kc, lp[0], lp[2], lp[1], lp[3], st))
        if world > 1:
            dist.all_gather_into_tensor(gathered, local, group=self.group)
        _lib.check(L.hr_merge_fuse_lists(self.index._h, gp[0], gp[2], gp[1], gp[3], world, 24 * n, nq, kc, top_k,
                                         _MODES[self.fusion], self.vector_weight, self.bm25_weight, oS.data_ptr(),
                                         oI.data_ptr(), st))
        return oS, oI
