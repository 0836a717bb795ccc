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


def comm_unique_id() -> bytes:
    """128-byte NCCL unique id (hr_comm_unique_id): rank 0 creates it, every other rank needs a copy."""
    import ctypes as C
    buf = C.create_string_buffer(128)
    _lib.check(_lib.lib().hr_comm_unique_id(buf))
    return buf.raw


class Comm:
    """One rank of a row-sharded corpus: owns the library's NCCL communicator (include/hr_b200.h: hr_comm_*).
    `unique_id` may be None when world == 1 (no NCCL involved)."""

    def __init__(self, rank: int, world: int, device: int, unique_id: bytes | None = None):
        import ctypes as C
        self.rank, self.world, self.device = int(rank), int(world), int(device)
        h = C.c_void_p()
        _lib.check(_lib.lib().hr_comm_init(unique_id, self.rank, self.world, self.device, C.byref(h)))
        self._h = h

    @classmethod
    def from_torch_distributed(cls, device: int, group=None) -> "Comm":
        """Rendezvous over an initialised torch.distributed group: rank 0's unique id is broadcast."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return cls(0, 1, device)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        on_gpu = dist.get_backend(group) == "nccl"
        t = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", device) if on_gpu else "cpu")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        return cls(rank, world, device, bytes(t.cpu().numpy().tobytes()))

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().hr_comm_destroy(self._h)
                self._h = None
        except Exception:
            pass


class ShardedRetriever:
    """Hybrid retrieval over row shards.  `index` / `bm25` hold THIS rank's shard with id_base set
    to the shard's first global row and BM25 built with global statistics.

    One step = ONE C-ABI call, hr_retrieve_sharded: local BM25 + dense top-k_c into one packed block
    (D | S | I | J, 24 bytes per candidate, + a 16-byte trailer), ONE ncclAllGather of that block on the same
    stream, merge of the gathered per-rank blocks in place (rank order == id order, comparator (score, id))
    and fusion; nothing waits for the host in between."""

    def __init__(self, index, bm25=None, vector_weight: float = 0.7, bm25_weight: float = 0.3,
                 fusion: str = "weighted", group=None, comm: Comm | None = None):
        self.index, self.bm25, self.group = index, bm25, group
        self.vector_weight, self.bm25_weight, self.fusion = vector_weight, bm25_weight, fusion
        self.comm = comm if comm is not None else Comm.from_torch_distributed(index.device, group)

    @staticmethod
    def _views(buf, n):
        """(D float32[n], S float32[n], I int64[n], J int64[n]) views of one packed block (first 24n bytes)."""
        import torch
        f = buf[:8 * n].view(torch.float32)
        i = buf[8 * n:24 * n].view(torch.int64)
        return f[:n], f[n:2 * n], i[:n], i[n:2 * n]

    def retrieve(self, query_embeddings, query_tokens=None, top_k: int = 10, k_c: int | None = None):
        """query_embeddings: float32 [nq, d], torch CUDA (results stay on the device) or numpy (host in / host
        out), replicated on every rank.  Returns (scores [nq, top_k], ids [nq, top_k]) — identical on every rank."""
        from .retriever import candidate_depth, _MODES
        from .bm25 import query_csr
        kc = candidate_depth(top_k) if k_c is None else int(k_c)
        use_bm = self.bm25 is not None and query_tokens is not None
        L = _lib.lib()
        st = _lib.current_stream_ptr(self.index.device)
        on_dev = type(query_embeddings).__module__.startswith("torch") and getattr(query_embeddings, "is_cuda", False)
        if on_dev:
            import torch
            q = query_embeddings.to(torch.float32).contiguous()
            if q.dim() != 2 or q.shape[1] != self.index.d:
                raise AssertionError(f"retrieve: expected [nq, {self.index.d}] embeddings, got {tuple(q.shape)}")
            nq = q.shape[0]
            qi = qt = None
            if use_bm:
                ip, tm = query_csr(query_tokens)
                if not hasattr(ip, "is_cuda"):
                    ip = torch.from_numpy(np.ascontiguousarray(ip)).to(q.device)
                    tm = torch.from_numpy(np.ascontiguousarray(tm)).to(q.device)
                qi, qt = ip.to(torch.int32).contiguous(), tm.to(torch.int32).contiguous()
                if qi.shape[0] != nq + 1:
                    raise AssertionError("retrieve: query_tokens and query_embeddings disagree on nq")
            oS = torch.empty((nq, top_k), dtype=torch.float32, device=q.device)
            oI = torch.empty((nq, top_k), dtype=torch.int64, device=q.device)
            _lib.check(L.hr_retrieve_sharded(self.comm._h, self.index._h, self.bm25._h if use_bm else None,
                                             q.data_ptr(), qi.data_ptr() if use_bm else None,
                                             qt.data_ptr() if use_bm else None, nq, int(qt.numel()) if use_bm else 0,
                                             top_k, kc, _MODES[self.fusion], self.vector_weight, self.bm25_weight,
                                             oS.data_ptr(), oI.data_ptr(), 1, st))
            return oS, oI
        q = np.ascontiguousarray(query_embeddings, dtype=np.float32)
        if q.ndim != 2 or q.shape[1] != self.index.d:
            raise AssertionError(f"retrieve: expected [nq, {self.index.d}] embeddings, got {q.shape}")
        nq = q.shape[0]
        qi = qt = None
        if use_bm:
            qi, qt = query_csr(query_tokens)
            if len(qi) != nq + 1:
                raise AssertionError("retrieve: query_tokens and query_embeddings disagree on nq")
        oS = np.empty((nq, top_k), dtype=np.float32)
        oI = np.empty((nq, top_k), dtype=np.int64)
        _lib.check(L.hr_retrieve_sharded(self.comm._h, self.index._h, self.bm25._h if use_bm else None, q.ctypes.data,
                                         qi.ctypes.data if use_bm else None, qt.ctypes.data if use_bm else None, nq,
                                         int(qt.size) if use_bm else 0, top_k, kc, _MODES[self.fusion],
                                         self.vector_weight, self.bm25_weight, oS.ctypes.data, oI.ctypes.data, 0, st))
        return oS, oI
