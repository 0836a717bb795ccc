"""Seeded synthetic workloads of the shapes BASELINE.json names.

Dense: unit-norm gaussian rows (the HF provider normalises embeddings,
/root/reference/rag/providers/hf/embeddings.py:34); half of the queries are planted near a
corpus row, half are random directions (near-tie stress).  Sparse: Zipf(s=1) term popularity over
V terms, chunk length ~ clip(round(lognormal(ln 128, 0.4)), 16, 512) tokens (the chunker's
~100-170 words, /root/reference/rag/ingest/node_aware_chunker.py:50-52); queries hold 3..12
distinct terms drawn Zipf with the 64 most frequent ranks ("stop words") excluded.

numpy generators feed the CPU-sized parity tests (same arrays go to the oracle and to the GPU);
torch generators build the bench-sized corpora directly in HBM.
"""
from __future__ import annotations

import numpy as np

DENSE_SEED = 1234
QUERY_SEED = 4321
SPARSE_SEED = 777
STOP_RANKS = 64


# ---------------------------------------------------------------- numpy (tests, CPU baseline)
def dense_corpus_np(n: int, d: int, seed: int = DENSE_SEED) -> np.ndarray:
    rng = np.random.default_rng(seed)
    x = rng.standard_normal((n, d), dtype=np.float32)
    x /= np.linalg.norm(x, axis=1, keepdims=True)
    return x


def dense_queries_np(x: np.ndarray, nq: int, seed: int = QUERY_SEED, noise: float = 0.1) -> np.ndarray:
    rng = np.random.default_rng(seed)
    n, d = x.shape
    q = rng.standard_normal((nq, d), dtype=np.float32)
    half = nq // 2
    if n and half:
        rows = rng.integers(0, n, size=half)
        q[:half] = x[rows] + noise * q[:half]
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q.astype(np.float32)


def zipf_cdf(vocab: int) -> np.ndarray:
    p = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    return np.cumsum(p / p.sum())


def sparse_corpus_np(n_docs: int, vocab: int, seed: int = SPARSE_SEED, mean_len: float = 128.0):
    """Returns (term_ids int64[T], doc_ids int64[T], doc_len int32[n_docs]) token occurrences."""
    rng = np.random.default_rng(seed)
    dl = np.clip(np.round(rng.lognormal(np.log(mean_len), 0.4, size=n_docs)), 16, 512).astype(np.int32)
    total = int(dl.sum())
    cdf = zipf_cdf(vocab)
    t = np.minimum(np.searchsorted(cdf, rng.random(total)), vocab - 1).astype(np.int64)
    dd = np.repeat(np.arange(n_docs, dtype=np.int64), dl)
    return t, dd, dl


def sparse_queries_np(nq: int, vocab: int, seed: int = SPARSE_SEED + 1, stop: int = STOP_RANKS):
    """List of nq lists of distinct term ids (3..12 each), Zipf over ranks >= stop."""
    rng = np.random.default_rng(seed)
    stop = min(stop, max(vocab - 16, 0))
    p = 1.0 / np.arange(stop + 1, vocab + 1, dtype=np.float64)
    cdf = np.cumsum(p / p.sum())
    out = []
    for _ in range(nq):
        m = int(rng.integers(3, 13))
        terms = []
        while len(terms) < min(m, vocab - stop):
            t = int(min(np.searchsorted(cdf, rng.random()), len(cdf) - 1)) + stop
            if t not in terms:
                terms.append(t)
        out.append(terms)
    return out


# ---------------------------------------------------------------- torch (bench-sized, on device)
def dense_corpus_into(index, n: int, d: int, device, seed: int = DENSE_SEED, chunk_rows: int = 1 << 18,
                      keep_rows: int = 0):
    """Generate n unit-norm rows on `device` in chunks and add them to `index` (never touches host
    memory).  Returns the first `keep_rows` rows (a torch tensor on device) for planting queries."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    index.reserve(index.ntotal + n)
    kept = []
    have = 0
    for r0 in range(0, n, chunk_rows):
        nr = min(chunk_rows, n - r0)
        x = torch.randn((nr, d), generator=g, device=device, dtype=torch.float32)
        x = torch.nn.functional.normalize(x, dim=1)
        index.add(x)
        if have < keep_rows:
            take = min(keep_rows - have, nr)
            kept.append(x[:take].clone())
            have += take
        del x
    return torch.cat(kept) if kept else None


def dense_queries_torch(planted_rows, nq: int, d: int, device, seed: int = QUERY_SEED, noise: float = 0.1):
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    q = torch.randn((nq, d), generator=g, device=device, dtype=torch.float32)
    half = nq // 2
    if planted_rows is not None and half:
        idx = torch.randint(0, planted_rows.shape[0], (half,), generator=g, device=device)
        q[:half] = planted_rows[idx] + noise * q[:half]
    return torch.nn.functional.normalize(q, dim=1)


def sparse_corpus_csr_torch(n_docs: int, vocab: int, device, seed: int = SPARSE_SEED, mean_len: float = 128.0,
                            chunk_docs: int = 1 << 19):
    """Builds the CSR-by-term arrays on `device`: (indptr int64[V+1], post_doc int32[nnz],
    post_tf int32[nnz], doc_len int32[n_docs]).  One global sort of (term, doc) keys."""
    import torch
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    p = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64, device=device)
    cdf = torch.cumsum(p / p.sum(), 0).to(torch.float32)
    keys, tfs, dls = [], [], []
    for d0 in range(0, n_docs, chunk_docs):
        nd = min(chunk_docs, n_docs - d0)
        z = torch.randn(nd, generator=g, device=device)
        dl = torch.clamp(torch.round(torch.exp(np.log(mean_len) + 0.4 * z)), 16, 512).to(torch.int64)
        total = int(dl.sum().item())
        u = torch.rand(total, generator=g, device=device)
        t = torch.clamp(torch.searchsorted(cdf, u), max=vocab - 1)
        dd = torch.repeat_interleave(torch.arange(d0, d0 + nd, device=device, dtype=torch.int64), dl)
        key = t * n_docs + dd
        uk, cnt = torch.unique(key, return_counts=True)
        keys.append(uk)
        tfs.append(cnt.to(torch.int32))
        dls.append(dl.to(torch.int32))
        del z, u, t, dd, key
    key = torch.cat(keys)
    tf = torch.cat(tfs)
    del keys, tfs
    key, order = torch.sort(key)
    tf = tf[order]
    del order
    bounds = torch.arange(0, vocab + 1, device=device, dtype=torch.int64) * n_docs
    indptr = torch.searchsorted(key, bounds)
    post_doc = (key % n_docs).to(torch.int32)
    return indptr.to(torch.int64), post_doc, tf, torch.cat(dls)


def sparse_corpus_csr_chunked(n_docs: int, vocab: int, device, seed: int = SPARSE_SEED, mean_len: float = 128.0,
                              chunk_docs: int = 1 << 20):
    """Same distributions as sparse_corpus_csr_torch, built without a global sort (BASELINE configs[3]: 50M
    chunks, 1M-term vocabulary, ~5e9 postings): two passes over doc chunks, every chunk regenerated from its own
    seed.  Pass 1 counts df per term -> indptr; pass 2 scatters each chunk's (term, doc, tf) triples, already
    sorted by (term, doc), behind the postings earlier chunks wrote for the same term.  Peak memory = the
    CSR itself + one chunk.  Returns (indptr int64[V+1], post_doc int32[nnz], post_tf int32[nnz], doc_len int32[n])."""
    import torch
    p = 1.0 / torch.arange(1, vocab + 1, dtype=torch.float64, device=device)
    cdf = torch.cumsum(p / p.sum(), 0)

    def chunk(ci: int, d0: int, nd: int):
        g = torch.Generator(device=device)
        g.manual_seed(seed * 1000003 + ci)
        z = torch.randn(nd, generator=g, device=device)
        dl = torch.clamp(torch.round(torch.exp(np.log(mean_len) + 0.4 * z)), 16, 512).to(torch.int64)
        total = int(dl.sum().item())
        u = torch.rand(total, generator=g, device=device, dtype=torch.float64)
        t = torch.clamp(torch.searchsorted(cdf, u), max=vocab - 1)
        del u
        dd = torch.repeat_interleave(torch.arange(nd, device=device, dtype=torch.int64), dl)
        uk, cnt = torch.unique(t * nd + dd, return_counts=True)       # sorted by (term, local doc)
        del t, dd
        return uk // nd, (uk % nd + d0).to(torch.int32), cnt.to(torch.int32), dl.to(torch.int32)

    spans = [(ci, d0, min(chunk_docs, n_docs - d0)) for ci, d0 in enumerate(range(0, n_docs, chunk_docs))]
    df = torch.zeros(vocab, dtype=torch.int64, device=device)
    dls = []
    for ci, d0, nd in spans:
        terms, _, _, dl = chunk(ci, d0, nd)
        df += torch.bincount(terms, minlength=vocab)
        dls.append(dl)
        del terms
    indptr = torch.zeros(vocab + 1, dtype=torch.int64, device=device)
    indptr[1:] = torch.cumsum(df, 0)
    nnz = int(indptr[-1].item())
    post_doc = torch.empty(nnz, dtype=torch.int32, device=device)
    post_tf = torch.empty(nnz, dtype=torch.int32, device=device)
    fill = indptr[:-1].clone()
    for ci, d0, nd in spans:
        terms, docs, cnt, _ = chunk(ci, d0, nd)
        per_term = torch.bincount(terms, minlength=vocab)
        first = torch.cumsum(per_term, 0) - per_term                 # first triple of each term in this chunk
        pos = fill[terms] + (torch.arange(terms.numel(), device=device, dtype=torch.int64) - first[terms])
        post_doc[pos] = docs
        post_tf[pos] = cnt
        fill += per_term
        del terms, docs, cnt, per_term, first, pos
    return indptr, post_doc, post_tf, torch.cat(dls)


def sparse_queries_csr(nq: int, vocab: int, seed: int = SPARSE_SEED + 1, stop: int = STOP_RANKS):
    from .bm25 import query_csr
    return query_csr(sparse_queries_np(nq, vocab, seed, stop))
