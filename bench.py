#!/usr/bin/env python
"""bench.py — the hybrid retrieval hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic queries:
retrieve(query_embeddings, query_tokens, top_k=10) = dense flat scan (tcgen05 filter + exact fp32
re-score) + BM25 posting sweep + weighted fusion + top-k.  Workload at N=1 is BASELINE.json
configs[1] (C1): 10M x 1024-d fp32 flat index + BM25 over 10M chunks (30k-term Zipf vocabulary),
batch of 1024 queries, hybrid top-10.  At N>1 the SAME 10M-row corpus is row-sharded over the N
ranks (strong scaling: the metric is QPS on a fixed corpus); candidates travel through one NCCL
all-gather issued inside the library (hr_retrieve_sharded), every rank merges + fuses.

`value`  : whole-job QPS with queries already resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the public API with HOST (pinned) query buffers and the results read
           back to the host inside the timed region.
`roofline`: the dominant kernel (scan_tc2_kernel) timed live with CUDA events on its launch stream;
           `roofline.bm25` the BM25 search (plan + sweep + merge kernels).
`parity_check`: every N: the timed answer of 32 queries against an fp64 evaluation of the same
           definitions on the device (torch matmul / index_add over the same shards, all-gathered).
`c4_latency`: every N: BASELINE configs[4], batch-1 hybrid retrieve(), host in / host out, p50/p99.
`c3`     : N=1: BASELINE configs[3], BM25-only stress (50M chunks, 1M-term vocabulary, 4096 queries).
`c2`     : N=8: BASELINE configs[2], 100M x 1024 bf16 rows row-sharded, hybrid top-100.
`cpu_baseline`: the CPU restatement (oracle/fast.py: sgemm + top-k, sparse product, fusion;
           faiss-cpu 1.7.4 is not installable offline) on a bounded row sample, all host threads.
`--impl reference`: that same CPU path alone (rank 0 only under torchrun), no GPU code involved; its
           line reports the MEASURED sample as its own config and the full-corpus projection separately.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if "--impl" in sys.argv and "reference" in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1: the CPU arm uses every host core, whatever launched it
    for _v in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS"):
        os.environ[_v] = str(os.cpu_count() or 1)

import numpy as np  # noqa: E402

METRIC = "hybrid QPS (top-10, 10M x 1024d)"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--rows", type=int, default=int(os.getenv("HR_BENCH_ROWS", 10_000_000)))
    ap.add_argument("--dim", type=int, default=1024)
    ap.add_argument("--nq", type=int, default=int(os.getenv("HR_BENCH_NQ", 1024)))
    ap.add_argument("--vocab", type=int, default=30_000)
    ap.add_argument("--topk", type=int, default=10)
    ap.add_argument("--cpu-sample-rows", type=int, default=int(os.getenv("HR_BENCH_CPU_ROWS", 500_000)))
    ap.add_argument("--storage", default=os.getenv("HR_BENCH_STORAGE", "f32+bf16"), choices=["f32", "f32+bf16", "bf16"],
                    help="f32: fp32 rows, TF32 filter; f32+bf16: fp32 rows + bf16 shadow for the filter (same answers, "
                         "1.5x the memory); bf16: bf16 rows")
    ap.add_argument("--index-metric", default="ip", choices=["ip", "l2"],
                    help="ip: IndexFlatIP (BASELINE.json); l2: IndexFlatL2, what the reference builds "
                         "(rag/storage/faiss_index.py:123); same ranking on unit-norm rows")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the nq sweep of the dense scan")
    ap.add_argument("--no-c3", action="store_true", help="skip the BM25 stress config (N=1)")
    ap.add_argument("--no-c2", action="store_true", help="skip the 100M-row config (N=8)")
    ap.add_argument("--no-c4", action="store_true", help="skip the batch-1 latency loop")
    ap.add_argument("--no-f32", action="store_true", help="skip the comparison run on plain fp32 storage (N=1)")
    ap.add_argument("--c3-rows", type=int, default=int(os.getenv("HR_BENCH_C3_ROWS", 50_000_000)))
    ap.add_argument("--c3-vocab", type=int, default=1_000_000)
    ap.add_argument("--c3-nq", type=int, default=4096)
    ap.add_argument("--c2-rows", type=int, default=int(os.getenv("HR_BENCH_C2_ROWS", 100_000_000)))
    ap.add_argument("--c2-min-gpus", type=int, default=8)
    ap.add_argument("--c4-calls", type=int, default=1000)
    ap.add_argument("--parity-queries", type=int, default=32)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": j["hbm_gbs"], "bf16_tflops": j["bf16_tflops"],
                "bf16_tflops_sustained": j.get("bf16_tflops_sustained", j["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_reference(args, steps, warmup):
    """The reference's CPU path for this workload on the host cores (oracle/fast.py): faiss-equivalent
    exhaustive search (sgemm blocks + top-k, the algorithm faiss 1.7.4 uses for nq >= 20), BM25 as one sparse
    product, weighted fusion.  Bounded sample: the first `cpu_sample_rows` rows of a corpus with the same
    distributions; the line reports what was MEASURED on that sample, the full-corpus projection
    (time scales linearly in rows) is a separate field."""
    import torch
    import intool_rag_b200  # noqa: F401  (generators only; no CUDA code runs in this arm)
    from intool_rag_b200 import synth
    from oracle import bm25 as obm25
    from oracle import fast, flat
    cores = int(os.cpu_count() or 1)
    torch.set_num_threads(cores)
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=cores)
    except Exception:
        pass
    n = min(args.cpu_sample_rows, args.rows)
    x = synth.dense_corpus_np(n, args.dim)
    q = synth.dense_queries_np(x, args.nq)
    t, dd, dl = synth.sparse_corpus_np(n, args.vocab)
    qs = synth.sparse_queries_np(args.nq, args.vocab)
    l2 = args.index_metric == "l2"
    corpus = obm25.BM25Corpus.from_token_matrix(t, dd, dl, args.vocab)
    fbm = fast.FastBM25(corpus)
    times, stages = [], {}
    for i in range(warmup + steps):
        tm = {} if i < warmup else stages
        t0 = time.perf_counter()
        fs, fi = fast.retrieve(x, l2, fbm, q, qs, args.topk, timings=tm)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = float(np.sum(times))
    ms_step = 1e3 * tot / len(times)
    qps_sample = args.nq * len(times) / tot
    # recall@10 of the fast path against the fp64 oracle on a query subset
    nsub = min(32, args.nq)
    oi = (flat.IndexFlatL2 if l2 else flat.IndexFlatIP)(args.dim)
    oi.add(x)
    from oracle import hybrid
    _, ri, _ = hybrid.retrieve(oi, corpus, q[:nsub], qs[:nsub], args.topk, precision="f64")
    hits = sum(len(set(a.tolist()) & set(b[b >= 0].tolist())) for a, b in zip(fi[:nsub], ri))
    recall = hits / max(1, int((ri >= 0).sum()))
    # what the reference service really does: one query at a time (rag/storage/faiss_index.py:81)
    n1 = min(16, args.nq)
    t0 = time.perf_counter()
    for i in range(n1):
        fast.retrieve(x, l2, fbm, q[i:i + 1], qs[i:i + 1], args.topk)
    ms_nq1 = 1e3 * (time.perf_counter() - t0) / n1
    scale = n / args.rows
    return {"value": qps_sample, "unit": UNIT, "cores": cores, "blas_threads": int(torch.get_num_threads()),
            "kind": "port", "rows": n, "ms_per_step": ms_step, "steps": len(times),
            "stage_ms": {k[:-2]: 1e3 * v / len(times) for k, v in stages.items()},
            "nq1_ms_per_query": ms_nq1, "recall_at_10_vs_fp64": recall,
            "extrapolated": {"rows": args.rows, "value": qps_sample * scale, "ms_per_step": ms_step / scale,
                             "nq1_ms_per_query": ms_nq1 / scale,
                             "how": f"time scales linearly in rows: x {args.rows}/{n}"},
            "sample": f"{args.nq} queries x first {n} of {args.rows} rows (sgemm + top-k, sparse-product BM25, fusion), "
                      f"{len(times)} step(s) of {ms_step / 1e3:.2f} s on {cores} threads; faiss-cpu 1.7.4 not installable "
                      "offline -> torch/scipy restatement (oracle/fast.py)"}


def run_reference(args):
    rank = int(os.getenv("RANK", "0"))
    if rank != 0:
        return
    # the requested K / W are honoured (a step of the 500k-row sample takes ~5 s on 16 cores); only absurd values are
    # clamped so that the arm always ends within a few minutes
    steps, warmup = max(1, min(args.steps, 20)), max(0, min(args.warmup, 5))
    cb = cpu_reference(args, steps, warmup)
    cfg = workload_config(args, args.gpus)
    cfg["rows"] = cb["rows"]
    cfg["workload"] = ("CPU arm, bounded sample of BASELINE configs[1]: " + cb["sample"])
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": cb["steps"], "warmup": warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "cpu_baseline": cb, "extrapolated": cb["extrapolated"],
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_json_out, flush=True)


def workload_config(args, n_gpus):
    mem = {"f32": 1.0, "f32+bf16": 1.5, "bf16": 0.5}[args.storage]
    return {"workload": f"BASELINE configs[1]: {args.rows}x{args.dim} fp32 flat index (IndexFlat{args.index_metric.upper()} semantics, storage {args.storage}) + BM25 over "
                        f"{args.rows} chunks ({args.vocab}-term Zipf vocabulary), batch {args.nq} queries, hybrid top-{args.topk}, "
                        "weighted fusion 0.7/0.3, candidate depth 50",
            "rows": args.rows, "dim": args.dim, "nq": args.nq, "top_k": args.topk, "vocab": args.vocab, "storage": args.storage,
            "storage_bytes": int(args.rows * args.dim * 4 * mem),
            "storage_note": "f32+bf16 keeps the fp32 rows plus a bf16 shadow the filter streams (1.5x the fp32 bytes); answers are "
                            "bit-identical to storage f32 (TF32 filter), which is the faiss-shaped module's default and runs the "
                            "scan at about half the rate (roofline.f32_default)",
            "sharding": f"row-sharded over {n_gpus} GPU(s), one NCCL all-gather of k_c candidates inside hr_retrieve_sharded" if n_gpus > 1 else "single GPU",
            "l2_hygiene": "inputs larger than L2 (20-41 GB of rows + 5-19 GB of postings streamed per step vs 126 MB L2)"}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
        self.proc = None
        self.idx = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_median": float(np.median(pw)) if pw else None}
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


# ------------------------------------------------------------------------------------ fp64 truth on the device
def fp64_hybrid_truth(torch, dist, world, dev, dense_chunks, q_sub, tokens_sub, csr, stats, lo, kc, top_k, metric_l2):
    """fp64 evaluation of the oracle's definitions (oracle/{flat,bm25,fusion}.py) on THIS rank's shard with torch,
    candidates of all ranks all-gathered, fused on every rank.  dense_chunks: iterator of (row0, fp32 rows)
    regenerated from the corpus seed; csr = (indptr, post_doc, post_tf, doc_len) of the local shard;
    stats = (df_global int64[V], n_docs_global, avgdl_global).  Returns (scores fp64 [nq, top_k'], ids)."""
    nq = q_sub.shape[0]
    q64 = q_sub.double()
    bs = torch.full((nq, kc), -float("inf"), dtype=torch.float64, device=dev)
    bi = torch.full((nq, kc), -1, dtype=torch.int64, device=dev)
    for row0, x in dense_chunks:
        s = q64 @ x.double().T
        if metric_l2:
            s = -((q64 * q64).sum(1, keepdim=True) + (x.double() ** 2).sum(1)[None, :] - 2.0 * s).clamp_(min=0.0)
        v, i = torch.topk(s, min(kc, s.shape[1]), dim=1)
        cs, ci = torch.cat([bs, v], 1), torch.cat([bi, i + row0 + lo], 1)
        bs, sel = torch.topk(cs, kc, dim=1)
        bi = torch.gather(ci, 1, sel)
    indptr, post_doc, post_tf, doc_len = csr
    df_g, n_g, avgdl_g = stats
    n_local = doc_len.shape[0]
    idf = torch.log((n_g - df_g.double() + 0.5) / (df_g.double() + 0.5) + 1.0)
    ss = torch.zeros((nq, kc), dtype=torch.float64, device=dev)
    si = torch.full((nq, kc), -1, dtype=torch.int64, device=dev)
    V = idf.shape[0]
    for qi, terms in enumerate(tokens_sub):
        acc = torch.zeros(n_local, dtype=torch.float64, device=dev)
        for t in terms:                                      # duplicates count per occurrence
            if not (0 <= t < V):
                continue
            a, b = int(indptr[t].item()), int(indptr[t + 1].item())
            if b <= a:
                continue
            docs = post_doc[a:b].long()
            tf = post_tf[a:b].double()
            dl = doc_len[docs].double()
            acc.index_add_(0, docs, idf[t] * tf * 2.5 / (tf + 1.5 * (0.25 + 0.75 * dl / avgdl_g)))
        kk = min(kc, n_local)
        v, i = torch.topk(acc, kk)
        kth = v[-1]
        cand = torch.nonzero(acc >= torch.clamp(kth, min=1e-300)).flatten()     # all boundary ties, score > 0 only
        order = torch.argsort(cand, stable=True)
        cand = cand[order]
        o2 = torch.argsort(-acc[cand], stable=True)[:kk]                       # (score desc, id asc)
        sel = cand[o2]
        ss[qi, :sel.numel()] = acc[sel]
        si[qi, :sel.numel()] = sel + lo
    if world > 1:
        def gather(tn):
            out = torch.empty((world,) + tuple(tn.shape), dtype=tn.dtype, device=dev)
            dist.all_gather_into_tensor(out, tn.contiguous())
            return out.permute(1, 0, 2).reshape(tn.shape[0], world * tn.shape[1])
        gs, gi, hs, hi_ = gather(bs), gather(bi), gather(ss), gather(si)
        bs, sel = torch.topk(gs, kc, dim=1)
        bi = torch.gather(gi, 1, sel)
        # sparse lists: ties need (score desc, id asc) across ranks; rank order == id order and topk is not stable,
        # so sort explicitly
        key = torch.argsort(hi_.masked_fill(hi_ < 0, 1 << 62), dim=1, stable=True)
        hs, hi_ = torch.gather(hs, 1, key), torch.gather(hi_, 1, key)
        key = torch.argsort(-hs, dim=1, stable=True)[:, :kc]
        ss, si = torch.gather(hs, 1, key), torch.gather(hi_, 1, key)
    # weighted fusion 0.7 / 0.3 over the union of the two depth-kc lists (oracle/fusion.py)
    bs, bi, ss, si = bs.cpu().numpy(), bi.cpu().numpy(), ss.cpu().numpy(), si.cpu().numpy()
    out_s = np.zeros((nq, 2 * top_k))
    out_i = np.full((nq, 2 * top_k), -1, np.int64)
    for qi in range(nq):
        fused = {}
        for s, i in zip(bs[qi], bi[qi]):
            if i >= 0:
                sim = (1.0 + s / 2.0) if metric_l2 else s       # s = -dist for L2
                fused[int(i)] = 0.7 * min(1.0, max(0.0, float(sim)))
        mx = float(ss[qi, 0]) if si[qi, 0] >= 0 else 0.0
        for s, i in zip(ss[qi], si[qi]):
            if i >= 0 and mx > 0:
                fused[int(i)] = fused.get(int(i), 0.0) + 0.3 * float(s) / mx
        items = sorted(fused.items(), key=lambda kv: (-kv[1], kv[0]))[:2 * top_k]
        for j, (i, s) in enumerate(items):
            out_s[qi, j], out_i[qi, j] = s, i
    return out_s, out_i


def compare_with_truth(S, I, truth_s, truth_i, top_k, rtol=1e-5):
    """ids must match position by position, except inside runs of truth scores closer than rtol (documented
    near-tie exemption); fused scores within rtol."""
    nq = I.shape[0]
    exact, tie_ok, bad, max_rel = 0, 0, 0, 0.0
    for qi in range(nq):
        for j in range(top_k):
            ti, ts = int(truth_i[qi, j]), float(truth_s[qi, j])
            if ti < 0 and int(I[qi, j]) < 0:
                exact += 1
                continue
            max_rel = max(max_rel, abs(float(S[qi, j]) - ts) / max(abs(ts), 1e-12))
            if int(I[qi, j]) == ti:
                exact += 1
                continue
            pos = np.nonzero(truth_i[qi] == int(I[qi, j]))[0]
            if pos.size and abs(float(truth_s[qi, pos[0]]) - ts) <= 2 * rtol * max(abs(ts), 1e-12):
                tie_ok += 1
            else:
                bad += 1
    return {"queries_checked": nq, "positions": nq * top_k, "ids_equal": exact, "ids_swapped_inside_near_ties": tie_ok,
            "ids_wrong": bad, "max_rel_score_err": max_rel, "score_rtol": rtol,
            "hybrid_top10_ids_equal_fp64": bad == 0 and max_rel <= rtol}


# ------------------------------------------------------------------------------------ GPU arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import intool_rag_b200  # noqa: F401
    from intool_rag_b200 import _lib, synth
    from intool_rag_b200 import bm25 as pbm25
    from intool_rag_b200 import faiss as hf
    from intool_rag_b200.retriever import HybridRetriever
    from intool_rag_b200.sharded import ShardedRetriever, shard_bounds, global_bm25_stats

    world = int(os.getenv("WORLD_SIZE", "1"))
    rank = int(os.getenv("RANK", "0"))
    local = int(os.getenv("LOCAL_RANK", "0"))
    _lib.require_gpu()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    pk = peaks()
    log = (lambda *a: print(*a, file=sys.stderr, flush=True)) if rank == 0 else (lambda *a: None)
    metric_l2 = args.index_metric == "l2"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def event_ms(fn, reps):
        ms = []
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        return float(np.median(ms)), ms

    # ---- C3 first (N=1): it needs the memory the C1 corpus will occupy --------------------------------
    c3 = None
    if world == 1 and not args.no_c3:
        t0 = time.time()
        ip3, pd3, tf3, dl3 = synth.sparse_corpus_csr_chunked(args.c3_rows, args.c3_vocab, dev, seed=synth.SPARSE_SEED + 31)
        nnz3 = int(ip3[-1].item())
        bm3 = pbm25.BM25Index.from_csr(ip3, pd3, tf3, dl3, args.c3_vocab, device=local)
        qs3 = synth.sparse_queries_np(args.c3_nq, args.c3_vocab, seed=synth.SPARSE_SEED + 32)
        qi3, qt3 = pbm25.query_csr(qs3)
        qi3d, qt3d = torch.from_numpy(qi3).to(dev), torch.from_numpy(qt3).to(dev)
        for _ in range(2):
            S3, I3, touched3 = bm3.search((qi3d, qt3d), 50, return_postings=True)
        ms3, all3 = event_ms(lambda: bm3.search((qi3d, qt3d), 50), 5)
        # spot check: 8 queries against an fp64 scatter on the device
        df3 = (ip3[1:] - ip3[:-1])
        avg3 = float(dl3.double().mean().item())
        idf3 = torch.log((args.c3_rows - df3.double() + 0.5) / (df3.double() + 0.5) + 1.0)
        worst, ids_ok = 0.0, True
        for qi in range(8):
            acc = torch.zeros(args.c3_rows, dtype=torch.float64, device=dev)
            for t in qs3[qi]:
                a, b = int(ip3[t].item()), int(ip3[t + 1].item())
                docs = pd3[a:b].long()
                tf = tf3[a:b].double()
                acc.index_add_(0, docs, idf3[t] * tf * 2.5 / (tf + 1.5 * (0.25 + 0.75 * dl3[docs].double() / avg3)))
            v, i = torch.topk(acc, 50)
            got_s, got_i = S3[qi].double(), I3[qi]
            worst = max(worst, float(((got_s - v).abs() / v.clamp(min=1e-30)).max().item()))
            ids_ok = ids_ok and bool((acc[got_i] - v).abs().max().item() <= 1e-5 * float(v[0].item()))
        c3 = {"workload": f"BASELINE configs[3]: BM25 only, {args.c3_rows} chunks, {args.c3_vocab}-term Zipf vocabulary, "
                          f"{args.c3_nq} queries (3..12 distinct terms, 64 stop ranks excluded), top-50",
              "rows": args.c3_rows, "vocab": args.c3_vocab, "nq": args.c3_nq, "nnz": nnz3, "ms": ms3, "ms_all": all3,
              "postings": int(touched3), "postings_per_s": touched3 / (ms3 / 1e3),
              "achieved": touched3 * 8 / (ms3 / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
              "frac": touched3 * 8 / (ms3 / 1e3) / 1e9 / pk["hbm_gbs"],
              "qps": args.c3_nq / (ms3 / 1e3), "build_s": time.time() - t0,
              "parity_check": {"queries_checked": 8, "max_rel_score_err_vs_fp64": worst,
                               "returned_docs_have_the_top50_scores": ids_ok}}
        log(f"[c3] {args.c3_rows} docs V={args.c3_vocab} nnz={nnz3}: {ms3:.2f} ms, {c3['achieved']:.0f} GB/s algorithmic "
            f"({c3['frac']:.3f} of peak), worst rel err {worst:.2e}, setup {time.time() - t0:.0f}s")
        # worst case: the same index, queries drawn WITHOUT the stop-rank exclusion: a third of the
        # terms are among the 64 most frequent ones, whose lists hold 10-100 % of the chunks
        nqw = min(512, args.c3_nq)
        qsw = synth.sparse_queries_np(nqw, args.c3_vocab, seed=synth.SPARSE_SEED + 33, stop=0)
        qiw, qtw = pbm25.query_csr(qsw)
        qiwd, qtwd = torch.from_numpy(qiw).to(dev), torch.from_numpy(qtw).to(dev)
        Sw, Iw, touchedw = bm3.search((qiwd, qtwd), 50, return_postings=True)
        msw, allw = event_ms(lambda: bm3.search((qiwd, qtwd), 50), 3)
        worstw, okw = 0.0, True
        for qi in range(2):
            acc = torch.zeros(args.c3_rows, dtype=torch.float64, device=dev)
            for t in qsw[qi]:
                a, b = int(ip3[t].item()), int(ip3[t + 1].item())
                docs = pd3[a:b].long()
                tf = tf3[a:b].double()
                acc.index_add_(0, docs, idf3[t] * tf * 2.5 / (tf + 1.5 * (0.25 + 0.75 * dl3[docs].double() / avg3)))
            v, i = torch.topk(acc, 50)
            worstw = max(worstw, float(((Sw[qi].double() - v).abs() / v.clamp(min=1e-30)).max().item()))
            okw = okw and bool((acc[Iw[qi]] - v).abs().max().item() <= 1e-5 * float(v[0].item()))
        c3["no_stop_exclusion"] = {
            "workload": f"same index, {nqw} queries drawn over ALL ranks (no stop-rank exclusion), top-50",
            "nq": nqw, "ms": msw, "ms_all": allw, "postings": int(touchedw),
            "postings_per_query": touchedw / nqw, "achieved": touchedw * 8 / (msw / 1e3) / 1e9, "unit": "GB/s",
            "frac": touchedw * 8 / (msw / 1e3) / 1e9 / pk["hbm_gbs"], "qps": nqw / (msw / 1e3),
            "parity_check": {"queries_checked": 2, "max_rel_score_err_vs_fp64": worstw,
                             "returned_docs_have_the_top50_scores": okw}}
        log(f"[c3] no stop exclusion, {nqw} queries: {msw:.1f} ms, {touchedw / nqw / 1e6:.1f}M postings per query, "
            f"{c3['no_stop_exclusion']['achieved']:.0f} GB/s algorithmic, worst rel err {worstw:.2e}")
        del Sw, Iw
        del bm3, ip3, pd3, tf3, dl3, S3, I3, acc, df3, idf3
        torch.cuda.empty_cache()

    # ---- build this rank's C1 shard on device -----------------------------------------------------------
    t0 = time.time()
    lo, hi = shard_bounds(args.rows, world, rank)
    n_local = hi - lo
    ix = (hf.IndexFlatL2 if metric_l2 else hf.IndexFlatIP)(args.dim, device=local, storage=args.storage)
    ix.set_id_base(lo)
    planted = synth.dense_corpus_into(ix, n_local, args.dim, dev, seed=synth.DENSE_SEED + rank, keep_rows=4096)
    if world > 1:
        dist.broadcast(planted, src=0)
    q_dev = synth.dense_queries_torch(planted, args.nq, args.dim, dev)
    log(f"[bench] rank0 dense shard {n_local}x{args.dim} in HBM after {time.time() - t0:.1f}s")
    indptr, post_doc, post_tf, doc_len = synth.sparse_corpus_csr_torch(n_local, args.vocab, dev,
                                                                        seed=synth.SPARSE_SEED + rank)
    df_local = (indptr[1:] - indptr[:-1]).contiguous()
    df_g, n_g, avgdl_g = global_bm25_stats(df_local, n_local, int(doc_len.sum().item()))
    bm = pbm25.BM25Index.from_csr(indptr, post_doc, post_tf, doc_len, args.vocab, n_docs_global=n_g,
                                  avgdl_global=avgdl_g, df_global=df_g, device=local)
    bm.set_id_base(lo)
    nnz_local = bm.nnz
    qs_list = synth.sparse_queries_np(args.nq, args.vocab)
    qi_np, qt_np = pbm25.query_csr(qs_list)
    qi_dev = torch.from_numpy(qi_np).to(dev)
    qt_dev = torch.from_numpy(qt_np).to(dev)
    log(f"[bench] BM25 shard: {nnz_local} postings; total setup {time.time() - t0:.1f}s")

    engine = HybridRetriever(ix, bm) if world == 1 else ShardedRetriever(ix, bm)
    step_dev = lambda: engine.retrieve(q_dev, (qi_dev, qt_dev), args.topk)  # noqa: E731

    # pinned host copies for the end-to-end arm
    q_host = torch.empty((args.nq, args.dim), dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    qi_host = torch.from_numpy(qi_np).pin_memory()
    qt_host = torch.from_numpy(qt_np).pin_memory()
    h2d = q_host.numel() * 4 + qi_host.numel() * 4 + qt_host.numel() * 4
    d2h = args.nq * args.topk * 12
    step_e2e = lambda: engine.retrieve(q_host.numpy(), (qi_host.numpy(), qt_host.numpy()), args.topk)  # noqa: E731

    dense_ms = [0.0]    # CUDA-event time of the dense part (pad .. fallback kernels) of a step, mean of the last timed() call

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on the launch stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scan_ms, flagged, deeper = 0.0, 0, 0
        dense_ms[0] = 0.0
        e0.record()
        for _ in range(steps):
            fn()
            st = ix.stats()
            scan_ms += st["scan_ms"]
            dense_ms[0] += st["total_ms"] / steps
            flagged += st["flagged"]
            deeper += st["deeper"]
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        cnt = torch.tensor([flagged, deeper], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt)     # fallbacks on ANY rank delay every rank at the all-gather
        return float(ms.item()), scan_ms / steps, (int(cnt[0].item()), int(cnt[1].item()))

    for _ in range(max(args.warmup, 3)):
        out = step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    total_ms, scan_ms, flagged = timed(step_dev, args.steps)
    dense_in_step_ms = dense_ms[0]
    launches = _lib.launch_count() - l0
    for _ in range(2):
        out_h = step_e2e()
    e2e_ms, _, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- parity at full size, every N: the timed answers of a query subset against fp64 on the device -----------
    npar = min(args.parity_queries, args.nq)

    def dense_chunks():
        g = torch.Generator(device=dev)
        g.manual_seed(synth.DENSE_SEED + rank)
        for r0 in range(0, n_local, 1 << 18):
            nr = min(1 << 18, n_local - r0)
            x = torch.nn.functional.normalize(torch.randn((nr, args.dim), generator=g, device=dev, dtype=torch.float32), dim=1)
            yield r0, x
    truth_s, truth_i = fp64_hybrid_truth(torch, dist, world, dev, dense_chunks(), q_dev[:npar], qs_list[:npar],
                                         (indptr, post_doc, post_tf, doc_len), (df_g, n_g, avgdl_g), lo, 50, args.topk,
                                         metric_l2)
    S, I = out
    check = compare_with_truth(S[:npar].cpu().numpy(), I[:npar].cpu().numpy(), truth_s, truth_i, args.topk)
    check["e2e_answer_equals_device_answer"] = bool(np.array_equal(out_h[1], I.cpu().numpy()) and
                                                    np.array_equal(out_h[0], S.cpu().numpy()))
    if world == 1:
        sub = q_dev[:16].contiguous()
        D_auto, I_auto = ix.search(sub, 50)
        ix.set_mode("exact")
        D_ex, I_ex = ix.search(sub, 50)
        ix.set_mode("auto")
        check.update({"dense_top50_ids_equal_exact_scan": bool(torch.equal(I_auto, I_ex)),
                      "dense_scores_equal_exact_scan": bool(torch.equal(D_auto, D_ex))})
    del indptr, post_doc, post_tf, doc_len, truth_s, truth_i
    torch.cuda.empty_cache()

    # ---- BM25 search alone (second roofline) -----------------------------------------------------------
    barrier()
    _, _, touched = bm.search((qi_dev, qt_dev), 50, return_postings=True)
    bm_ms, _ = event_ms(lambda: bm.search((qi_dev, qt_dev), 50), 5)

    # ---- C4: batch-1 latency, host in / host out, every N (all ranks call in lockstep) ----------------------
    c4 = None
    if not args.no_c4:
        qh = q_host.numpy()
        calls = max(args.c4_calls, 10)
        for i in range(20):
            engine.retrieve(qh[i % args.nq:i % args.nq + 1], [qs_list[i % args.nq]], args.topk)
        barrier()
        lat = np.empty(calls)
        for i in range(calls):
            j = i % args.nq
            t1 = time.perf_counter()
            engine.retrieve(qh[j:j + 1], [qs_list[j]], args.topk)
            lat[i] = time.perf_counter() - t1
        barrier()
        filt = 4 if args.storage == "f32" else 2
        c4 = {"workload": f"BASELINE configs[4]: batch-1 hybrid retrieve() over {args.rows}x{args.dim} rows on {world} GPU(s), "
                          "host query in, host top-10 out, wall clock per call on rank 0",
              "calls": calls, "p50_ms": float(np.percentile(lat, 50) * 1e3), "p99_ms": float(np.percentile(lat, 99) * 1e3),
              "mean_ms": float(lat.mean() * 1e3), "max_ms": float(lat.max() * 1e3),
              "hbm_floor_ms": {"filter_rows_streamed_once": n_local * args.dim * filt / (pk["hbm_gbs"] * 1e9) * 1e3,
                               "bf16_rows": n_local * args.dim * 2 / (pk["hbm_gbs"] * 1e9) * 1e3,
                               "fp32_rows": n_local * args.dim * 4 / (pk["hbm_gbs"] * 1e9) * 1e3},
              "fallback_queries": int(ix.stats()["flagged"])}
        log(f"[c4] N={world}: p50 {c4['p50_ms']:.3f} ms  p99 {c4['p99_ms']:.3f} ms over {calls} calls")

    ms_per_step = total_ms / args.steps
    qps = args.nq / (ms_per_step / 1e3)
    e2e_qps = args.nq / (e2e_ms / args.steps / 1e3)
    flops = 2.0 * args.nq * n_local * args.dim            # per launch of this rank's scan kernel
    filt_elem = 4 if args.storage == "f32" else 2     # bytes per element of the rows the filter streams
    corpus_bytes = float(n_local) * args.dim * filt_elem
    scan_s = max(scan_ms, 1e-6) / 1e3
    t_hbm = corpus_bytes / (pk["hbm_gbs"] * 1e9)
    t_tensor = flops / (pk["bf16_tflops_sustained"] * 1e12)
    bound = "tensor" if t_tensor > t_hbm else "hbm"
    if bound == "tensor":
        achieved, peak, runit = flops / scan_s / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s"
    else:
        achieved, peak, runit = corpus_bytes / scan_s / 1e9, pk["hbm_gbs"], "GB/s"
    kname = ("scan_tc2_kernel" if args.nq > 128 else "scan_tc_kernel") + (f"<tf32,{args.index_metric}>" if args.storage == "f32" else f"<bf16,{args.index_metric}>")
    note = ("kind::tf32 MMA runs at half the bf16 rate the peak was measured with (cuBLAS bf16, sustained)"
            if args.storage == "f32" else
            "kind::f16 (bf16 operands, fp32 accumulate in TMEM) over the bf16 rows; answers are exact fp32 after the re-score; "
            "peak = sustained cuBLAS bf16 (the kernel is timed inside a long step)")
    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture of this
    # exact default workload (profiles/r2_scan_ncu.md, round 2; round 1: 20.83 GB); other workloads: not captured
    traffic = None
    if (world, args.rows, args.dim, args.nq, args.storage) == (1, 10_000_000, 1024, 1024, "f32+bf16"):
        traffic = 20.756e9 + 0.0375e9
    t_min_step = max(t_hbm, t_tensor) + touched * 8 / (pk["hbm_gbs"] * 1e9)
    roofline = {"kernel": kname, "bound": bound, "achieved": achieved, "peak": peak, "unit": runit,
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu --set full, profiles/r2_scan_ncu.md)",
                "peak_source": pk["src"],
                "kernel_ms": scan_ms, "share_of_step": scan_ms / ms_per_step,
                "step_split_ms": {"dense_search_total": dense_in_step_ms, "scan_kernel": scan_ms,
                                  "bm25_fusion_exchange_and_gaps": ms_per_step - dense_in_step_ms,
                                  "note": "CUDA events of the library around the dense part of the timed steps; the rest of the step "
                                          "is the BM25 search (timed alone below: roofline.bm25), fusion, at N > 1 the all-gather and "
                                          "the merges, and what the power-capped clock costs the kernels that follow the scan"},
                "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": corpus_bytes,
                "hbm_gbs_algorithmic": corpus_bytes / scan_s / 1e9,
                "hbm_frac_algorithmic": corpus_bytes / scan_s / 1e9 / pk["hbm_gbs"],
                "note": note,
                "whole_step_frac_of_t_min": t_min_step * 1e3 / ms_per_step,
                "bm25": {"kernel": "bm25_sweep_kernel (+ plan and merge kernels)", "bound": "hbm", "postings": int(touched), "ms": bm_ms,
                         "achieved": touched * 8 / (bm_ms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": touched * 8 / (bm_ms / 1e3) / 1e9 / pk["hbm_gbs"],
                         # dram bytes of the sweep kernel + the two cursor-plan kernels, ncu capture of this default
                         # workload (profiles/r2_bm25_sweep_ncu.md); other workloads: not captured
                         "traffic": (5.859e9 + 0.022e9 + 0.645e9 + 2.418e9 + 0.208e9) if traffic is not None else None,
                         "algorithmic_bytes": "8 B per posting of every (query, term) pair; lists shared by the queries of a "
                                              "batch are read from HBM once per doc window and from L2 afterwards "
                                              "(profiles/r2_bm25_sweep_ncu.md)"}}
    line = {"metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "roofline": roofline,
            "fallback_queries_in_timed_region": int(flagged[0]),
            "second_stage_rescore_queries_in_timed_region": int(flagged[1]), "parity_check": check}
    if c4 is not None:
        line["c4_latency"] = c4
    if c3 is not None:
        line["c3"] = c3

    if world == 1 and not args.no_sweep:
        # the HBM-bound regime of the same scan kernel family: top-10 dense search for growing batches (scan kernel
        # time from the library's CUDA events; bytes = the rows the filter streams, once per batch)
        sweep = {}
        for nq in (1, 8, 32, 128, 256, 512, 1024):
            if nq > args.nq:
                break
            qq = q_dev[:nq].contiguous()
            # the board is power-capped: the SM clock needs a few hundred ms to settle after the batch size changes
            # (profiles/r2_nq_dip.md), so every point runs back to back for ~0.9 s and reports the median of the
            # last two thirds
            ms, t_start = [], time.time()
            while time.time() - t_start < 0.9 and len(ms) < 400:
                ix.search(qq, 10)
                ms.append(ix.stats()["scan_ms"])
            m = float(np.median(ms[len(ms) // 3:]))
            fl = 2.0 * nq * n_local * args.dim
            tmin = max(corpus_bytes / (pk["hbm_gbs"] * 1e9), fl / (pk["bf16_tflops_sustained"] * 1e12)) * 1e3
            sweep[str(nq)] = {"scan_ms": m, "hbm_gbs_algorithmic": corpus_bytes / m / 1e6,
                              "hbm_frac": corpus_bytes / m / 1e6 / pk["hbm_gbs"],
                              "tflops": fl / m / 1e9, "t_min_ms": tmin, "frac_of_t_min": tmin / m}
            log(f"[sweep] nq={nq:5d} scan {m:8.3f} ms  algorithmic HBM {corpus_bytes / m / 1e6:8.1f} GB/s "
                f"({corpus_bytes / m / 1e6 / pk['hbm_gbs']:.3f} of peak)  {fl / m / 1e9:8.1f} TFLOP/s  t_min/t {tmin / m:.2f}")
        line["roofline"]["nq_sweep_dense_top10"] = sweep

    # ---- the faiss-shaped module's DEFAULT storage (plain fp32 rows, TF32 filter) on the same workload ----------
    if world == 1 and not args.no_f32 and args.storage == "f32+bf16":
        ixf = (hf.IndexFlatL2 if metric_l2 else hf.IndexFlatIP)(args.dim, device=local, storage="f32")
        synth.dense_corpus_into(ixf, n_local, args.dim, dev, seed=synth.DENSE_SEED + rank)
        engf = HybridRetriever(ixf, bm)
        for _ in range(3):
            of = engf.retrieve(q_dev, (qi_dev, qt_dev), args.topk)
        msf, _ = event_ms(lambda: engf.retrieve(q_dev, (qi_dev, qt_dev), args.topk), 5)
        line["roofline"]["f32_default"] = {
            "storage": "f32", "storage_bytes": int(n_local * args.dim * 4), "ms_per_step": msf,
            "value": args.nq / (msf / 1e3), "unit": UNIT, "scan_kernel_ms": ixf.stats()["scan_ms"],
            "answers_equal_f32_bf16_storage": bool(torch.equal(of[1], out[1]) and torch.equal(of[0], out[0])),
            "note": "same corpus, same queries, HR_STORAGE=f32 (what IndexFlatL2(d) builds by default): kind::tf32 MMA at "
                    "half the bf16 rate, twice the streamed bytes, 2/3 of the memory"}
        log(f"[f32] default storage: {msf:.2f} ms/step = {args.nq / (msf / 1e3):.0f} QPS, answers equal: "
            f"{line['roofline']['f32_default']['answers_equal_f32_bf16_storage']}")
        del engf, ixf
        torch.cuda.empty_cache()

    # ---- C2 (N >= 8): 100M x 1024 bf16 rows over the ranks, hybrid top-100 ---------------------------------
    if world >= args.c2_min_gpus and not args.no_c2:
        del engine, ix, bm
        torch.cuda.empty_cache()
        t0 = time.time()
        lo2, hi2 = shard_bounds(args.c2_rows, world, rank)
        n2 = hi2 - lo2
        ix2 = (hf.IndexFlatL2 if metric_l2 else hf.IndexFlatIP)(args.dim, device=local, storage="bf16")
        ix2.set_id_base(lo2)
        planted2 = synth.dense_corpus_into(ix2, n2, args.dim, dev, seed=synth.DENSE_SEED + 100 + rank, keep_rows=4096)
        dist.broadcast(planted2, src=0)
        q2 = synth.dense_queries_torch(planted2, args.nq, args.dim, dev)
        ip2, pd2, tf2, dl2 = synth.sparse_corpus_csr_torch(n2, args.vocab, dev, seed=synth.SPARSE_SEED + 100 + rank)
        dfg2, ng2, avg2 = global_bm25_stats((ip2[1:] - ip2[:-1]).contiguous(), n2, int(dl2.sum().item()))
        bm2 = pbm25.BM25Index.from_csr(ip2, pd2, tf2, dl2, args.vocab, n_docs_global=ng2, avgdl_global=avg2,
                                       df_global=dfg2, device=local)
        bm2.set_id_base(lo2)
        del ip2, pd2, tf2, dl2
        torch.cuda.empty_cache()
        eng2 = ShardedRetriever(ix2, bm2)
        ix = ix2    # timed() reads the scan statistics of `ix`
        step2 = lambda: eng2.retrieve(q2, (qi_dev, qt_dev), 100, k_c=100)  # noqa: E731
        for _ in range(3):
            o2 = step2()
        ms2, scan2, fl2 = timed(step2, 5)
        ms2 /= 5
        # parity: the filter + re-score answer of 16 queries equals the exhaustive exact scan of the bf16 rows
        sub = q2[:16].contiguous()
        Da, Ia = ix2.search(sub, 100)
        ix2.set_mode("exact")
        De, Ie = ix2.search(sub, 100)
        ix2.set_mode("auto")
        okc = torch.tensor([int(torch.equal(Ia, Ie) and torch.equal(Da, De))], device=dev)
        dist.all_reduce(okc, op=dist.ReduceOp.MIN)
        fl = 2.0 * args.nq * n2 * args.dim
        line["c2"] = {"workload": f"BASELINE configs[2]: {args.c2_rows}x{args.dim} bf16 rows (fp32 accumulate) + BM25 over {args.c2_rows} "
                                  f"chunks, row-sharded over {world} GPUs, batch {args.nq}, hybrid top-100 (k_c = 100), NCCL all-gather merge",
                      "rows": args.c2_rows, "rows_per_rank": n2, "ms_per_step": ms2, "qps": args.nq / (ms2 / 1e3),
                      "scan_ms_rank0": scan2, "scan_tflops_rank0": fl / (scan2 / 1e3) / 1e12,
                      "scan_frac_of_sustained_bf16_peak": fl / (scan2 / 1e3) / 1e12 / pk["bf16_tflops_sustained"],
                      "allgather_bytes_per_rank": int(args.nq * 100 * 24 + 16), "allgather_bytes_total": int(world * (args.nq * 100 * 24 + 16)),
                      "fallback_queries_in_timed_region": int(fl2[0]), "setup_s": time.time() - t0,
                      "parity_check": {"dense_top100_equal_exact_scan_16_queries_every_rank": bool(okc.item())}}
        log(f"[c2] {args.c2_rows} rows over {world} GPUs: {ms2:.2f} ms/step, scan {scan2:.2f} ms = {fl / (scan2 / 1e3) / 1e12:.0f} TFLOP/s per GPU")

    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(args, steps=2, warmup=1)   # ~15 s of CPU work on 16 cores
        print(json.dumps(line), file=_json_out, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    # stdout carries exactly ONE JSON line: anything a library prints there (NCCL's version banner, ...) is
    # sent to stderr instead; the JSON goes to a private duplicate of the original stdout
    sys.stdout.flush()
    _json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
