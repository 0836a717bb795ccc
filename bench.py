#!/usr/bin/env python
"""bench.py — the hybrid retrieval hot path on B200 (contract in the task statement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A "step" is one pass of the hot path over one batch of synthetic queries:
retrieve(query_embeddings, query_tokens, top_k=10) = dense flat scan (tcgen05 filter + exact fp32
re-score) + BM25 posting scatter + weighted fusion + top-k.  Workload at N=1 is BASELINE.json
configs[1]: 10M x 1024-d fp32 flat index + BM25 over 10M chunks (30k-term Zipf vocabulary),
batch of 1024 queries, hybrid top-10.  At N>1 the SAME 10M-row corpus is row-sharded over the N
ranks (strong scaling: the metric is QPS on a fixed corpus), candidates travel through one NCCL
all-gather, every rank merges + fuses.

`value`  : whole-job QPS with queries already resident in HBM (CUDA events, max over ranks).
`e2e`    : same metric through the public API with HOST (pinned) query buffers and the results read
           back to the host inside the timed region.
`roofline`: the dominant kernel (scan_tc_kernel) timed live with CUDA events on its launch stream.
`cpu_baseline`: the CPU oracle (numpy/OpenBLAS sgemm restatement of faiss IndexFlat + BM25 + fusion;
           faiss-cpu 1.7.4 is not installable offline) on a bounded row sample, scaled to the full corpus.
`--impl reference`: that same CPU path alone (rank 0 only under torchrun), no GPU code involved.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

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
    ap.add_argument("--cpu-sample-rows", type=int, default=int(os.getenv("HR_BENCH_CPU_ROWS", 200_000)))
    ap.add_argument("--storage", default=os.getenv("HR_BENCH_STORAGE", "f32+bf16"), choices=["f32", "f32+bf16", "bf16"],
                    help="f32: fp32 rows, TF32 filter; f32+bf16: fp32 rows + bf16 shadow for the filter (same answers); "
                         "bf16: bf16 rows")
    ap.add_argument("--index-metric", default="ip", choices=["ip", "l2"],
                    help="ip: IndexFlatIP (BASELINE.json); l2: IndexFlatL2, what the reference builds "
                         "(rag/storage/faiss_index.py:123); same ranking on unit-norm rows")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--sweep", action="store_true", help="(kept for compatibility: the nq sweep is part of the line)")
    ap.add_argument("--no-sweep", action="store_true", help="skip the nq sweep of the dense scan")
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
    """The reference's CPU path for this workload on the host cores: faiss-equivalent exhaustive
    search (OpenBLAS sgemm in 1024-row blocks + top-k, the algorithm faiss 1.7.4 uses for nq >= 20),
    BM25 over a CSR index, weighted fusion — the oracle package.  Bounded sample: the first
    `cpu_sample_rows` rows of a corpus with the same distributions; time scales linearly in rows, so
    QPS(full) = QPS(sample) * sample_rows / rows."""
    import intool_rag_b200  # noqa: F401  (generators only; no CUDA code runs in this arm)
    from intool_rag_b200 import synth
    from oracle import bm25 as obm25
    from oracle import flat, hybrid
    n = min(args.cpu_sample_rows, args.rows)
    x = synth.dense_corpus_np(n, args.dim)
    q = synth.dense_queries_np(x, args.nq)
    t, dd, dl = synth.sparse_corpus_np(n, args.vocab)
    qs = synth.sparse_queries_np(args.nq, args.vocab)
    ix = flat.IndexFlatIP(args.dim) if args.index_metric == "ip" else flat.IndexFlatL2(args.dim)
    ix.add(x)
    corpus = obm25.BM25Corpus.from_token_matrix(t, dd, dl, args.vocab)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        hybrid.retrieve(ix, corpus, q, qs, args.topk)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    tot = float(np.sum(times))
    qps_sample = args.nq * len(times) / tot
    qps_full = qps_sample * n / args.rows
    try:
        import torch
        threads = torch.get_num_threads()
    except Exception:
        threads = os.cpu_count()
    return {"value": qps_full, "unit": UNIT, "cores": int(os.cpu_count() or 1), "blas_threads": int(threads),
            "kind": "port",
            "sample": f"{args.nq} queries x first {n} of {args.rows} rows (dense sgemm+top-k, BM25, fusion); "
                      f"QPS scaled by {n}/{args.rows}; {len(times)} steps, {tot / len(times):.2f} s/step; "
                      "faiss-cpu 1.7.4 not installable offline -> numpy/OpenBLAS restatement (oracle/)",
            "ms_per_step_sample": 1e3 * tot / len(times)}


def run_reference(args):
    rank = int(os.getenv("RANK", "0"))
    if rank != 0:
        return
    steps, warmup = max(1, min(args.steps, 3)), min(args.warmup, 1)
    cb = cpu_reference(args, steps, warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": warmup, "ms_per_step": cb["ms_per_step_sample"] * args.rows / min(args.cpu_sample_rows, args.rows),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, args.gpus), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), file=_json_out, flush=True)


def workload_config(args, n_gpus):
    return {"workload": f"BASELINE configs[1]: {args.rows}x{args.dim} fp32 flat index (IndexFlat{args.index_metric.upper()} semantics, storage {args.storage}) + BM25 over "
                        f"{args.rows} chunks ({args.vocab}-term Zipf vocabulary), batch {args.nq} queries, hybrid top-{args.topk}, "
                        "weighted fusion 0.7/0.3, candidate depth 50",
            "rows": args.rows, "dim": args.dim, "nq": args.nq, "top_k": args.topk, "vocab": args.vocab, "storage": args.storage,
            "sharding": f"row-sharded over {n_gpus} GPU(s), one NCCL all-gather of k_c candidates" if n_gpus > 1 else "single GPU",
            "l2_hygiene": "inputs larger than L2 (40.96 GB corpus + 8-11 GB postings streamed per step vs 126 MB L2)"}


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
        sm, mx, reasons = [], [], set()
        for line in open(self.path):
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out = {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                   "samples": len(sm)}
        try:
            os.unlink(self.path)
        except OSError:
            pass
        return out


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

    # ---- build this rank's shard on device -------------------------------------------------------
    t0 = time.time()
    lo, hi = shard_bounds(args.rows, world, rank)
    n_local = hi - lo
    ix = (hf.IndexFlatIP if args.index_metric == "ip" else hf.IndexFlatL2)(args.dim, device=local, storage=args.storage)
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
    del indptr, post_doc, post_tf, doc_len
    torch.cuda.empty_cache()
    qi_np, qt_np = synth.sparse_queries_csr(args.nq, args.vocab)
    qi_dev = torch.from_numpy(qi_np).to(dev)
    qt_dev = torch.from_numpy(qt_np).to(dev)
    log(f"[bench] BM25 shard: {nnz_local} postings; total setup {time.time() - t0:.1f}s")

    if world == 1:
        engine = HybridRetriever(ix, bm)
        step_dev = lambda: engine.retrieve(q_dev, (qi_dev, qt_dev), args.topk)  # noqa: E731
    else:
        engine = ShardedRetriever(ix, bm)
        step_dev = lambda: engine.retrieve(q_dev, (qi_dev, qt_dev), args.topk)  # noqa: E731

    # pinned host copies for the end-to-end arm
    q_host = torch.empty((args.nq, args.dim), dtype=torch.float32, pin_memory=True)
    q_host.copy_(q_dev)
    qi_host = torch.from_numpy(qi_np).pin_memory()
    qt_host = torch.from_numpy(qt_np).pin_memory()
    h2d = q_host.numel() * 4 + qi_host.numel() * 4 + qt_host.numel() * 4
    d2h = args.nq * args.topk * 12

    def step_e2e():
        if world == 1:
            S, I = engine.retrieve(q_host.numpy(), (qi_host.numpy(), qt_host.numpy()), args.topk)
            return S, I
        qd = q_host.to(dev, non_blocking=True)
        qid = qi_host.to(dev, non_blocking=True)
        qtd = qt_host.to(dev, non_blocking=True)
        S, I = engine.retrieve(qd, (qid, qtd), args.topk)
        return S.cpu(), I.cpu()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize; CUDA events on the launch stream; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        scan_ms, flagged, deeper = 0.0, 0, 0
        e0.record()
        for _ in range(steps):
            fn()
            st = ix.stats()
            scan_ms += st["scan_ms"]
            flagged += st["flagged"]
            deeper += st["deeper"]
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        cnt = torch.tensor([flagged, deeper], device=dev, dtype=torch.int64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(cnt)     # fallbacks on ANY rank stall every rank at the all-gather
        return float(ms.item()), scan_ms / steps, (int(cnt[0].item()), int(cnt[1].item()))

    for _ in range(max(args.warmup, 3)):
        out = step_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = _lib.launch_count()
    total_ms, scan_ms, flagged = timed(step_dev, args.steps)
    launches = _lib.launch_count() - l0
    for _ in range(2):
        step_e2e()
    e2e_ms, _, _ = timed(step_e2e, args.steps)
    clocks = sampler.stop() if rank == 0 else {}

    # ---- sanity inside the bench: the timed answer equals the exact SIMT scan on a few queries ------
    S, I = out
    check = {}
    if world == 1:
        sub = q_dev[:16].contiguous()
        D_auto, I_auto = ix.search(sub, 50)
        ix.set_mode("exact")
        D_ex, I_ex = ix.search(sub, 50)
        ix.set_mode("auto")
        check = {"dense_top50_ids_equal_exact_scan": bool(torch.equal(I_auto, I_ex)),
                 "dense_scores_equal_exact_scan": bool(torch.equal(D_auto, D_ex)), "queries_checked": 16}
    # ---- BM25 kernel alone (for the second roofline) ---------------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _, _, touched = bm.search((qi_dev, qt_dev), 50, return_postings=True)
    e1.record()
    torch.cuda.synchronize()
    bm_ms = e0.elapsed_time(e1)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
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
            "kind::f16 (bf16 operands, fp32 accumulate in TMEM) over the bf16 rows; answers are exact fp32 after the re-score")
    # dram__bytes_read.sum + dram__bytes_write.sum of the dominant kernel from the committed ncu capture of this
    # exact default workload (profiles/r1_scan_bm25_ncu_full.md); other workloads: not captured
    traffic = None
    if (world, args.rows, args.dim, args.nq, args.storage) == (1, 10_000_000, 1024, 1024, "f32+bf16"):
        traffic = 20.791e9 + 0.039e9
    roofline = {"kernel": kname, "bound": bound, "achieved": achieved, "peak": peak, "unit": runit,
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "bytes per launch (ncu, profiles/r1_scan_bm25_ncu_full.md)",
                "peak_source": pk["src"],
                "kernel_ms": scan_ms, "share_of_step": scan_ms / ms_per_step,
                "algorithmic_flops_per_launch": flops, "algorithmic_bytes_per_launch": corpus_bytes,
                "hbm_gbs_algorithmic": corpus_bytes / scan_s / 1e9,
                "hbm_frac_algorithmic": corpus_bytes / scan_s / 1e9 / pk["hbm_gbs"],
                "note": note,
                "bm25": {"kernel": "bm25_slice_kernel (+ plan kernels)", "bound": "hbm", "postings": int(touched), "ms": bm_ms,
                         "achieved": touched * 8 / (bm_ms / 1e3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": touched * 8 / (bm_ms / 1e3) / 1e9 / pk["hbm_gbs"]}}
    line = {"metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(args, world),
            "clocks": clocks,
            "e2e": {"value": e2e_qps, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches), "roofline": roofline,
            "fallback_queries_in_timed_region": int(flagged[0]),
            "second_stage_rescore_queries_in_timed_region": int(flagged[1]), "parity_check": check}
    if world == 1 and not args.no_sweep:
        # the HBM-bound regime of the same scan kernel family: top-10 dense search for growing batches (scan kernel
        # time from the library's CUDA events; bytes = the rows the filter streams, once per batch)
        sweep = {}
        for nq in (1, 8, 32, 128, 256, 512, 1024):
            if nq > args.nq:
                break
            qq = q_dev[:nq].contiguous()
            for _ in range(3):
                ix.search(qq, 10)
            ms = []
            for _ in range(5):
                ix.search(qq, 10)
                ms.append(ix.stats()["scan_ms"])
            m = float(np.median(ms))
            sweep[str(nq)] = {"scan_ms": m, "hbm_gbs_algorithmic": corpus_bytes / m / 1e6,
                              "hbm_frac": corpus_bytes / m / 1e6 / pk["hbm_gbs"],
                              "tflops": 2.0 * nq * n_local * args.dim / m / 1e9}
            log(f"[sweep] nq={nq:5d} scan {m:8.3f} ms  algorithmic HBM {corpus_bytes / m / 1e6:8.1f} GB/s "
                f"({corpus_bytes / m / 1e6 / pk['hbm_gbs']:.3f} of peak)  {2.0 * nq * n_local * args.dim / m / 1e9:8.1f} TFLOP/s")
        line["roofline"]["nq_sweep_dense_top10"] = sweep
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference(args, steps=1, warmup=1)
    print(json.dumps(line), file=_json_out, flush=True)
    if world > 1:
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
