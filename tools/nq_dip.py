#!/usr/bin/env python
"""Why the dense scan sits at 0.8-0.9 of t_min for nq = 256..512: SM clock and board power sampled
(nvidia-smi, 20 ms) while the scan runs back to back for ~2 s per batch size.  Prints one row per nq:
scan kernel ms, t_hbm, t_tensor (sustained cuBLAS peak), SM MHz, W, and TFLOP/s per GHz (the tensor pipe's work per
clock: constant when the pipe is saturated, so the time then follows the clock the power cap allows).
Run on the GPU box: python tools/nq_dip.py [rows]"""
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import intool_rag_b200  # noqa: F401,E402
from intool_rag_b200 import faiss as hf, synth  # noqa: E402

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = 1024
dev = torch.device("cuda", 0)
pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else \
    {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "bf16_tflops": 1590.0}
ix = hf.IndexFlatIP(d, storage="f32+bf16")
planted = synth.dense_corpus_into(ix, rows, d, dev, keep_rows=4096)
q_all = synth.dense_queries_torch(planted, 1024, d, dev)


def sample(fn, seconds=2.0):
    path = tempfile.mktemp(suffix=".csv")
    p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw,clocks.mem", "--format=csv,noheader,nounits",
                          "-lms", "20"], stdout=open(path, "w"), stderr=subprocess.DEVNULL)
    ms = []
    t0 = time.time()
    while time.time() - t0 < seconds:
        fn()
        ms.append(ix.stats()["scan_ms"])
    p.terminate()
    p.wait()
    rows_ = [l.split(",") for l in open(path) if l.count(",") == 2]
    os.unlink(path)
    clk = [float(r[0]) for r in rows_[5:]] or [0.0]
    pw = [float(r[1]) for r in rows_[5:]] or [0.0]
    half = ms[len(ms) // 2:]
    return float(np.median(half)), float(np.median(clk)), float(np.median(pw)), len(ms)


print("| nq | scan ms | t_hbm ms | t_tensor ms (sustained peak) | t_min/t | SM MHz | board W | TFLOP/s | TFLOP/s per GHz | GB/s |")
print("|---|---|---|---|---|---|---|---|---|---|")
for nq in (1, 32, 128, 192, 256, 384, 512, 768, 1024):
    q = q_all[:nq].contiguous()
    for _ in range(3):
        ix.search(q, 10)
    ms, clk, pw, n = sample(lambda: ix.search(q, 10))
    fl = 2.0 * nq * rows * d
    by = rows * d * 2.0
    th, tt = by / (pk["hbm_gbs"] * 1e9) * 1e3, fl / (pk["bf16_tflops_sustained"] * 1e12) * 1e3
    tf = fl / ms / 1e9
    print(f"| {nq} | {ms:.3f} | {th:.3f} | {tt:.3f} | {max(th, tt) / ms:.2f} | {clk:.0f} | {pw:.0f} | {tf:.0f} | "
          f"{tf / max(clk, 1) * 1000:.0f} | {by / ms / 1e6:.0f} |", flush=True)
    time.sleep(1.0)
