#!/usr/bin/env python
"""Bring-up diagnostics for the CUDA path (run on the GPU box).  Each stage runs in its own
subprocess under a timeout so a hung kernel in one stage does not hide the others' output."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def stage_exact():
    import numpy as np
    import intool_rag_b200  # noqa
    from intool_rag_b200 import faiss as hf, synth
    from oracle import flat
    for metric in ("ip", "l2"):
        n, d, nq = 5000, 64, 37
        x = synth.dense_corpus_np(n, d)
        q = synth.dense_queries_np(x, nq)
        ix = hf.IndexFlatIP(d) if metric == "ip" else hf.IndexFlatL2(d)
        ix.set_mode("exact")
        ix.add(x)
        D, I = ix.search(q, 10)
        o = flat.IndexFlatIP(d) if metric == "ip" else flat.IndexFlatL2(d)
        o.add(x)
        Dr, Ir = o.search(q, 10, precision="f64")
        print(f"exact/{metric}: ids equal {(I == Ir).mean():.4f} max|dD| {np.abs(D - Dr).max():.2e} stats {ix.stats()}")


def stage_filter(d=64, n=5000, nq=37, metric="ip", storage="f32"):
    import ctypes as C
    import numpy as np
    import intool_rag_b200  # noqa
    from intool_rag_b200 import faiss as hf, synth, _lib
    from oracle import flat
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, nq)
    ix = hf.IndexFlat(d, hf.METRIC_INNER_PRODUCT if metric == "ip" else hf.METRIC_L2, storage=storage)
    ix.add(x)
    D, I = ix.search(q, 10)
    st = ix.stats()
    o = flat.IndexFlatIP(d) if metric == "ip" else flat.IndexFlatL2(d)
    xs = x
    if storage == "bf16":
        import torch
        xs = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()
    o.add(xs)
    Dr, Ir = o.search(q, 10, precision="f64")
    print(f"filter/{metric}/{storage} n={n} d={d} nq={nq}: ids equal {(I == Ir).mean():.4f} "
          f"max|dD| {np.abs(D - Dr).max():.2e} stats {st}")
    G, KL = st["grid"], st["list_len"]
    lists = np.zeros((G, nq, KL, 2), dtype=np.uint32)
    cnts = np.zeros((G, nq), dtype=np.int32)
    tau = np.zeros(nq, dtype=np.uint32)
    tprime = np.zeros(nq, dtype=np.float32)
    rc = _lib.lib().hr_index_debug_dump(ix._h, nq, lists.ctypes.data, cnts.ctypes.data, tau.ctypes.data, None,
                                        tprime.ctypes.data)
    if rc != 0:
        print("  debug_dump:", _lib.last_error())
        return
    sc = lists[..., 0].view(np.float32)
    rows = lists[..., 1]
    errs = []
    nbad_row = 0
    for g in range(G):
        for qi in range(nq):
            c = cnts[g, qi]
            if c == 0:
                continue
            r = rows[g, qi, :c].astype(np.int64)
            if (r >= n).any():
                nbad_row += int((r >= n).sum())
                r = np.minimum(r, n - 1)
            true = xs[r].astype(np.float64) @ q[qi].astype(np.float64)
            if metric == "l2":
                true = true - 0.5 * (xs[r].astype(np.float64) ** 2).sum(1)
            errs.append(np.abs(sc[g, qi, :c] - true).max())
    errs = np.array(errs) if errs else np.zeros(1)
    print(f"  lists: total entries {int(cnts.sum())}, rows out of range {nbad_row}, "
          f"approx-vs-true score err max {errs.max():.3e} median {np.median(errs):.3e}; tprime[:4] {tprime[:4]}")
    print(f"  sample q0 list g0: {[(float(a), int(b)) for a, b in zip(sc[0, 0, :4], rows[0, 0, :4])]}")


STAGES = {
    "exact": stage_exact,
    "filter_small": lambda: stage_filter(),
    "filter_l2": lambda: stage_filter(metric="l2"),
    "filter_d1024": lambda: stage_filter(d=1024, n=40000, nq=300),
    "filter_bf16": lambda: stage_filter(d=256, n=20000, nq=130, storage="bf16"),
}

if __name__ == "__main__":
    if len(sys.argv) > 1:
        STAGES[sys.argv[1]]()
        sys.exit(0)
    for name in STAGES:
        print(f"=== {name}", flush=True)
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), name], timeout=240,
                               capture_output=True, text=True)
            print(r.stdout[-6000:], r.stderr[-3000:], f"[exit {r.returncode}]", flush=True)
        except subprocess.TimeoutExpired as e:
            print(f"TIMEOUT in {name}: {(e.stdout or b'')[-2000:]} {(e.stderr or b'')[-2000:]}", flush=True)
