import sys, time, numpy as np
sys.path.insert(0, "/root/repo")
import importlib
synth = importlib.import_module("intool-rag_b200.synth") if False else None
import importlib.util
spec = importlib.util.spec_from_file_location("synth", "/root/repo/intool-rag_b200/synth.py"); synth = importlib.util.module_from_spec(spec); spec.loader.exec_module(synth)
from oracle import bm25 as ob
N = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
V = 30000; KC = 50; SLICE = 3072
t0 = time.time()
t, dd, dl = synth.sparse_corpus_np(N, V)
c = ob.BM25Corpus.from_token_matrix(t, dd, dl, V); del t, dd
print("built", time.time() - t0, "nnz", len(c.post_doc))
maximp = np.zeros(V); 
for tt in range(V):
    a, b = c.indptr[tt], c.indptr[tt+1]
    if b > a: maximp[tt] = c.impact[a:b].max()
qs = synth.sparse_queries_np(64, V)
tot_post = 0; ess_post = 0; hot_tot = 0; lookups = 0; ess_post_early = 0; hot_early = 0; look_early = 0
for q in qs:
    w = np.array([c.idf[x] for x in q]); order = np.argsort(-w, kind="stable"); q = [q[i] for i in order]; w = w[order]
    ub = w * np.array([maximp[x] for x in q]) * 1.00001
    acc = c.scores(q); 
    top = np.sort(acc)[::-1]
    tau_final = top[KC-1]
    # early tau: 50th best of the first 1/55 of the docs
    tau_early = np.sort(acc[:N//55])[::-1][KC-1]
    df = np.array([c.indptr[x+1]-c.indptr[x] for x in q])
    tot_post += df.sum()
    for tau, tag in ((tau_final, "final"), (tau_early, "early")):
        # choose m: cost model  sweep cost = essential postings ; lookup cost = hot_docs_est * (nt-m) * CL
        best = None
        nt = len(q)
        for m in range(nt, 0, -1):
            rest = ub[m:].sum()
            if rest >= tau: break
            thr = tau - rest
            # exact hot count: docs with partial >= thr
            part = np.zeros(N)
            for x, wx in zip(q[:m], w[:m]):
                a, b = c.indptr[x], c.indptr[x+1]
                part[c.post_doc[a:b]] += wx * c.impact[a:b]
            hot = int((part >= thr).sum())
            cost = df[:m].sum() + hot * (nt - m) * 40.0
            if best is None or cost < best[0]: best = (cost, m, hot)
        cost, m, hot = best
        if tag == "final":
            ess_post += df[:m].sum(); hot_tot += hot; lookups += hot * (nt - m)
        else:
            ess_post_early += df[:m].sum(); hot_early += hot; look_early += hot * (nt - m)
    print(len(q), "df%", np.round(100*df/N,2), "ub", np.round(ub,1), "tau", round(tau_final,2), round(tau_early,2), "m", m, "hot", hot)
print("postings", tot_post, "essential(final tau)", ess_post/tot_post, "hot/query", hot_tot/len(qs), "lookups/query", lookups/len(qs))
print("essential(early tau)", ess_post_early/tot_post, "hot/query", hot_early/len(qs), "lookups/query", look_early/len(qs))
