#!/usr/bin/env python
"""Generates the committed golden fixtures in tests/golden/.

Run in the BUILD container only (needs /root/reference, read-only):
    python tests/golden/make_golden.py

Two kinds of fixtures:
 1. ``ref_wrapper.json``/``.npz`` — outputs of the reference's UNMODIFIED storage wrapper
    (/root/reference/rag/storage/faiss_index.py: create_faiss_index, save_faiss_index,
    FAISSIndexReader.search, search_faiss_by_vector, and the page ranking of
    /root/reference/rag/query/page_retriever.py:145-236) executed on top of the numpy oracle
    registered as ``sys.modules["faiss"]`` (faiss-cpu itself is not installable offline,
    SURVEY.md §8c).  They pin this repo's storage mirror and the oracle's file format
    against what the reference code actually does with a faiss-shaped module.
 2. ``kat_*.json`` — hand-computable known answers (closed form) for flat search, BM25 and
    fusion; ``synthetic_small.npz`` — seeded inputs + fp64-oracle outputs for GPU parity.
"""
import asyncio
import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import flat, bm25, fusion, hybrid  # noqa: E402
import intool_rag_b200  # noqa: E402,F401
from intool_rag_b200 import synth  # noqa: E402


def ref_wrapper_fixture():
    ref = "/root/reference"
    if not os.path.isdir(ref):
        print("reference not mounted; skipping ref_wrapper fixture")
        return
    tmp = tempfile.mkdtemp(prefix="golden_ref_")
    os.environ["STORAGE_DIR"] = os.path.join(tmp, "storages")
    os.environ["CACHE_DIR"] = os.path.join(tmp, "cache")
    os.environ["LOG_LEVEL"] = "WARNING"
    cwd = os.getcwd()
    os.chdir(tmp)
    sys.modules["faiss"] = flat  # the oracle stands in for faiss-cpu
    sys.path.insert(0, ref)
    try:
        from rag.storage import faiss_index as ref_fi
        from rag.query.page_retriever import PageLevelRetriever, RetrievedChunk
        assert ref_fi.HAS_FAISS
        rng = np.random.default_rng(20261018)
        n, d = 37, 24
        x = rng.standard_normal((n, d)).astype(np.float32)
        x /= np.linalg.norm(x, axis=1, keepdims=True)
        x[5] = x[3]  # duplicate vector -> exact tie (reference quirk)
        queries = np.stack([x[3] * 0.9 + 0.1 * x[7], x[11], -x[2], rng.standard_normal(d).astype(np.float32)])
        queries = (queries / np.linalg.norm(queries, axis=1, keepdims=True)).astype(np.float32)
        index = ref_fi.create_faiss_index([list(map(float, r)) for r in x])
        doc_id = "docA"
        index_path = os.path.join(os.environ["STORAGE_DIR"], f"{doc_id}_faiss.index")
        ref_fi.save_faiss_index(index, index_path)
        raw = open(index_path, "rb").read()
        chunks = [{"chunk_id": f"c_{i // 4}_{i % 4:02d}", "page": i // 4 + 1, "text": f"chunk text {i}",
                   "chunk_index": i} for i in range(n)]
        with open(os.path.join(os.environ["STORAGE_DIR"], f"{doc_id}_chunks.json"), "w") as f:
            json.dump({"total": n, "chunks": chunks}, f)
        reader = ref_fi.FAISSIndexReader(index_path)
        out = {"n": n, "d": d, "doc_id": doc_id, "index_sha256": hashlib.sha256(raw).hexdigest(),
               "index_header_hex": raw[:45].hex(), "index_nbytes": len(raw),
               "reader_dimension": reader.get_dimension(), "reader_size": reader.get_size(),
               "reader_search": {}, "search_by_vector": {}, "page_ranking": {}}
        for qi, q in enumerate(queries):
            for k in (1, 5, 10, 50):
                res = reader.search(list(map(float, q)), top_k=k)
                out["reader_search"][f"q{qi}_k{k}"] = [[int(i), float(s)] for i, s in res]
            for limit in (7, 50):
                res = asyncio.run(ref_fi.search_faiss_by_vector(list(map(float, q)), limit=limit))
                out["search_by_vector"][f"q{qi}_l{limit}"] = res
            res = asyncio.run(ref_fi.search_faiss_by_vector(list(map(float, q)), limit=20))
            pr = PageLevelRetriever(top_chunks=20, top_pages=3)
            rc = [RetrievedChunk(chunk_id=r["chunk_id"], text=r["text"], score=r["score"], page=r["page"],
                                 metadata={"chapter": r.get("chapter"), "section": r.get("section")})
                  for r in res]
            groups = pr.group_chunks_by_page(rc)
            ranked = pr.rank_pages(groups)
            top = pr.select_top_pages(ranked)
            out["page_ranking"][f"q{qi}"] = [[int(p.page), float(p.score), len(p.chunks)] for p in top]
        np.savez_compressed(os.path.join(HERE, "ref_wrapper.npz"), x=x, queries=queries,
                            index_bytes=np.frombuffer(raw, dtype=np.uint8))
        with open(os.path.join(HERE, "ref_wrapper.json"), "w") as f:
            json.dump(out, f, indent=1)
        print("wrote ref_wrapper.{npz,json}")
    finally:
        os.chdir(cwd)
        sys.modules.pop("faiss", None)
        sys.path.remove(ref)


def kat_fixtures():
    # ---- flat index, d=4, hand-computed ----
    x = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1], [1, 0, 0, 0],
                  [-1, 0, 0, 0], [0.6, 0.8, 0, 0]], dtype=np.float32)
    q = np.array([[1, 0, 0, 0], [0, 0.6, 0.8, 0]], dtype=np.float32)
    kat = {
        "x": x.tolist(), "q": q.tolist(),
        # q0: |x-q|^2 = 0 (rows 0,4 tie -> id asc), 0.8 (row 6), 2 (rows 1,2,3), 4 (row 5)
        "l2_k7_ids": [[0, 4, 6, 1, 2, 3, 5], [6, 1, 2, 0, 3, 4, 5]],
        "l2_k7_dist": [[0.0, 0.0, 0.8, 2.0, 2.0, 2.0, 4.0],
                       [1.04, 0.8, 0.4, 2.0, 2.0, 2.0, 2.0]],
        "ip_k7_ids": [[0, 4, 6, 1, 2, 3, 5], [2, 1, 6, 0, 3, 4, 5]],
        "ip_k7_score": [[1.0, 1.0, 0.6, 0.0, 0.0, 0.0, -1.0], [0.8, 0.6, 0.48, 0.0, 0.0, 0.0, 0.0]],
        # reference transform clamp(1 - d/2, 0, 1): d=0 ->1, 0.8->0.6, 2->0, 4->0 (clamped from -1)
        "ref_score_q0": [1.0, 1.0, 0.6, 0.0, 0.0, 0.0, 0.0],
        "k9_pad_ids_q0": [0, 4, 6, 1, 2, 3, 5, -1, -1],
    }
    # fix q1 L2 order by actually sorting the hand distances (1.04 is row 6: |(0.6,0.8)-(0,0.6,0.8)|^2)
    d_q1 = [2.0, 0.8, 0.4, 2.0, 2.0, 2.0, 0.36 + 0.04 + 0.64]
    order = sorted(range(7), key=lambda i: (d_q1[i], i))
    kat["l2_k7_ids"][1] = order
    kat["l2_k7_dist"][1] = [d_q1[i] for i in order]
    with open(os.path.join(HERE, "kat_flat.json"), "w") as f:
        json.dump(kat, f, indent=1)

    # ---- BM25, 3 docs, hand-computed (item 5) ----
    docs = [[0, 1, 1, 2], [1, 3], [0, 0, 0, 2, 2, 4, 4, 4]]
    V = 5
    N, avgdl = 3, (4 + 2 + 8) / 3.0
    k1, b = 1.5, 0.75

    def idf_l(df):
        return float(np.log((N - df + 0.5) / (df + 0.5) + 1.0))

    def imp(tf, dl):
        return tf * (k1 + 1) / (tf + k1 * (1 - b + b * dl / avgdl))

    # query [1, 0]: df(1)=2, df(0)=2
    s0 = idf_l(2) * imp(2, 4) + idf_l(2) * imp(1, 4)
    s1 = idf_l(2) * imp(1, 2)
    s2 = idf_l(2) * imp(3, 8)
    # query [4, 4, 3]: term 4 twice (counts twice), df(4)=1, df(3)=1
    t0 = 0.0
    t1 = idf_l(1) * imp(1, 2)
    t2 = 2 * idf_l(1) * imp(3, 8)
    katb = {"docs": docs, "vocab": V, "k1": k1, "b": b, "queries": [[1, 0], [4, 4, 3], [9, -1]],
            "scores": [[s0, s1, s2], [t0, t1, t2], [0.0, 0.0, 0.0]]}
    with open(os.path.join(HERE, "kat_bm25.json"), "w") as f:
        json.dump(katb, f, indent=1)

    # ---- fusion on 2x5 lists with overlap (item 6) ----
    dense_ids = [[10, 11, 12, 13, 14]]
    dense_sim = [[0.9, 0.8, 0.7, 1.2, -0.1]]          # 1.2 clamps to 1, -0.1 clamps to 0
    bm_ids = [[12, 20, 10, 21, -1]]
    bm_s = [[8.0, 6.0, 4.0, 2.0, 0.0]]
    wv, wb = 0.7, 0.3
    weighted = {10: wv * 0.9 + wb * 0.5, 11: wv * 0.8, 12: wv * 0.7 + wb * 1.0, 13: wv * 1.0, 14: 0.0,
                20: wb * 0.75, 21: wb * 0.25}
    rrf = {10: 1 / 61 + 1 / 63, 11: 1 / 62, 12: 1 / 63 + 1 / 61, 13: 1 / 64, 14: 1 / 65, 20: 1 / 62, 21: 1 / 64}
    katf = {"dense_ids": dense_ids, "dense_sim": dense_sim, "bm25_ids": bm_ids, "bm25": bm_s,
            "w_vec": wv, "w_bm25": wb,
            "weighted": sorted([[i, s] for i, s in weighted.items()], key=lambda t: (-t[1], t[0])),
            "rrf": sorted([[i, s] for i, s in rrf.items()], key=lambda t: (-t[1], t[0]))}
    with open(os.path.join(HERE, "kat_fusion.json"), "w") as f:
        json.dump(katf, f, indent=1)
    print("wrote kat_{flat,bm25,fusion}.json")


def synthetic_small():
    """Seeded small hybrid case + fp64-oracle answers (GPU parity without recomputing the oracle)."""
    n, d, V, nq = 3000, 96, 400, 64
    x = synth.dense_corpus_np(n, d)
    x[100] = x[7]
    q = synth.dense_queries_np(x, nq)
    t, dd, dl = synth.sparse_corpus_np(n, V, mean_len=40.0)
    qt = synth.sparse_queries_np(nq, V, stop=8)
    out = {}
    for name, metric in (("ip", flat.METRIC_INNER_PRODUCT), ("l2", flat.METRIC_L2)):
        ix = flat.IndexFlat(d, metric)
        ix.add(x)
        D, I = ix.search(q, 50, precision="f64")
        out[f"{name}_D"], out[f"{name}_I"] = D, I
        corpus = bm25.BM25Corpus.from_token_matrix(t, dd, dl, V)
        for mode in ("weighted", "rrf"):
            fs, fi, parts = hybrid.retrieve(ix, corpus, q, qt, 10, mode=mode, precision="f64")
            out[f"{name}_{mode}_S"], out[f"{name}_{mode}_I"] = fs, fi
        out["bm25_S"], out["bm25_I"] = parts["bm25_S"], parts["bm25_I"]
    qi = np.zeros(nq + 1, np.int32)
    for i, ql in enumerate(qt):
        qi[i + 1] = qi[i] + len(ql)
    np.savez_compressed(os.path.join(HERE, "synthetic_small.npz"), n=n, d=d, V=V, nq=nq,
                        q_indptr=qi, q_terms=np.concatenate([np.asarray(a, np.int32) for a in qt]), **out)
    print("wrote synthetic_small.npz (inputs are regenerated from seeds by intool_rag_b200.synth)")


if __name__ == "__main__":
    ref_wrapper_fixture()
    kat_fixtures()
    synthetic_small()
