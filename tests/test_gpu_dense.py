"""GPU parity: the flat index (libhr_b200 through the faiss-shaped module / C ABI) vs the oracle."""
import json
import os

import numpy as np
import pytest

import intool_rag_b200  # noqa: F401
from intool_rag_b200 import faiss as hf
from intool_rag_b200 import synth
from oracle import flat
from helpers import assert_topk_matches, recall_at_k

pytestmark = pytest.mark.gpu
FLT_MAX = np.float32(3.4028234663852886e38)
TOL_F32 = 3e-6   # exact fp32 re-score vs fp64 truth on unit-norm data (measured ~5e-7), stated bound


def _mk(metric, d, storage="f32"):
    return hf.IndexFlatL2(d, storage=storage) if metric == "l2" else hf.IndexFlatIP(d, storage=storage)


def _oracle(metric, d):
    return flat.IndexFlatL2(d) if metric == "l2" else flat.IndexFlatIP(d)


@pytest.mark.parametrize("mode", ["auto", "exact"])
def test_kat_flat(gpu, golden_dir, mode):
    k = json.load(open(os.path.join(golden_dir, "kat_flat.json")))
    x, q = np.array(k["x"], np.float32), np.array(k["q"], np.float32)
    for metric, ids, sc in (("l2", "l2_k7_ids", "l2_k7_dist"), ("ip", "ip_k7_ids", "ip_k7_score")):
        ix = _mk(metric, 4)
        ix.set_mode(mode)
        ix.add(x)
        assert (ix.d, ix.ntotal, ix.is_trained) == (4, 7, True)
        D, I = ix.search(q, 7)
        assert I.dtype == np.int64 and D.dtype == np.float32
        assert I.tolist() == k[ids]
        np.testing.assert_allclose(D, np.array(k[sc], np.float32), atol=1e-6)
    ix = _mk("l2", 4)
    ix.set_mode(mode)
    ix.add(x)
    D, I = ix.search(q[:1], 9)
    assert I[0].tolist() == k["k9_pad_ids_q0"] and (D[0, 7:] == FLT_MAX).all()


def test_empty_index_and_padding(gpu):
    for metric, pad in (("l2", FLT_MAX), ("ip", -FLT_MAX)):
        ix = _mk(metric, 16)
        D, I = ix.search(np.ones((3, 16), np.float32), 5)
        assert (I == -1).all() and (D == pad).all()
        ix.add(np.eye(16, dtype=np.float32)[:2])
        D, I = ix.search(np.eye(16, dtype=np.float32)[:1], 4)
        assert I[0, :2].tolist() == [0, 1] and I[0, 2:].tolist() == [-1, -1] and (D[0, 2:] == pad).all()
        ix.reset()
        assert ix.ntotal == 0
        D, I = ix.search(np.ones((1, 16), np.float32), 2)
        assert (I == -1).all()


def test_shape_and_dtype_errors(gpu):
    ix = _mk("l2", 8)
    with pytest.raises(AssertionError):
        ix.add(np.zeros((3, 7), np.float32))
    with pytest.raises(AssertionError):
        ix.search(np.zeros((1, 9), np.float32), 1)
    with pytest.raises(AssertionError):
        ix.search(np.zeros(8, np.float32), 1)
    with pytest.raises(RuntimeError):
        ix.search(np.zeros((1, 8), np.float32), 5000)
    ix.add(np.zeros((2, 8)))          # float64 input is converted like faiss's SWIG wrapper does
    assert ix.ntotal == 2


@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_synthetic_small_vs_golden(gpu, golden_dir, metric):
    z = np.load(os.path.join(golden_dir, "synthetic_small.npz"))
    n, d, nq = int(z["n"]), int(z["d"]), int(z["nq"])
    x = synth.dense_corpus_np(n, d)
    x[100] = x[7]
    q = synth.dense_queries_np(x, nq)
    ix = _mk(metric, d)
    ix.add(x)
    D, I = ix.search(q, 50)
    st = ix.stats()
    assert st["mode_used"] == 0 and st["list_len"] == 128
    assert_topk_matches(D, I, z[f"{metric}_D"], z[f"{metric}_I"], TOL_F32, metric == "ip", f"auto/{metric}")
    assert recall_at_k(I[:, :10], z[f"{metric}_I"][:, :10]) == 1.0
    # the exact SIMT path (the certified fallback) gives bit-identical results
    ix.set_mode("exact")
    D2, I2 = ix.search(q, 50)
    assert np.array_equal(I, I2) and np.array_equal(D, D2)
    assert ix.stats()["mode_used"] == 1


def test_filter_certifies_random_queries_without_fallback(gpu):
    n, d, nq = 40000, 256, 300
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, nq)
    ix = _mk("ip", d)
    ix.add(x)
    D, I = ix.search(q, 10)
    st = ix.stats()
    assert st["flagged"] == 0 and st["overflow"] == 0, st   # tensor-core filter, no exact fallback needed
    o = flat.IndexFlatIP(d)
    o.add(x)
    Dr, Ir = o.search(q, 10, precision="f64")
    assert_topk_matches(D, I, Dr, Ir, TOL_F32, True, "random/ip")


def test_duplicates_force_fallback_and_keep_tie_rule(gpu):
    """300 identical rows: the filter cannot separate rank k from rank KL, the certificate must
    send the query to the exact scan, and ties resolve by ascending id."""
    d = 64
    x = synth.dense_corpus_np(5000, d)
    x[1000:1300] = x[17]
    ix = _mk("ip", d)
    ix.add(x)
    D, I = ix.search(x[17:18], 10)
    assert I[0].tolist() == [17] + list(range(1000, 1009))
    assert ix.stats()["flagged"] == 1
    ixl = _mk("l2", d)
    ixl.add(x)
    D, I = ixl.search(x[17:18], 10)
    assert I[0].tolist() == [17] + list(range(1000, 1009)) and np.allclose(D, 0, atol=1e-6)


@pytest.mark.parametrize("storage", ["f32", "f32+bf16"])
def test_near_duplicates_take_the_second_rescore_stage(gpu, storage):
    """200 rows within 1e-4 of each other, scattered over the corpus, next to the query: the filter cannot
    separate rank 10 from rank KL, but every candidate above the thresholds fits the deep shortlist, so the second
    re-score stage certifies the query without the exhaustive exact scan; the answer equals the exact scan bit for
    bit.  (The same rows stored contiguously fill one CTA's list, push the threshold into the cluster and
    rightly end in the exact scan: test_duplicates_force_fallback_and_keep_tie_rule.)"""
    d = 256
    rng = np.random.default_rng(8)
    x = synth.dense_corpus_np(60000, d)
    base = x[123].copy()
    where = np.sort(rng.choice(np.arange(1000, 60000), size=200, replace=False))
    x[where] = base + 1e-4 * rng.standard_normal((200, d)).astype(np.float32)
    q = np.concatenate([base[None, :], synth.dense_queries_np(x, 40)]).astype(np.float32)
    ix = _mk("ip", d, storage=storage)
    ix.add(x)
    D, I = ix.search(q, 10)
    st = ix.stats()
    assert st["deeper"] >= 1 and st["flagged"] == 0, st
    ix.set_mode("exact")
    De, Ie = ix.search(q, 10)
    assert np.array_equal(I, Ie) and np.array_equal(D, De)
    assert set(I[0]) <= set(where.tolist()) | {123}


@pytest.mark.parametrize("d", [1, 3, 31, 100, 384, 1000])
def test_odd_dimensions(gpu, d):
    rng = np.random.default_rng(d)
    x = rng.standard_normal((700, d)).astype(np.float32)
    q = rng.standard_normal((5, d)).astype(np.float32)
    for metric in ("ip", "l2"):
        ix, o = _mk(metric, d), _oracle(metric, d)
        ix.add(x), o.add(x)
        D, I = ix.search(q, 8)
        Dr, Ir = o.search(q, 8, precision="f64")
        scale = max(1.0, float(np.abs(Dr).max()))
        assert_topk_matches(D, I, Dr, Ir, 4e-6 * scale, metric == "ip", f"d={d}/{metric}")


def test_incremental_add_reconstruct_and_ragged_tiles(gpu):
    d = 40
    x = synth.dense_corpus_np(1000, d)
    ix = _mk("l2", d)
    for lo, hi in ((0, 1), (1, 257), (257, 600), (600, 1000)):   # crosses the 256-row tile boundary
        ix.add(x[lo:hi])
    assert ix.ntotal == 1000
    np.testing.assert_array_equal(ix.reconstruct_n(0, 1000), x)
    np.testing.assert_array_equal(ix.reconstruct(999), x[999])
    o = flat.IndexFlatL2(d)
    o.add(x)
    q = synth.dense_queries_np(x, 9)
    D, I = ix.search(q, 5)
    Dr, Ir = o.search(q, 5, precision="f64")
    assert_topk_matches(D, I, Dr, Ir, TOL_F32, False, "incremental")


def test_many_queries_cross_launch_batches(gpu):
    """nq > 2048 exercises the per-launch query batching; nq not a multiple of 128 the ragged tile."""
    n, d, nq = 3000, 64, 2500
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, nq)
    ix, o = _mk("ip", d), flat.IndexFlatIP(d)
    ix.add(x), o.add(x)
    D, I = ix.search(q, 3)
    Dr, Ir = o.search(q, 3, precision="f64")
    assert_topk_matches(D, I, Dr, Ir, TOL_F32, True, "nq=2500")


@pytest.mark.parametrize("k", [1, 16, 17, 100, 128, 200])
def test_k_range(gpu, k):
    n, d = 6000, 48
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, 20)
    ix, o = _mk("ip", d), flat.IndexFlatIP(d)
    ix.add(x), o.add(x)
    D, I = ix.search(q, k)
    Dr, Ir = o.search(q, k, precision="f64")
    assert_topk_matches(D, I, Dr, Ir, TOL_F32, True, f"k={k}")


def test_bf16_storage_matches_oracle_on_rounded_corpus(gpu):
    import torch
    n, d = 20000, 128
    x = synth.dense_corpus_np(n, d)
    xb = torch.from_numpy(x).to(torch.bfloat16).to(torch.float32).numpy()   # what bf16 storage holds
    q = synth.dense_queries_np(x, 50)
    for metric in ("ip", "l2"):
        ix, o = _mk(metric, d, storage="bf16"), _oracle(metric, d)
        ix.add(x), o.add(xb)
        assert ix.storage == "bf16"
        np.testing.assert_array_equal(ix.reconstruct_n(0, 10), xb[:10])
        D, I = ix.search(q, 10)
        Dr, Ir = o.search(q, 10, precision="f64")
        assert_topk_matches(D, I, Dr, Ir, TOL_F32, metric == "ip", f"bf16/{metric}")


@pytest.mark.parametrize("metric", ["ip", "l2"])
def test_bf16_shadow_filter_gives_the_fp32_answers_bit_for_bit(gpu, metric):
    """storage 'f32+bf16': the filter streams a bf16 copy of the rows, the re-score reads the fp32
    rows, so D and I must equal the plain fp32 index exactly (and reconstruct / file bytes too)."""
    n, d, nq = 60000, 200, 300           # d not a multiple of 64: both row layouts are padded
    x = synth.dense_corpus_np(n, d)
    x[500:520] = x[3]                     # a run of duplicates: certificate -> exact fallback
    q = synth.dense_queries_np(x, nq)
    q[0] = x[3]
    a, b = _mk(metric, d), _mk(metric, d, storage="f32+bf16")
    a.add(x[:25000]), a.add(x[25000:])
    b.add(x[:25000]), b.add(x[25000:])
    assert b.storage == "f32+bf16"
    np.testing.assert_array_equal(b.reconstruct_n(24990, 20), x[24990:25010])
    for k in (10, 50, 128):
        Da, Ia = a.search(q, k)
        Db, Ib = b.search(q, k)
        assert np.array_equal(Ia, Ib) and np.array_equal(Da, Db), (metric, k)
        st = b.stats()
        assert st["mode_used"] == 0 and st["list_len"] == min(256, 2 * a.stats()["list_len"])
    assert b.stats()["flagged"] <= 3      # only the duplicate query (and near-ties) may need the exact scan
    o = _oracle(metric, d)
    o.add(x)
    Dr, Ir = o.search(q, 10, precision="f64")
    Db, Ib = b.search(q, 10)
    assert_topk_matches(Db, Ib, Dr, Ir, TOL_F32, metric == "ip", f"shadow/{metric}")


def test_l2_ip_equivalence_on_unit_norm(gpu):
    n, d = 30000, 96
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, 64)
    a, b = _mk("l2", d), _mk("ip", d)
    a.add(x), b.add(x)
    Dl, Il = a.search(q, 10)
    Di, Ii = b.search(q, 10)
    np.testing.assert_allclose(1.0 - Dl / 2.0, Di, atol=3e-6)
    assert (Il == Ii).mean() > 0.99


def test_c0_config_recall(gpu):
    """BASELINE config C0 dense part: 100k x 1024 fp32, 1000 queries, top-10, vs the fp32 oracle
    (sgemm in 1024-row blocks, the faiss nq >= 20 algorithm)."""
    n, d, nq = 100_000, 1024, 1000
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, nq)
    ix, o = _mk("ip", d), flat.IndexFlatIP(d)
    ix.add(x), o.add(x)
    D, I = ix.search(q, 10)
    st = ix.stats()
    Dr, Ir = o.search(q, 10)
    assert_topk_matches(D, I, Dr, Ir, TOL_F32, True, "C0")
    assert recall_at_k(I, Ir) >= 0.9999
    assert st["flagged"] <= 5, st


def test_index_file_bytes_equal_reference_and_roundtrip(gpu, golden_dir, tmp_path):
    z = np.load(os.path.join(golden_dir, "ref_wrapper.npz"))
    ix = hf.IndexFlatL2(z["x"].shape[1])
    ix.add(z["x"])
    p = tmp_path / "a_faiss.index"
    hf.write_index(ix, str(p))
    assert p.read_bytes() == z["index_bytes"].tobytes()      # byte-identical to the reference-written file
    back = hf.read_index(str(p))
    assert (back.d, back.ntotal, back.metric_type) == (ix.d, ix.ntotal, hf.METRIC_L2)
    np.testing.assert_array_equal(back.reconstruct_n(), z["x"])
    ip = hf.IndexFlatIP(8)
    ip.add(np.eye(8, dtype=np.float32))
    hf.write_index(ip, str(tmp_path / "b.index"))
    assert (tmp_path / "b.index").read_bytes()[:4] == b"IxFI"
    assert hf.read_index(str(tmp_path / "b.index")).metric_type == hf.METRIC_INNER_PRODUCT
    with pytest.raises(RuntimeError):
        hf.read_index(str(tmp_path / "missing.index"))
    (tmp_path / "bad.index").write_bytes(b"nope" + b"\0" * 64)
    with pytest.raises(RuntimeError):
        hf.read_index(str(tmp_path / "bad.index"))


def test_device_tensors_in_out(gpu):
    import torch
    n, d = 5000, 64
    x = synth.dense_corpus_np(n, d)
    q = synth.dense_queries_np(x, 16)
    ix = _mk("ip", d)
    ix.add(torch.from_numpy(x).cuda())
    D, I = ix.search(torch.from_numpy(q).cuda(), 7)
    assert D.is_cuda and I.is_cuda and I.dtype == torch.int64
    Dh, Ih = ix.search(q, 7)
    assert np.array_equal(I.cpu().numpy(), Ih) and np.array_equal(D.cpu().numpy(), Dh)


def test_large_scan_properties(gpu):
    """Size-independent properties at a corpus far larger than L2 (2M x 1024 fp32 = 8 GB, built on
    device): planted queries return their planted row first with score ~0.995; a 2-shard split
    merged with hr_merge_topk equals the single index bit for bit; no query needs the fallback."""
    import torch
    from intool_rag_b200 import _lib
    n, d, nq, k = 2_000_000, 1024, 256, 10
    dev = torch.device("cuda", 0)
    full = hf.IndexFlatIP(d)
    planted = synth.dense_corpus_into(full, n, d, dev, keep_rows=4096)
    assert full.ntotal == n
    g = torch.Generator(device=dev)
    g.manual_seed(99)
    rows = torch.randint(0, 4096, (nq,), generator=g, device=dev)
    q = torch.nn.functional.normalize(planted[rows] + 0.1 / 32 * torch.randn((nq, d), generator=g, device=dev), dim=1)
    D, I = full.search(q, k)
    st = full.stats()
    assert (I[:, 0] == rows).all(), "planted row must rank first"
    assert (D[:, 0] > 0.99).all() and (D[:, 1] < 0.3).all()
    assert (D[:, :-1] >= D[:, 1:]).all()                       # sortedness
    assert st["flagged"] == 0 and st["mode_used"] == 0, st
    # idempotence
    D2, I2 = full.search(q, k)
    assert torch.equal(I, I2) and torch.equal(D, D2)
    # shard merge == full (uses two more indexes over the same rows, regenerated from the seed)
    half = n // 2
    parts = []
    gen = torch.Generator(device=dev)
    gen.manual_seed(synth.DENSE_SEED)
    a, b = hf.IndexFlatIP(d), hf.IndexFlatIP(d)
    a.reserve(half), b.reserve(n - half)
    b.set_id_base(half)
    done = 0
    for r0 in range(0, n, 1 << 18):
        nr = min(1 << 18, n - r0)
        xx = torch.nn.functional.normalize(torch.randn((nr, d), generator=gen, device=dev), dim=1)
        lo = max(0, half - r0)
        if lo > 0:
            a.add(xx[:min(lo, nr)])
        if lo < nr:
            b.add(xx[max(lo, 0):])
        done += nr
    assert a.ntotal == half and b.ntotal == n - half
    for ix in (a, b):
        parts.append(ix.search(q, k))
    Dg = torch.cat([parts[0][0], parts[1][0]], 1).contiguous()
    Ig = torch.cat([parts[0][1], parts[1][1]], 1).contiguous()
    Dm = torch.empty_like(D)
    Im = torch.empty_like(I)
    _lib.check(_lib.lib().hr_merge_topk(Dg.data_ptr(), Ig.data_ptr(), nq, 2 * k, k, 1, -3.4e38, Dm.data_ptr(),
                                        Im.data_ptr(), 0, _lib.current_stream_ptr(0)))
    torch.cuda.synchronize()
    assert torch.equal(Im, I) and torch.equal(Dm, D)
