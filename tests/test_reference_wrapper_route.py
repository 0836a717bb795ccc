"""INTEGRATION.md's primary route: the reference's UNMODIFIED storage wrapper
(/root/reference/rag/storage/faiss_index.py:13-17 import, :54 read_index, :83 search, :123-124 IndexFlatL2 + add,
:133 write_index) running on ``sys.modules["faiss"] = intool_rag_b200.faiss``.

Needs both the reference tree and a B200 (the GPU box has no /root/reference, the build container has no GPU), so
each test is gated on what it needs; where only the reference exists, the route is still exercised up to the first
CUDA call, which must fail loudly (no CPU fallback)."""
import asyncio
import importlib
import json
import os
import sys

import numpy as np
import pytest

import intool_rag_b200  # noqa: F401
from intool_rag_b200 import _lib

REF = os.environ.get("HR_REFERENCE_DIR", "/root/reference")
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "rag", "storage")), reason="reference tree not mounted")


def _import_reference_wrapper(tmp_path, monkeypatch):
    from intool_rag_b200 import faiss as hr_faiss
    monkeypatch.setenv("STORAGE_DIR", str(tmp_path / "storages"))
    monkeypatch.setenv("CACHE_DIR", str(tmp_path / "cache"))
    monkeypatch.setenv("LOG_LEVEL", "WARNING")
    monkeypatch.chdir(tmp_path)                      # rag.config creates ./storages relative to the CWD
    monkeypatch.setitem(sys.modules, "faiss", hr_faiss)
    monkeypatch.syspath_prepend(REF)
    for name in [m for m in sys.modules if m == "rag" or m.startswith("rag.")]:
        monkeypatch.delitem(sys.modules, name)
    return importlib.import_module("rag.storage.faiss_index")


@needs_ref
@pytest.mark.skipif(_lib.device_count() > 0, reason="checks the no-GPU failure mode of the route")
def test_reference_wrapper_accepts_the_module_and_fails_loudly_without_a_gpu(tmp_path, monkeypatch):
    ref_fi = _import_reference_wrapper(tmp_path, monkeypatch)
    assert ref_fi.HAS_FAISS                           # the reference took the module for faiss
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ref_fi.create_faiss_index([[0.0, 1.0], [1.0, 0.0]])


@needs_ref
@pytest.mark.gpu
def test_unmodified_reference_wrapper_on_the_cuda_module(gpu, golden_dir, tmp_path, monkeypatch):
    """The recorded reference run (tests/golden/ref_wrapper.*, made with the oracle as faiss) must be reproduced
    by the same unmodified code over the CUDA module: file bytes, (id, score) lists, enriched hits."""
    z = np.load(os.path.join(golden_dir, "ref_wrapper.npz"))
    g = json.load(open(os.path.join(golden_dir, "ref_wrapper.json")))
    ref_fi = _import_reference_wrapper(tmp_path, monkeypatch)
    assert ref_fi.HAS_FAISS
    index = ref_fi.create_faiss_index([list(map(float, r)) for r in z["x"]])
    sd = tmp_path / "storages"
    path = sd / f"{g['doc_id']}_faiss.index"
    ref_fi.save_faiss_index(index, str(path))
    assert path.read_bytes() == z["index_bytes"].tobytes()
    n = g["n"]
    chunks = [{"chunk_id": f"c_{i // 4}_{i % 4:02d}", "page": i // 4 + 1, "text": f"chunk text {i}", "chunk_index": i}
              for i in range(n)]
    (sd / f"{g['doc_id']}_chunks.json").write_text(json.dumps({"total": n, "chunks": chunks}))
    reader = ref_fi.FAISSIndexReader(str(path))
    assert reader.get_dimension() == g["reader_dimension"] and reader.get_size() == g["reader_size"]
    for qi, q in enumerate(z["queries"]):
        for k in (1, 5, 10, 50):
            got = reader.search(list(map(float, q)), top_k=k)
            want = g["reader_search"][f"q{qi}_k{k}"]
            assert [i for i, _ in got] == [w[0] for w in want]
            np.testing.assert_allclose([s for _, s in got], [w[1] for w in want], atol=2e-6)
        got = asyncio.run(ref_fi.search_faiss_by_vector(list(map(float, q)), limit=7))
        want = g["search_by_vector"][f"q{qi}_l7"]
        assert [h["chunk_id"] for h in got] == [h["chunk_id"] for h in want]
