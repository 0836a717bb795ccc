"""CPU: the reference arm of bench.py (the oracle timed on the host cores) prints exactly one JSON line on
stdout with the keys the driver reads; nothing under intool-rag_b200's CUDA path is touched."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--rows", "20000", "--cpu-sample-rows", "4000", "--nq", "32",
                          "--vocab", "2000", "--dim", "64"], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    j = json.loads(lines[0])
    assert j["impl"] == "reference" and j["unit"] == "queries/s" and j["higher_is_better"] is True
    assert j["value"] > 0 and j["cpu_baseline"]["kind"] == "port" and j["cpu_baseline"]["cores"] >= 1
    assert j["e2e"] == {"value": j["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in j["config"] and j["metric"].startswith("hybrid QPS")
    # the line reports what was MEASURED (the driver checks ms_per_step * steps against its own clock); the
    # full-corpus projection is a separate field, and the arm uses every host core whatever launched it
    cb = j["cpu_baseline"]
    assert j["config"]["rows"] == 4000 and cb["rows"] == 4000
    assert abs(j["value"] - 32 / (j["ms_per_step"] / 1e3)) < 1e-6 * j["value"]
    assert j["extrapolated"]["rows"] == 20000 and abs(j["extrapolated"]["value"] - j["value"] / 5) < 1e-9 * j["value"]
    assert cb["blas_threads"] == cb["cores"] == (os.cpu_count() or 1)
    assert set(cb["stage_ms"]) == {"dense", "bm25", "fusion"} and cb["nq1_ms_per_query"] > 0
    assert cb["recall_at_10_vs_fp64"] == 1.0


def test_reference_arm_ignores_torchrun_thread_limit():
    env = dict(os.environ, OMP_NUM_THREADS="1", RANK="0", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "0", "--rows", "8000", "--cpu-sample-rows", "2000", "--nq", "16",
                          "--vocab", "500", "--dim", "32"], capture_output=True, text=True, timeout=300, env=env)
    assert out.returncode == 0, out.stderr[-2000:]
    j = json.loads([l for l in out.stdout.splitlines() if l.strip()][0])
    assert j["cpu_baseline"]["blas_threads"] == (os.cpu_count() or 1)


def test_reference_arm_other_ranks_do_no_work():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""
