"""GPU: a short run of the randomised parity soak (tools/fuzz_parity.py): random shapes, metrics, storages,
duplicates, ragged / empty / out-of-vocabulary queries; dense auto path == exhaustive exact scan bit for bit, BM25
and hybrid scores within 1e-5 / 3e-5 relative of the oracle."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_fuzz_soak_20s(gpu):
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "fuzz_parity.py"), "20", "7"],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "fuzz ok" in out.stdout
