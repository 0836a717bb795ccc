"""ORACLE — test infrastructure only (CPU restatement of the reference's hot path).

Imported only by tests/, __graft_entry__.smoke() and bench.py's CPU baseline / reference arm.
The product package never imports this; without the CUDA library the product fails loudly.
See oracle/flat.py, oracle/bm25.py, oracle/fusion.py for per-function reference citations and
the parity status (unpinned by the reference: it ships no tests, no BM25 and no fusion).
"""
from . import flat, bm25, fusion  # noqa: F401
from .hybrid import retrieve  # noqa: F401
