"""genozip_b200 — B200 (sm_100a) implementation of genozip's per-VBlock codec path.

The product is the C-ABI shared library ``libgzb200.so`` (see include/gzb200.h); this package is only the
Python-side binding used by the tests and by bench.py.  It never imports anything under ``oracle/`` and has
no CPU fallback: loading fails loudly if the CUDA library is missing, and every call fails on a box without
a B200.
"""
from .lib import load, GzbError, Engine, CODEC, est_size  # noqa: F401
