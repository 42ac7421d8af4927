"""The product's CUDA kernels, executed WITHOUT a GPU: genozip_b200/csrc/*.cu compiled by g++ against tests/host/simt (a
stand-in cuda_runtime.h plus a lock-step SIMT emulator: one fibre per thread, warp collectives and __syncthreads as
rendez-vous points that also detect collectives in divergent code) into tests/host/_build/libgzb200_simt.so, and the -m gpu
parity tests run against it through the same C-ABI and the same ctypes binding (`pytest -m gpu --simt`).

This checks the LOGIC of every kernel — rANS and arithmetic chains, histograms, table construction, container framing, ACGT,
DOMQ, PBWT, LONGR — byte for byte against the reference's compiled objects, on every commit, on a machine without a GPU.  It
does not replace the GPU run: timing, the memory model and the real instruction set are not emulated.

Here: a quick slice (about a minute).  The full GPU suite takes ~11 minutes on the emulator: `python -m pytest tests -m gpu --simt`."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
QUICK = ("edge_sizes or acgt or domq_ragged or domq_edges or domq_fastq_batch or pbwt or longr or share_warp or soft_fail or corrupt "
         "or golden or hot_streams and not 450001 and not 3000000")


import pytest


def test_kernels_do_not_depend_on_the_lane_order():
    """between two rendez-vous points the emulator runs the lanes of a warp one after the other; results must not depend on which
    lane goes first (SIMT_LANE_ORDER) — code that only works because lane 0 happens to run first would rely on more than the
    markers GZB_WARP_READS_DONE / AR_READS_DONE state.  (The two orders run side by side: two processes.)"""
    sys.path.insert(0, os.path.join(ROOT, "tests", "host", "simt"))
    import build as simt_build
    simt_build.build()                                   # (built once here, not by the two processes at the same time)
    k = ("(edge_sizes and (RANB or ARTB or ARTw)) or acgt or domq_ragged or domq_edges or domq_tiles or pbwt or longr or share_warp "
         "or oq_batch or smux or tmpl or pacb or homp or b250 or transpose")
    ps = {order: subprocess.Popen([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--simt", "-k", k, "-x", "-q", "-p", "no:cacheprovider"],
                                  cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                                  env=dict(os.environ, GZB_SIMT_QUICK="1", SIMT_LANE_ORDER=order)) for order in ("desc", "random")}
    outs = {order: (p.communicate(timeout=1500)[0], p.returncode) for order, p in ps.items()}
    for order, (out, rc) in outs.items():
        tail = out[-4000:]
        assert rc == 0 and " passed" in tail and "failed" not in tail, (order, tail)


def test_kernels_on_the_simt_emulator():
    env = dict(os.environ, GZB_SIMT_QUICK="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--simt", "-k", QUICK, "-x", "-q", "-p", "no:cacheprovider"],
                       cwd=ROOT, capture_output=True, text=True, timeout=1500, env=env)
    tail = (r.stdout + r.stderr)[-4000:]
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
