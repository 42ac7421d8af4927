"""The -m gpu parity tests, run WITHOUT a GPU (`--dry-gpu`, tests/conftest.py): genozip_b200.lib.Engine's own marshalling code
on top of tests/mock_gzb.py, which answers the C-ABI's entry points with the CPU checkers.  This proves nothing about the
kernels — it keeps the GPU tests themselves (inputs, expectations, the comparisons with the reference's compiled objects) and
the ctypes binding from rotting between GPU runs, so that a failure on the B200 box is a kernel failure."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_gpu_tests_pass_against_the_cpu_checkers():
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--dry-gpu", "-x", "-q", "-p", "no:cacheprovider"],
                       cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert " passed" in tail and "failed" not in tail, tail
