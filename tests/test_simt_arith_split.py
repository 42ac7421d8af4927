"""The split arithmetic encoder (genozip_b200/csrc/arith_split.cu: per-context model warps + one range-coder warp per leaf) on
the SIMT emulator, with its threshold lowered (GZB_AR_SPLIT_MIN=64) so that the edge sizes, every stream kind and the fuzzer's
shapes go through it: the bytes must be the reference's, exactly as from the fused chain."""
import os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout):
    env = dict(os.environ, GZB_AR_SPLIT_MIN="64", GZB_SIMT_QUICK="1")
    r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    return r.stdout + r.stderr


def test_parity_tests_through_the_split_encoder():
    out = _run(["-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--simt", "-x", "-q", "-p", "no:cacheprovider",
                "-k", "(edge_sizes and ART) or golden or soft_fail"], 1500)
    assert " passed" in out and "failed" not in out, out[-2000:]


def test_fuzz_through_the_split_encoder():
    out = _run([os.path.join(ROOT, "tools", "fuzz_simt.py"), "--seconds", "20", "--seed", "31", "--max-n", "30000"], 600)
    assert out.strip().splitlines()[-1].startswith("ok:"), out[-2000:]
