"""The long-leaf arithmetic decoder (genozip_b200/csrc/arith_chain.cu k_arith_decode_long: a mirror of every context's model head
and first 16 / 32 entries in shared memory) on the SIMT emulator, with its threshold lowered (GZB_AR_LONG_MIN=64) so that the
edge sizes, every stream kind, the golden vectors and the fuzzer's shapes — damaged streams included — go through it: bit-exact
output and the reference's verdicts, exactly as from the general kernel."""
import os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, timeout, ent):
    env = dict(os.environ, GZB_AR_LONG_MIN="64", GZB_AR_LONG_ENT=str(ent), GZB_SIMT_QUICK="1")
    r = subprocess.run([sys.executable] + args, cwd=ROOT, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, (r.stdout + r.stderr)[-4000:]
    return r.stdout + r.stderr


@pytest.mark.parametrize("ent", [16])      # (32 entries per context: the same code, a constant; the kernel is off by default)
def test_parity_tests_through_the_long_decoder(ent):
    out = _run(["-m", "pytest", os.path.join(ROOT, "tests"), "-m", "gpu", "--simt", "-x", "-q", "-p", "no:cacheprovider",
                "-k", "(edge_sizes and ART) or golden or corrupt"], 1500, ent)
    assert " passed" in out and "failed" not in out, out[-2000:]


def test_fuzz_through_the_long_decoder():
    out = _run([os.path.join(ROOT, "tools", "fuzz_simt.py"), "--seconds", "25", "--seed", "47", "--max-n", "30000"], 600, 16)
    assert out.strip().splitlines()[-1].startswith("ok:"), out[-2000:]
