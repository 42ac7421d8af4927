"""The SIMT emulator (tests/host/simt) checks itself: every collective against its definition, sub-masks, partial warps, barriers
with exited threads, shared and dynamic shared memory, atomics, the integer/float intrinsics — and its diagnostics: a collective
inside divergent code, lanes of one mask at different collectives, __syncthreads in divergent code and a plain deadlock must
each abort the run with a report (on the GPU: a hang or garbage)."""
import os, subprocess, pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SIMT = os.path.join(HERE, "host", "simt")
EXE = os.path.join(HERE, "host", "_build", "simt_selftest")


@pytest.fixture(scope="module")
def exe():
    os.makedirs(os.path.dirname(EXE), exist_ok=True)
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g1", "-w", "-I", SIMT, os.path.join(SIMT, "selftest.cpp"), os.path.join(SIMT, "simt.cpp"), "-o", EXE],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return EXE


def test_collectives_barriers_intrinsics(exe):
    r = subprocess.run([exe, "0"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "selftest: ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("mode,what", [(1, "deadlock"), (2, "deadlock"), (3, "DIFFERENT __syncthreads"), (4, "deadlock")])
def test_diagnostics(exe, mode, what):
    r = subprocess.run([exe, str(mode)], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and what in r.stderr and "waiting at" in r.stderr, r.stdout + r.stderr
