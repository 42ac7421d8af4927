"""CPU-only: the C-ABI library builds, loads and exports every symbol include/gzb200.h declares; no compute without a GPU."""
import ctypes, os, re
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared():
    h = open(os.path.join(ROOT, "include", "gzb200.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    names = set(re.findall(r"\b(gzb_[A-Za-z0-9_]+)\s*\(", h))
    names |= set(re.findall(r"GZB_(?:UN)?COMPRESS\s*\(\s*(gzb_\w+)\s*\)", h))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from genozip_b200 import build as b
    lib = ctypes.CDLL(b.build())
    names = declared()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, f"declared in include/gzb200.h but not exported: {missing}"


def test_no_cpu_fallback():
    """without a CUDA device the product refuses to run instead of falling back"""
    import genozip_b200
    L = genozip_b200.load()
    if L.gzb_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(genozip_b200.GzbError):
        genozip_b200.Engine(0)


def test_est_size_matches_reference_bounds():
    """gzb_est_size == codec_*_est_size = 1 KB + rans_compress_bound_4x16 / arith_compress_bound (codec_htscodecs.c:26-33)"""
    import numpy as np, orc
    import genozip_b200
    rng = np.random.default_rng(1)
    sizes = list(range(0, 3000, 7)) + [int(x) for x in rng.integers(0, 2**31 - 2**27, 3000)]
    for name, kind in (("RANB", "rans"), ("RANW", "rans"), ("RANb", "rans"), ("RANw", "rans"), ("ARTB", "arith"), ("ARTW", "arith"), ("ARTb", "arith"), ("ARTw", "arith")):
        for n in sizes:
            assert genozip_b200.est_size(name, n) == orc.est_size(kind, n, orc.ORDER[name]), (name, n)


def test_vb_round_robin():
    import genozip_b200
    L = genozip_b200.load()
    assert [L.gzb_vb_device(i, 8) for i in range(1, 18)] == [(i - 1) % 8 for i in range(1, 18)]
    assert L.gzb_vb_device(5, 1) == 0
