"""Pins the CPU restatement (oracle/hts_port.c) against the reference's own compiled htscodecs
(oracle/_ref/libhts_ref.so): byte-identical compressed output and exact round trips, for all eight
genozip order bytes (codec_htscodecs.c:17-20)."""
import numpy as np, pytest
import orc
from datagen import stream, KINDS, EDGE_SIZES

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built and /root/reference absent")

CODECS = [("rans", "RANB"), ("rans", "RANW"), ("rans", "RANb"), ("rans", "RANw"),
          ("arith", "ARTB"), ("arith", "ARTW"), ("arith", "ARTb"), ("arith", "ARTw")]


def check(kind, name, data):
    order = orc.ORDER[name]
    a = orc.compress("ref", kind, data, order)
    b = orc.compress("port", kind, data, order)
    assert a.size == b.size and np.array_equal(a, b), f"{name} n={data.size}: port != ref (len {b.size} vs {a.size})"
    if data.size:
        assert np.array_equal(orc.uncompress("port", kind, a, data.size), data)
        assert np.array_equal(orc.uncompress("ref", kind, b, data.size), data)


@pytest.mark.parametrize("kind,name", CODECS)
def test_edge_sizes(kind, name):
    for n in EDGE_SIZES:
        for dk in ("skew8", "uniform256", "two", "const"):
            check(kind, name, stream(dk, n, 7 + n))


@pytest.mark.parametrize("kind,name", CODECS)
@pytest.mark.parametrize("dk", KINDS)
def test_kinds(kind, name, dk):
    for n in (777, 50021, 300000):
        check(kind, name, stream(dk, n, 11))


@pytest.mark.parametrize("kind,name", [("rans", "RANB"), ("rans", "RANw"), ("arith", "ARTB")])
def test_large(kind, name):
    check(kind, name, stream("qual", 3_000_000, 5))   # > 500000 takes hist1_4's split-table path (utils.h:146)
    check(kind, name, stream("skew8", 1_200_001, 6))


def test_bounds():
    L, R = orc.port(), orc.ref()
    rng = np.random.default_rng(3)
    sizes = list(range(0, 5000)) + [int(x) for x in rng.integers(0, 2**31 - 2**27, size=20000)]
    for n in sizes:
        for o in (0x01, 0x19, 0x81, 0x99, 0, 0x10, 0x11, 0x90, 0x50):
            assert L.orc_rans_bound(n, o) == R.rans_compress_bound_4x16(n, o)
            assert L.orc_arith_bound(n, o) == R.arith_compress_bound(n, o)


def _shift_of(comp, n):
    off = 1 + (1 if n < 128 else 2 if n < 16384 else 3 if n < 2097152 else 4)
    return int(comp[off] >> 4)


def test_o1_shift_decision_fuzz():
    """compute_shift (rANS_static4x16pr.c:626-687) is double arithmetic with FMA contraction; fuzz many
    order-1 streams so that both outcomes (10- and 12-bit tables) are exercised."""
    rng = np.random.default_rng(99)
    seen = {10: 0, 12: 0}
    for t in range(300):
        big = t % 5 == 0
        k = int(rng.integers(8, 60)) if big else int(rng.integers(2, 40))
        n = int(rng.integers(50000, 300000)) if big else int(rng.integers(64, 6000))
        p = rng.dirichlet(np.full(k, 0.05 if big else rng.uniform(0.05, 2.0))) + 1e-5
        p /= p.sum()
        data = rng.choice(np.arange(k, dtype=np.uint8) + 40, size=n, p=p).astype(np.uint8)
        a = orc.compress("ref", "rans", data, 0x01)
        b = orc.compress("port", "rans", data, 0x01)
        assert np.array_equal(a, b), f"trial {t}"
        if a[0] & 1:
            seen[_shift_of(a, n)] += 1
    assert seen[10] > 10 and seen[12] > 5, seen
