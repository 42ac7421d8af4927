"""GPU: the hot-transition path of the order-1 rANS kernels (rans_chain.cu: a symbol that follows itself with probability
>= 63/64 is coded from registers, 16 steps per block) on the streams that take it — low-entropy exception-like streams of
many sizes, with rare interruptions, a second rare symbol, and the not-quite-hot case that must fall back — alone and
mixed with ordinary leaves inside one warp job.  Bytes vs the oracle, and round trip."""
import numpy as np, pytest
import orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("n", [100, 1000, 4099, 20000, 450001, 3000000])
def test_hot_streams(eng, n):
    rng = np.random.default_rng(5 + n)
    for p, sym in ((0.001, 78), (0.02, 1), (0.0, 0), (0.3, 5)):          # 78 = 'N' in the ACGT exception stream
        x = np.zeros(n, np.uint8)
        x[rng.random(n) < p] = sym
        if p == 0.02:
            x[rng.random(n) < 0.001] = 200
        for codec in ("RANB", "RANb", "RANW"):
            want = orc.compress("port", "rans", x, orc.ORDER[codec])
            got = eng.compress([(codec, x)])[0]
            assert got.size == want.size and np.array_equal(got, want), (n, p, codec)
            assert np.array_equal(eng.uncompress([(codec, want, n)])[0], x), (n, p, codec)


def test_hot_and_ordinary_leaves_share_warp_jobs(eng):
    rng = np.random.default_rng(9)
    items = []
    for i in range(40):
        n = int(rng.integers(50, 60000))
        x = np.zeros(n, np.uint8) if i % 3 else rng.integers(0, 9, n).astype(np.uint8)
        x[rng.random(n) < 0.002] = 78
        items.append(("RANB", x))
    got = eng.compress(items)
    for (c, x), g in zip(items, got):
        w = orc.compress("port", "rans", x, orc.ORDER[c])
        assert g.size == w.size and np.array_equal(g, w), x.size
    for (c, x), o in zip(items, eng.uncompress([(c, g, x.size) for (c, x), g in zip(items, got)])):
        assert np.array_equal(o, x), x.size
