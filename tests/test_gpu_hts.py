"""GPU parity tests for the simple codecs: the CUDA path (through the C-ABI) must emit the same bytes as the
reference codec function — checked against oracle/_ref (the reference's own objects) when it travelled to this
box, and always against the CPU restatement — and decode bit-exactly."""
import numpy as np, pytest
import orc
from datagen import stream, KINDS, EDGE_SIZES

pytestmark = pytest.mark.gpu

RANS = ["RANB", "RANW", "RANb", "RANw"]
ARITH = ["ARTB", "ARTW", "ARTb", "ARTw"]


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def kind_of(name):
    return "rans" if name.startswith("RAN") else "arith"


def expected(name, data):
    impl = "ref" if orc.have_ref() else "port"
    return orc.compress(impl, kind_of(name), data, orc.ORDER[name])


def run_batch(eng, cases):
    """cases: list of (codec, data). One batched GPU call each way; compare with the oracle."""
    comp = eng.compress(cases)
    bad = []
    for (name, data), c in zip(cases, comp):
        want = expected(name, data)
        if c.size != want.size or not np.array_equal(c, want):
            first = int(np.argmax(c[:min(c.size, want.size)] != want[:min(c.size, want.size)])) if min(c.size, want.size) else 0
            bad.append((name, data.size, c.size, want.size, first))
    assert not bad, f"compressed bytes differ from the reference (codec, n, got_len, want_len, first_diff): {bad[:8]}"
    # decode the REFERENCE's bytes on the GPU (and ours are identical to them)
    dec_cases = [(name, c, data.size) for (name, data), c in zip(cases, comp) if data.size]
    outs = eng.uncompress(dec_cases)
    k = 0
    for (name, data) in cases:
        if not data.size:
            continue
        assert np.array_equal(outs[k], data), f"{name} n={data.size}: GPU decode mismatch"
        k += 1


@pytest.mark.parametrize("name", RANS + ARITH)
def test_edge_sizes(eng, name):
    cases = []
    for n in EDGE_SIZES:
        for dk in ("skew8", "uniform256", "two", "const"):
            cases.append((name, stream(dk, n, 7 + n)))
    run_batch(eng, cases)


@pytest.mark.parametrize("name", RANS + ARITH)
def test_kinds(eng, name):
    cases = []
    for dk in KINDS:
        for n in (777, 50021, 300000):
            cases.append((name, stream(dk, n, 11)))
    run_batch(eng, cases)


def test_mixed_batch(eng):
    """all eight codecs in one batch, ragged sizes — the shape a VBlock's sections have"""
    rng = np.random.default_rng(1)
    cases = []
    for i in range(200):
        name = (RANS + ARITH)[i % 8]
        dk = KINDS[int(rng.integers(0, len(KINDS)))]
        n = int(rng.integers(1, 40000))
        cases.append((name, stream(dk, n, 100 + i)))
    run_batch(eng, cases)


@pytest.mark.parametrize("name", ["RANB", "RANw", "ARTB"])
def test_large(eng, name):
    run_batch(eng, [(name, stream("qual", 3_000_000, 5)), (name, stream("skew8", 1_200_001, 6))])


def test_o1_shift_decision(eng):
    """both outcomes of the double-precision 10/12-bit table decision (rANS_static4x16pr.c:626-687)"""
    rng = np.random.default_rng(99)
    cases = []
    for t in range(120):
        big = t % 3 == 0
        k = int(rng.integers(8, 60)) if big else int(rng.integers(2, 40))
        n = int(rng.integers(50000, 300000)) if big else int(rng.integers(64, 6000))
        p = rng.dirichlet(np.full(k, 0.05 if big else rng.uniform(0.05, 2.0))) + 1e-5
        p /= p.sum()
        cases.append(("RANB", rng.choice(np.arange(k, dtype=np.uint8) + 40, size=n, p=p).astype(np.uint8)))
    run_batch(eng, cases)


def test_soft_fail(eng):
    """capacity below est_size: the reference returns false under soft_fail (src/compressor.c:90)"""
    from genozip_b200.lib import Section, CODEC
    import ctypes as C
    data = stream("skew8", 5000, 1)
    out = np.empty(100, np.uint8)
    s = (Section * 1)()
    s[0].codec = CODEC["RANB"]; s[0].in_ = data.ctypes.data; s[0].in_len = data.size
    s[0].out = out.ctypes.data; s[0].out_cap = out.size
    eng.compress_raw(s, 1)
    assert s[0].status == 1 and s[0].out_len == 0


def test_corrupt_is_reported(eng):
    from genozip_b200 import GzbError
    data = stream("skew8", 5000, 1)
    comp = eng.compress([("RANB", data)])[0].copy()
    comp[0] = 0x08 | 0x40   # claims STRIPE + rubbish
    with pytest.raises(GzbError):
        eng.uncompress([("RANB", comp, data.size)])


def test_packed_output(eng):
    """gzb_compress_sections_packed: the same bytes as the per-section call, appended to one buffer; a buffer that is too small
    is reported with the size that is needed and nothing is written"""
    from datagen import stream
    items = [(c, stream(k, n, 5 + i)) for i, (c, k, n) in enumerate((("RANB", "skew8", 40000), ("ARTW", "u32le", 9000), ("RANw", "qual", 77), ("ARTb", "qual", 30001),
                                                                     ("RANW", "u32le", 20), ("ARTB", "skew8", 1), ("RANb", "const", 100000)))]
    want = eng.compress(items)
    got, used = eng.compress_packed(items, arena_cap=64)                    # too small at first: grown to the reported size
    assert used == sum((len(w) + 15) & ~15 for w in want)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)
    got2, _ = eng.compress_packed(items)
    for g, w in zip(got2, want):
        assert np.array_equal(g, w)
