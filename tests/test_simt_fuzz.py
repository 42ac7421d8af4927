"""Differential fuzzing of the CUDA codec kernels on the SIMT emulator (tools/fuzz_simt.py) — a short seeded pass in the CPU
suite: random streams (alphabets of 1..256 symbols, iid / order-1 / run / striped / hot shapes, edge sizes) through all eight
codecs against the reference's own compiled htscodecs, and damaged streams through both decoders: same verdict, same bytes.
Plus the regressions the fuzzer found."""
import os, sys
import numpy as np, pytest
import orc

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref/libhts_ref.so not built and /root/reference absent")


@pytest.fixture(scope="module")
def eng():
    from simt_lib import simt_engine_class
    e = simt_engine_class()(0)
    yield e
    e.close()


def test_valid_streams():
    import fuzz_simt
    n, nbytes, _, _ = fuzz_simt.run(600, 101, False, 20000, max_streams=90)      # a fixed number of streams: the same cases on any machine
    assert n >= 90 and nbytes > 50000


def test_damaged_streams():
    import fuzz_simt
    stricter = {}
    n, _, n_damaged, n_rejected = fuzz_simt.run(600, 102, True, 8000, stricter=stricter, max_streams=150)
    assert n_damaged > 40 and 0 < n_rejected < n_damaged
    assert sum(stricter.values()) <= n_damaged // 10, stricter            # the documented stricter classes stay the exception


@pytest.mark.skipif(not orc.have_gz_ref(), reason="oracle/_ref/libgz_ref.so not built")
def test_genozip_codec_kernels():
    """ACGT / DOMQ / PBWT / LONGR kernels against the reference's compiled codec objects on random VBlocks (tools/fuzz_simt_gz.py), then the
    round-2 kernels: OQ, SMUX, PACB, HOMP / T0, b250"""
    import fuzz_simt_gz
    count = fuzz_simt_gz.run(600, 103, which=["domq", "acgt", "pbwt", "longr"], max_cases=160)
    assert all(v >= 40 for v in count.values()), count
    count = fuzz_simt_gz.run(600, 104, which=["oq", "smux", "pacb", "homp", "b250"], max_cases=100)
    assert all(v >= 20 for v in count.values()), count


def _rejects(eng, codec, comp, n):
    from genozip_b200 import GzbError
    try:
        eng.uncompress([(codec, np.asarray(comp, np.uint8), n)])
    except GzbError:
        return True
    return False


def test_container_without_payload_is_refused(eng):
    """flags + size and nothing else: htscodecs returns an empty result (rANS_static4x16pr.c:1598-1601, arith_dynamic.c:1076-1079)
    and the plug-in's out_len == uncompressed_len check aborts (codec_htscodecs.c:111,126)"""
    for codec, flags in (("RANB", 0x20), ("RANB", 0x00), ("RANB", 0x01), ("ARTB", 0x20), ("ARTB", 0x00), ("ARTb", 0x01)):
        for n in (1, 5, 100):
            assert _rejects(eng, codec, [flags, n], n), (codec, flags, n)
            with pytest.raises(AssertionError):
                orc.uncompress("ref", "rans" if codec.startswith("RAN") else "arith", np.array([flags, n], np.uint8), n)
    # ... but a one-symbol PACK map needs no payload: a constant stream decodes from its meta data alone — whatever the other flags say
    for flags in (0x80 | 0x20 | 0x04, 0x80 | 0x40 | 0x10 | 0x04 | 0x01):
        body = [flags] + ([] if flags & 0x10 else [4]) + [1, 232, 0]
        assert list(orc.uncompress("ref", "arith", np.array(body, np.uint8), 4)) == [232] * 4
        assert list(eng.uncompress([("ARTb", np.array(body, np.uint8), 4)])[0]) == [232] * 4
    x = np.full(1000, 65, np.uint8)
    for codec in ("RANb", "ARTb"):
        comp = orc.compress("ref", "rans" if codec.startswith("RAN") else "arith", x, orc.ORDER[codec])
        assert np.array_equal(eng.uncompress([(codec, comp, x.size)])[0], x)


def test_rans_state_below_the_bound_is_refused(eng):
    """RansDecInit ... if (R < RANS_BYTE_L) goto err (rANS_static4x16pr.c:555-558, :1023-1026)"""
    from datagen import stream
    for codec, kind in (("RANB", "skew8"), ("RANb", "qual"), ("RANW", "u32le")):
        x = stream(kind, 5000, 3)
        comp = orc.compress("ref", "rans", x, orc.ORDER[codec])
        assert np.array_equal(eng.uncompress([(codec, comp, x.size)])[0], x)
        if codec == "RANW":
            continue
        bad = comp.copy()
        bad[-16:] = 0                                                       # (the four final states are the LAST 16 bytes only by construction of
        for k in range(comp.size - 16, 16, -1):                            #  this loop: find the window whose zeroing the reference refuses)
            bad = comp.copy(); bad[k:k + 16] = 0
            try:
                orc.uncompress("ref", "rans", bad, x.size)
            except AssertionError:
                break
        else:
            pytest.skip("no refusing window found")
        assert _rejects(eng, codec, bad, x.size)


def test_order1_context_without_frequencies_is_defined(eng):
    """a context of the alphabet whose frequency row is empty (rANS_static4x16pr.c:994-997): a valid stream never enters it; a
    damaged one that does must find a defined row, not arena leftovers — same bytes on every run"""
    # alphabet {0, 33, 215}; context 0 -> 33, context 33 -> {33, 215}, context 215: no frequencies
    body = [0x01, 40, 0xA0, 0, 33, 215, 0] + [0, 0, 4, 0] + [0, 0, 3, 1] + [0, 2] + [0, 128, 0, 0] * 4 + [7, 9, 11, 13, 200, 100, 50, 25]
    outs = []
    for _ in range(3):
        try:
            outs.append(eng.uncompress([("RANB", np.array(body, np.uint8), 40)])[0].copy())
        except Exception:
            outs.append(None)
        eng.compress([("RANB", np.random.default_rng(len(outs)).integers(0, 256, 30000, dtype=np.uint8))])   # other traffic through the arena
    assert all((o is None) == (outs[0] is None) and (o is None or np.array_equal(o, outs[0])) for o in outs)


def test_frequency_sums_that_wrap_are_refused(eng):
    """a varint is any 32-bit value: F['A'] = 0xFFFFFFFF and F['C'] = 4097 sum to 4096 modulo 2^32.  The reference checks every
    frequency against what is left of the table (rANS_static4x16pr.c:538, :1003-1005); a table build that only looks at the
    32-bit total would fill billions of LUT slots past its arena"""
    states = [0, 128, 0, 0] * 4
    o0 = [0x00, 100, 0x41, 0x43, 0x00, 0x8F, 0xFF, 0xFF, 0xFF, 0x7F, 0xA0, 0x01] + states
    assert _rejects(eng, "RANB", o0, 100)
    with pytest.raises(AssertionError):
        orc.uncompress("ref", "rans", np.array(o0, np.uint8), 100)
    ctl = [0x00, 100, 0x41, 0x43, 0x00, 0x90, 0x00, 0x90, 0x00] + states       # the control: 2048 / 2048 decodes
    assert eng.uncompress([("RANB", np.array(ctl, np.uint8), 100)])[0].size == 100
    # order 1: alphabet {0, 65}; context 0's row carries the wrapping pair
    o1 = [0x01, 40, 0xC0, 0, 65, 0] + [0x8F, 0xFF, 0xFF, 0xFF, 0x7F, 0xA0, 0x01] + [0x90, 0x00, 0x90, 0x00] + states
    assert _rejects(eng, "RANb", o1, 40)
