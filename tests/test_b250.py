"""gzb_b250_generate_batch: b250_zip_generate (src/b250.c:202-297) — the segmenter's b250 buffer (type in the last byte of every word)
converted to the PIZ form.  CPU: the restatement against the reference's compiled b250.c (oracle/_ref).  GPU (-m gpu, also --simt):
the chunked backward walk against both."""
import numpy as np
import pytest

import orc

ONE_UP, EMPTY, MISSING = -2, -3, -4


def seg_word(wi, force4=False):
    """b250_set_wi (:89-121) in the segmenter's form: little endian, the type in the last byte; a node new to the VBlock always takes 4 bytes (:155-158)"""
    if force4:
        enc, n = (7 << 29) | wi, 4
    elif 0 <= wi <= 126:
        enc, n = wi, 1
    elif 127 <= wi <= 16508:
        enc, n = (2 << 14) | (wi - 127), 2
    elif 16509 <= wi <= 2113660:
        enc, n = (6 << 21) | (wi - 16509), 3
    elif wi == EMPTY:
        enc, n = 0xBFFE, 2
    elif wi == MISSING:
        enc, n = 0xBFFF, 2
    else:
        enc, n = (7 << 29) | wi, 4
    return [(enc >> (8 * k)) & 0xff for k in range(n)]


def make_ctx(seed, n_words, ol_len, n_new, runs=True):
    """a context whose dictionary has ol_len words from earlier VBlocks and n_new nodes of its own (segged with 4 bytes each time they appear)"""
    rng = np.random.default_rng(seed)
    ni2wi = (ol_len + rng.permutation(n_new)).astype(np.int32)          # the merge gave the new nodes these word indices
    out = []
    wi = int(rng.integers(0, max(1, ol_len)))
    for _ in range(n_words):
        r = rng.random()
        if runs and r < 0.3:
            wi = wi + 1 if wi + 1 < ol_len else 0                       # consecutive words: ONE_UP candidates
            out += seg_word(wi)
        elif r < 0.35:
            out += seg_word(EMPTY if rng.random() < 0.5 else MISSING)
        elif r < 0.55 and n_new:
            out += seg_word(ol_len + int(rng.integers(0, n_new)), force4=True)
        else:
            wi = int(rng.integers(0, max(1, ol_len)))
            out += seg_word(wi)
    return np.array(out, np.uint8), ni2wi


CASES = [(1, 3000, 100, 10), (2, 5000, 2000, 50), (3, 20000, 40000, 3000), (4, 1, 5, 0), (5, 64, 3000000, 100), (6, 700, 900, 200), (7, 0, 10, 0)]


@pytest.mark.parametrize("seed,n_words,ol_len,n_new", CASES)
def test_port_matches_reference(seed, n_words, ol_len, n_new):
    if not orc.have_gz_ref():
        pytest.skip("the reference is not here")
    b, t = make_ctx(seed, n_words, ol_len, n_new)
    up = ol_len + n_new > 1024
    p = orc.b250_generate(b, t, ol_len, up, "port")
    r = orc.b250_generate(b, t, ol_len, up, "ref")
    assert p is not None and r is not None
    assert p[1] == n_words and np.array_equal(p[0], r[0]), (p[0][:20], r[0][:20])
    if up and n_words > 1000:
        assert (p[0] == 127).sum() > 0                                  # ONE_UP was used


def test_port_rejects_a_buffer_the_words_do_not_tile():
    b, t = make_ctx(9, 500, 3000, 10)
    bad = np.concatenate([[0x80], b]).astype(np.uint8)                  # one byte too many at the start: the first word would begin before the buffer
    assert orc.b250_generate(bad, t, 3000, True, "port") is None


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    return Engine(0)


@pytest.mark.gpu
def test_gpu_b250_batch(eng):
    items = []
    for seed, n_words, ol_len, n_new in CASES + [(11, 100000, 70000, 9000), (12, 63, 10, 3), (13, 65, 10, 3)]:
        b, t = make_ctx(seed, n_words, ol_len, n_new)
        items.append((b, t, ol_len, ol_len + n_new > 1024))
    got = eng.b250_generate(items)
    for (b, t, ol, up), (g, nw) in zip(items, got):
        w = orc.b250_generate(b, t, ol, up, "port")
        assert nw == w[1] and np.array_equal(g, w[0]), (b.size, g[:16], w[0][:16])
        if orc.have_gz_ref():
            assert np.array_equal(g, orc.b250_generate(b, t, ol, up, "ref")[0])
    from genozip_b200.lib import GzbError
    b, t = make_ctx(9, 500, 3000, 10)
    with pytest.raises(GzbError):
        eng.b250_generate([(np.concatenate([[0x80], b]).astype(np.uint8), t, 3000, True)])
