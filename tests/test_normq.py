"""NORMQ (reference src/codec_normq.c): the restatement pinned against the reference's compiled codec_normq.c (oracle/_ref), and the
CUDA path (gzb_normq_gather / gzb_normq_reconstruct through the C-ABI) against both: ragged and empty lines, reverse-complemented
reads, SAM lines without quality (' ' in the stream, '*' in the text) at the start, in runs and at the end, a stream that does not
fit its lines."""
import numpy as np, pytest
import orc


def _vb(seed, n_lines=300, max_len=200, p_rev=0.4, p_missing=0.0, p_empty=0.05):
    """(txt, off, zip-side lengths, is_rev, seq_len per line, missing flags): qualities '!'..'~'; a line without quality is the one byte ' '"""
    rng = np.random.default_rng(seed)
    seq_len = rng.integers(1, max_len + 1, n_lines).astype(np.uint32)
    seq_len[rng.random(n_lines) < p_empty] = 0
    missing = (rng.random(n_lines) < p_missing) & (seq_len > 0)
    zlen = np.where(missing, 1, seq_len).astype(np.uint32)
    parts, off, pos = [], [], 0
    for i in range(n_lines):
        gap = int(rng.integers(0, 7))                                       # the lines lie anywhere in the text (vb->txt_data)
        parts.append(rng.integers(33, 127, gap, dtype=np.uint8)); pos += gap
        off.append(pos)
        q = np.full(1, 32, np.uint8) if missing[i] else rng.integers(33, 127, int(zlen[i]), dtype=np.uint8)
        parts.append(q); pos += q.size
    txt = np.concatenate(parts + [np.zeros(1, np.uint8)])
    is_rev = (rng.random(n_lines) < p_rev).astype(np.uint8)
    return txt, np.asarray(off, np.uint64), zlen, is_rev, seq_len, missing.astype(np.uint8)


def _text_of(out, lens, miss):
    """what the reference leaves in txt_data: every line's bytes, a line without quality as the one character '*'"""
    parts, pos = [], 0
    for L, m in zip(lens, miss):
        parts.append(out[pos:pos + 1] if m else out[pos:pos + L]); pos += int(L)
    return np.concatenate(parts) if parts else np.zeros(0, np.uint8)


CASES = [dict(seed=1), dict(seed=2, p_missing=0.1), dict(seed=3, p_missing=0.6, p_rev=0.9), dict(seed=4, n_lines=1, p_empty=0), dict(seed=5, n_lines=2000, max_len=30, p_missing=0.02),
         dict(seed=6, p_missing=1.0, p_empty=0), dict(seed=7, n_lines=40, max_len=5000)]


@pytest.mark.parametrize("kw", CASES)
def test_restatement_is_the_reference(kw):
    if not orc.have_gz_ref():
        pytest.skip("oracle/_ref/libgz_ref.so not built here")
    txt, off, zlen, rev, seq_len, missing = _vb(**kw)
    local = orc.normq_encode(txt, off, zlen, rev)
    assert np.array_equal(local, orc.ref_normq_encode(txt, off, zlen, rev))
    out, miss = orc.normq_decode(local, seq_len, rev)
    assert np.array_equal(miss, missing)
    assert np.array_equal(_text_of(out, seq_len, miss), orc.ref_normq_decode(local, seq_len, rev))


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_normq_batch(eng):
    vbs = [_vb(**kw) for kw in CASES]
    locals_ = eng.normq_gather([(t, o, z, r) for t, o, z, r, _, _ in vbs])
    for (t, o, z, r, sl, ms), got in zip(vbs, locals_):
        want = orc.ref_normq_encode(t, o, z, r) if orc.have_gz_ref() else orc.normq_encode(t, o, z, r)
        assert np.array_equal(got, want)
    outs = eng.normq_reconstruct([(loc, sl, r) for (_, _, _, r, sl, _), loc in zip(vbs, locals_)])
    for (t, o, z, r, sl, ms), loc, (out, miss) in zip(vbs, locals_, outs):
        assert np.array_equal(miss, ms)
        want = orc.ref_normq_decode(loc, sl, r) if orc.have_gz_ref() else _text_of(*orc.normq_decode(loc, sl, r)[:1], sl, ms)
        assert np.array_equal(_text_of(out, sl, miss), want)
    # is_rev NULL and an empty VBlock
    t, o, z, r, sl, ms = vbs[0]
    assert np.array_equal(eng.normq_gather([(t, o, z, None)])[0], orc.normq_encode(t, o, z, None))
    assert eng.normq_gather([(np.zeros(1, np.uint8), np.zeros(0, np.uint64), np.zeros(0, np.uint32), None)])[0].size == 0


@pytest.mark.gpu
def test_gpu_normq_refuses_a_stream_that_does_not_fit(eng):
    from genozip_b200 import GzbError
    t, o, z, r, sl, ms = _vb(seed=9, p_missing=0.05)
    loc = orc.normq_encode(t, o, z, r)
    for bad in (loc[:-3], np.concatenate([loc, loc[:5]])):
        assert orc.normq_decode(bad, sl, r) is None
        with pytest.raises(GzbError):
            eng.normq_reconstruct([(bad, sl, r)])
