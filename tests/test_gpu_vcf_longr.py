"""GPU parity tests for PBWT and LONGR: CUDA path through the C-ABI vs the CPU restatement (oracle/gz_port.c) and, where
oracle/_ref/libgz_ref.so travelled, vs the reference's own compiled codec_pbwt.c / codec_longr.c — word- and byte-exact —
and back through both decoders."""
import numpy as np, pytest
import orc
from datagen import haplotype_matrix, longread_vb

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("shape,multi", [((200, 50), False), ((300, 77), True), ((64, 1000), False), ((3, 5), False), ((1, 40), True), ((40, 10000), False)])
def test_pbwt(eng, shape, multi):
    n_lines, n_samples = shape
    ht = haplotype_matrix(n_lines, n_samples, 4 + n_lines, multi=multi)
    runs, fgrc = eng.pbwt_encode(ht)
    wr, wf = orc.pbwt_encode(ht)
    assert runs.size == wr.size and np.array_equal(runs, wr), "RUNS differ from the oracle"
    assert fgrc.size == wf.size and np.array_equal(fgrc, wf), "FGRC differ from the oracle"
    if orc.have_gz_ref():
        rr, rf = orc.ref_pbwt_encode(ht)
        assert np.array_equal(runs, rr) and np.array_equal(fgrc, rf), "differs from the reference's compiled codec_pbwt.c"
    back = eng.pbwt_decode(wr, wf, ht.shape[0], ht.size)
    assert np.array_equal(back.reshape(ht.shape), ht)


def test_pbwt_batch(eng):
    """several VBlocks' matrices of different shapes in one call (one CTA each), both directions"""
    hts = [haplotype_matrix(nl, ns, 90 + nl, multi=m) for nl, ns, m in ((120, 60, False), (1, 7, True), (77, 333, True), (300, 1000, False), (5, 9000, False))]
    got = eng.pbwt_encode_batch(hts)
    want = [orc.ref_pbwt_encode(h) if orc.have_gz_ref() else orc.pbwt_encode(h) for h in hts]
    for (r, f), (wr, wf) in zip(got, want):
        assert np.array_equal(r, wr) and np.array_equal(f, wf)
    back = eng.pbwt_decode_batch(want, [h.shape[0] for h in hts], [h.size for h in hts])
    for b, h in zip(back, hts):
        assert np.array_equal(b.reshape(h.shape), h)


def test_pbwt_capacity_is_reported(eng):
    ht = haplotype_matrix(50, 40, 3)
    from genozip_b200 import GzbError
    with pytest.raises(GzbError):
        eng.pbwt_encode_batch([ht], runs_cap=8, fgrc_cap=64)


def test_pbwt_corrupt_runs_are_refused(eng):
    ht = haplotype_matrix(30, 20, 4)
    r, f = orc.pbwt_encode(ht)
    from genozip_b200 import GzbError
    with pytest.raises(GzbError):
        eng.pbwt_decode_batch([(r[: r.size // 2], f)], [30], [ht.size])          # the runs no longer cover the matrix


@pytest.mark.fullsize
def test_pbwt_config_size(eng):
    """BASELINE configs[3]: 38 000 lines x 2000 haplotypes per VBlock, vs the reference's compiled codec_pbwt.c, both directions"""
    import torch, bench_domain
    ht = bench_domain.synth_vcf_vb(38000, 2000, 4242, torch.device("cpu")).numpy()
    (r, f), = eng.pbwt_encode_batch([ht], runs_cap=ht.size // 4, fgrc_cap=ht.size // 8)
    wr, wf = orc.ref_pbwt_encode(ht) if orc.have_gz_ref() else orc.pbwt_encode(ht)
    assert np.array_equal(r, wr) and np.array_equal(f, wf)
    back, = eng.pbwt_decode_batch([(wr, wf)], [38000], [ht.size])
    assert np.array_equal(back.reshape(ht.shape), ht)
    if orc.have_gz_ref():
        assert np.array_equal(orc.ref_pbwt_decode(r, f, 38000, ht.size).reshape(ht.shape), ht)


def _longr_vb(seed, n_reads, mean_len, rev):
    seq, qual, lens = longread_vb(n_reads, mean_len, seed)
    n = int(lens.sum())
    txt = np.concatenate([seq, qual])
    seq_off = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    qual_off = (seq_off + np.uint64(n)).astype(np.uint64)
    is_rev = (np.arange(lens.size) % 2).astype(np.uint8) if rev else None
    return (txt, seq_off, qual_off, lens, is_rev, orc.longr_bins(qual)), qual


@pytest.mark.parametrize("rev", [False, True])
def test_longr(eng, rev):
    vbs, quals = zip(*[_longr_vb(5 + s, 10 + s, 2500, rev) for s in range(3)])
    got = eng.longr_encode(list(vbs))
    for vb, (vals, lb) in zip(vbs, got):
        wv, wl = orc.longr_encode(vb[0], vb[1], vb[2], vb[3], vb[4], vb[5])
        assert np.array_equal(lb, wl), "channel lengths differ from the oracle"
        assert np.array_equal(vals, wv), "sorted values differ from the oracle"
        if orc.have_gz_ref():
            rb, rv, rl = orc.ref_longr_encode(vb[0], vb[1], vb[2], vb[3], vb[4])
            assert np.array_equal(rb, vb[5]) and np.array_equal(vals, rv) and np.array_equal(lb, rl), "differs from the reference's compiled codec_longr.c"
    back = eng.longr_decode(list(vbs), [g[0] for g in got], [g[1] for g in got])
    for q, b in zip(quals, back):
        assert np.array_equal(b, q)


def test_longr_bins(eng):
    """codec_longr_segconf_calculate_bins: histogram on the GPU, bins on the host"""
    for seed in (1, 2):
        vb, qual = _longr_vb(seed, 20, 1500, False)
        assert np.array_equal(eng.longr_calculate_bins(vb), vb[5])
    if orc.have_gz_ref():
        vb, qual = _longr_vb(9, 30, 900, True)
        rb, _, _ = orc.ref_longr_encode(vb[0], vb[1], vb[2], vb[3], vb[4])
        assert np.array_equal(eng.longr_calculate_bins(vb), rb)


def test_longr_missing_quality(eng):
    """a SAM line without quality is the single byte ' ' (value 255) whatever its seq_len (codec_longr.c:188-192, :278):
    the encoder takes one value from it, the decoder stops the line there and flags it"""
    vb, qual = _longr_vb(21, 12, 800, True)
    txt, seq_off, qual_off, lens, is_rev, v2b = vb
    txt = txt.copy(); qlens = lens.copy()
    for li in (0, 5, 11):
        txt[int(qual_off[li])] = ord(" "); qlens[li] = 1
    v2b = eng.longr_calculate_bins((txt, seq_off, qual_off, lens, is_rev, v2b, qlens))     # the lines without quality do not count (:88)
    vb2 = (txt, seq_off, qual_off, lens, is_rev, v2b, qlens)
    (vals, lb), = eng.longr_encode([vb2])
    wv, wl = orc.longr_encode(txt, seq_off, qual_off, qlens, is_rev, v2b, seq_lens=lens)
    assert np.array_equal(vals, wv) and np.array_equal(lb, wl)
    if orc.have_gz_ref():
        rb, rv, rl = orc.ref_longr_encode(txt, seq_off, qual_off, qlens, is_rev, seq_lens=lens)
        assert np.array_equal(rb, v2b) and np.array_equal(vals, rv) and np.array_equal(lb, rl), "differs from the reference's compiled codec_longr.c"
    back, = eng.longr_decode([vb2], [vals], [lb])
    miss = eng.longr_missing[0]
    assert list(np.nonzero(miss)[0]) == [0, 5, 11]
    off = np.concatenate([[0], np.cumsum(lens[:-1])]).astype(np.int64)
    n = int(lens.sum())
    for li in range(lens.size):
        got = back[off[li]: off[li] + lens[li]]
        if miss[li]:
            assert got[0] == ord("*")
        else:
            assert np.array_equal(got, txt[n + off[li]: n + off[li] + lens[li]])


def test_longr_corrupt_lengths_are_refused(eng):
    vb, qual = _longr_vb(31, 6, 500, False)
    (vals, lb), = eng.longr_encode([vb])
    bad = lb.copy(); bad[int(np.nonzero(lb)[0][0])] = np.uint32(1 << 24).byteswap()      # channel lengths no longer add up
    from genozip_b200 import GzbError
    with pytest.raises(GzbError):
        eng.longr_decode([vb], [vals], [bad])


@pytest.mark.fullsize
@pytest.mark.parametrize("rev", [False, True])
def test_longr_config_size(eng, rev):
    """BASELINE configs[4]: 50 kb reads, 8 M qualities in the VBlock, vs the reference's compiled codec_longr.c, both directions"""
    import torch, bench_domain
    txt, so, qo, ln = [t.numpy() for t in bench_domain.synth_longread_vb(8_000_000, 50000, 77 + rev, torch.device("cpu"))]
    so, qo, ln = so.astype(np.uint64), qo.astype(np.uint64), ln.astype(np.uint32)
    is_rev = (np.arange(ln.size) % 3 == 1).astype(np.uint8) if rev else None
    if orc.have_gz_ref():
        v2b, wv, wl = orc.ref_longr_encode(txt, so, qo, ln, is_rev)
    else:
        n = int(ln.sum()); v2b = orc.longr_bins(txt[n:]); wv, wl = orc.longr_encode(txt, so, qo, ln, is_rev, v2b)
    vb = (txt, so, qo, ln, is_rev, v2b)
    assert np.array_equal(eng.longr_calculate_bins(vb), v2b)
    (vals, lb), = eng.longr_encode([vb])
    assert np.array_equal(lb, wl) and np.array_equal(vals, wv)
    back, = eng.longr_decode([vb], [wv], [wl])
    n = int(ln.sum())
    assert np.array_equal(back, txt[n:2 * n])
