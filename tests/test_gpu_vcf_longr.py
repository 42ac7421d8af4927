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
