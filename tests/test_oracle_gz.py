"""CPU tests of the restatement of genozip's own codecs (oracle/gz_port.c): every encoder is round-tripped through
the independently restated reference DECODER, plus stream-format properties stated in the reference source."""
import numpy as np, pytest
import orc
from datagen import fastq_vb, line_table, ragged_quals, haplotype_matrix, longread_vb


def test_acgt_roundtrip_and_layout():
    seq, _ = fastq_vb(300, 151, 1, lower_frac=0.01, n_frac=0.01)
    packed, x, allz = orc.acgt_pack(seq)
    assert packed.size == ((2 * seq.size + 63) // 64) * 8 and not allz
    assert np.array_equal(orc.acgt_unpack(packed, x, seq.size), seq)
    # base i lives in bits [2i,2i+1] of LE words (codec_acgt.c:45-55)
    s2 = np.frombuffer(b"ACGTTGCA", np.uint8)
    p2, x2, z2 = orc.acgt_pack(s2)
    assert z2 and p2[0] == (0 | 1 << 2 | 2 << 4 | 3 << 6) and p2[1] == (3 | 2 << 2 | 1 << 4 | 0 << 6)
    assert np.array_equal(orc.acgt_unpack(p2, None, 8), s2)
    for n in (0, 1, 31, 32, 33, 63, 64, 65):
        s = seq[:n]
        p, x, _ = orc.acgt_pack(s)
        assert np.array_equal(orc.acgt_unpack(p, x, n), s)


@pytest.mark.parametrize("seed", range(6))
def test_domq_roundtrip_ragged(seed):
    txt, off, lens = ragged_quals(seed)
    enc = orc.domq_encode(txt, off, lens)
    out = orc.domq_decode(enc, lens)
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(off, lens)]) if lens.sum() else np.zeros(0, np.uint8)
    assert np.array_equal(out, want)
    assert enc["mplx"].size == int((lens > 0).sum())


def test_domq_fastq_and_edge_cases():
    _, qual = fastq_vb(2000, 150, 3)
    off, lens = line_table(2000, 150)
    enc = orc.domq_encode(qual, off, lens)
    assert np.array_equal(orc.domq_decode(enc, lens), qual)
    assert enc["qual"].size < qual.size // 3            # DOMQ's point: few non-dom values remain
    # all-dom VB: QUAL.local is the single byte 'X', no runs (codec_domq.c:473-500)
    q = np.full(10 * 100, ord("F"), np.uint8)
    off, lens = line_table(10, 100)
    enc = orc.domq_encode(q, off, lens)
    assert enc["qual"].tobytes() == b"X" and enc["runs"].size == 0
    assert np.array_equal(orc.domq_decode(enc, lens), q)
    # runs of exactly 254/255/508 doms: 254 -> [254]; 255 -> [255,1]; 508 -> [255,254] (codec_domq.c:368-377)
    for r, want in ((254, [254]), (255, [255, 1]), (508, [255, 254]), (509, [255, 255, 1])):
        q = np.concatenate([np.full(r, ord("F"), np.uint8), [ord("#")], np.full(600 - r - 1, ord("F"), np.uint8)]).astype(np.uint8)
        off, lens = line_table(1, 600)
        enc = orc.domq_encode(q, off, lens)
        assert list(enc["runs"][:len(want)]) == want
        assert np.array_equal(orc.domq_decode(enc, lens), q)


@pytest.mark.parametrize("multi", [False, True])
def test_pbwt_roundtrip(multi):
    ht = haplotype_matrix(200, 50, 4, multi=multi)
    runs, fgrc = orc.pbwt_encode(ht)
    assert runs.sum() == ht.size                          # run lengths tile the matrix
    assert int(fgrc[-2]) | int(fgrc[-1]) << 32 == ht.size  # trailing 64-bit matrix length (codec_pbwt.c:274-276)
    back = orc.pbwt_decode(runs, fgrc, ht.shape[0], ht.size)
    assert np.array_equal(back.reshape(ht.shape), ht)


@pytest.mark.parametrize("rev", [False, True])
def test_longr_roundtrip(rev):
    seq, qual, lens = longread_vb(12, 3000, 5)
    n = int(lens.sum())
    txt = np.concatenate([seq, qual])
    seq_off = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    qual_off = (seq_off + np.uint64(n)).astype(np.uint64)
    is_rev = (np.arange(lens.size) % 2).astype(np.uint8) if rev else None
    v2b = orc.longr_bins(qual)
    values, lens_be = orc.longr_encode(txt, seq_off, qual_off, lens, is_rev, v2b)
    assert int(lens_be.byteswap().sum()) == n
    assert np.array_equal(np.sort(values), np.sort(qual - 33))
    back = orc.longr_decode(txt, seq_off, lens, is_rev, v2b, values, lens_be)
    assert np.array_equal(back, qual)
