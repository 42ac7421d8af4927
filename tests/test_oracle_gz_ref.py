"""Pins the CPU restatement of genozip's own codecs (oracle/gz_port.c) against the REFERENCE's own compiled translation units
codec_domq.c, codec_acgt.c, codec_pbwt.c, codec_longr.c (oracle/_ref/libgz_ref.so: unmodified, hosted by oracle/ref_gz_shim.c
with a hand-made VBlock; nucleotide tables from the reference's compiled reference.c):
  DOMQ   the four streams, the de-normalisation table and the section parameter — FASTQ-like VBlocks, ragged and empty lines,
         all-dominant and all-diverse VBlocks, dom runs across lines, the 254/255 run-length escapes, ties between qualities
         (dom choice and libc qsort order of the rank tables);
  ACGT   the 2-bit words handed to the sub-codec, the exception stream, acgt_no_x — IUPAC codes, lower case, odd characters;
  PBWT   RUNS and FGRC — bi- and multi-allelic matrices incl. the pseudo alleles;
  LONGR  the value-to-bin map, the channel-sorted values and the 65,536 big-endian channel lengths — forward and
         reverse-complemented reads.
The PIZ side is pinned the same way: the reference's codec_acgt_uncompress / codec_xcgt_uncompress, codec_pbwt_uncompress,
codec_domq_reconstruct (line by line) and codec_longr_reconstruct (read by read) must reproduce the input from the streams, and
the restated decoders must agree with them byte for byte (incl. on streams the reference encoder produced)."""
import numpy as np, pytest
import orc
from datagen import fastq_vb, line_table, ragged_quals, haplotype_matrix, longread_vb

pytestmark = pytest.mark.skipif(not orc.have_gz_ref(), reason="oracle/_ref/libgz_ref.so not built and /root/reference absent")


def check(txt, off, lens):
    r = orc.ref_domq_encode(txt, off, lens)
    w = orc.domq_encode(txt, off, lens)
    for k in ("qual", "runs", "mplx", "divr", "denorm"):
        assert r[k].size == w[k].size and np.array_equal(r[k], w[k]), f"{k}: restatement != reference (len {w[k].size} vs {r[k].size})"
    assert r["num_norm_qs"] == w["num_norm_qs"] and bool(r["has_diverse"]) == bool(w["has_diverse"])
    # PIZ side: the reference's codec_domq_reconstruct and the restated decoder, both on the reference's streams
    lens = np.asarray(lens, np.uint32)
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(off, lens)]) if lens.sum() else np.zeros(0, np.uint8)
    assert np.array_equal(orc.ref_domq_decode(r, lens), want), "reference codec_domq_reconstruct does not give back the input"
    assert np.array_equal(orc.domq_decode(r, lens), want), "restated DOMQ decoder != reference"


@pytest.mark.parametrize("n_reads,read_len,seed", [(200, 150, 1), (3000, 151, 2), (50, 37, 3), (1000, 100, 4), (20000, 150, 5), (7, 1, 6), (1, 150, 7)])
def test_fastq_like(n_reads, read_len, seed):
    _, qual = fastq_vb(n_reads, read_len, seed)
    off, lens = line_table(n_reads, read_len)
    check(qual, off, lens)


@pytest.mark.parametrize("seed", range(8))
def test_ragged_and_empty_lines(seed):
    check(*ragged_quals(seed))


def test_run_length_escapes_and_extremes():
    for r in (1, 253, 254, 255, 256, 508, 509, 762, 763, 2000):
        q = np.concatenate([np.full(r, ord("F"), np.uint8), [ord("#")], np.full(2600 - r - 1, ord("F"), np.uint8)]).astype(np.uint8)
        check(q, *line_table(1, 2600))
        check(q, *line_table(26, 100))                                     # the same text as 26 lines: runs span lines
    q = np.full(10 * 100, ord("F"), np.uint8)
    check(q, *line_table(10, 100))                                         # all dominant: QUAL.local = 'X'
    rng = np.random.default_rng(3)
    q = rng.integers(33, 75, 5000).astype(np.uint8)
    check(q, *line_table(50, 100))                                         # all diverse
    q = np.tile(np.frombuffer(b"FFFF::::", np.uint8), 500)                  # ties: two qualities equally frequent in every line
    check(q, *line_table(40, 100))
    q = np.concatenate([np.full(100, ord("F"), np.uint8), np.full(100, ord(","), np.uint8), np.full(100, ord("F"), np.uint8)])
    check(q, *line_table(3, 100))                                          # different dom per line, final run


def test_many_qualities_and_rank_ties():
    rng = np.random.default_rng(11)
    for t in range(6):
        n_lines, ln = 300, 120
        doms = rng.choice(np.arange(40, 80), 3, replace=False)
        q = np.empty(n_lines * ln, np.uint8)
        for i in range(n_lines):
            d = doms[i % 3]
            line = np.full(ln, d, np.uint8)
            k = rng.integers(0, 40)
            line[rng.integers(0, ln, k)] = rng.integers(33, 127, k)      # many distinct rare qualities: equal counts -> qsort tie order
            q[i * ln:(i + 1) * ln] = line
        check(q, *line_table(n_lines, ln))


# ------------------------------------------------------------------------------------------------ ACGT
def test_acgt_against_reference():
    seq, _ = fastq_vb(300, 151, 1, lower_frac=0.01, n_frac=0.01)
    pure = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(2).integers(0, 4, 100000)].copy()
    odd = np.frombuffer(b"ACGTNacgtnRYSWKMBDHVUryswkmbdhvu*-.", np.uint8).copy()
    allb = np.arange(256, dtype=np.uint8)                                   # every byte value
    for s in [seq, pure, odd, allb] + [seq[:n].copy() for n in (1, 5, 31, 32, 33, 63, 64, 65, 1000, 4097)]:
        p, x, nox = orc.ref_acgt_pack(s)
        pw, xw, zw = orc.acgt_pack(s)
        assert np.array_equal(p, pw) and np.array_equal(x, xw) and nox == zw, f"n={s.size}"
        back = orc.ref_acgt_unpack(p, None if nox else x, s.size)           # codec_acgt_uncompress [+ codec_xcgt_uncompress]
        lossless = s > 1                                                    # the format itself maps bytes 0 and 1 to 'A' and 'a' (exception codes 0 / 1)
        assert np.array_equal(back[lossless], s[lossless]) and set(back[~lossless]) <= {65, 97}, f"n={s.size}: reference unpack does not give back the input"
        assert np.array_equal(orc.acgt_unpack(p, None if nox else x, s.size), back), f"n={s.size}: restated unpack != reference"
        if nox:                                                             # an all-zero exception stream given explicitly is a no-op
            assert np.array_equal(orc.ref_acgt_unpack(p, x, s.size), back)


# ------------------------------------------------------------------------------------------------ PBWT
@pytest.mark.parametrize("n_lines,n_samples,multi", [(50, 40, False), (50, 40, True), (200, 1000, False), (200, 1000, True), (7, 3, False), (300, 17, True), (1, 5, False)])
def test_pbwt_against_reference(n_lines, n_samples, multi):
    ht = haplotype_matrix(n_lines, n_samples, n_lines + n_samples, multi=multi)
    r, f = orc.ref_pbwt_encode(ht)
    rw, fw = orc.pbwt_encode(ht)
    assert r.size == rw.size and np.array_equal(r, rw) and f.size == fw.size and np.array_equal(f, fw)
    back = orc.ref_pbwt_decode(r, f, n_lines, ht.size)                      # codec_pbwt_uncompress
    assert np.array_equal(back, ht.reshape(-1)), "reference codec_pbwt_uncompress does not give back the matrix"
    assert np.array_equal(orc.pbwt_decode(r, f, n_lines, ht.size), back), "restated PBWT decoder != reference"


def test_pbwt_pseudo_alleles_against_reference():
    rng = np.random.default_rng(5)
    alleles = np.frombuffer(b"0011122.*%-&", np.uint8)
    ht = alleles[rng.integers(0, alleles.size, (60, 48))]
    r, f = orc.ref_pbwt_encode(ht)
    rw, fw = orc.pbwt_encode(ht)
    assert np.array_equal(r, rw) and np.array_equal(f, fw)
    back = orc.ref_pbwt_decode(r, f, 60, ht.size)
    assert np.array_equal(back, ht.reshape(-1)) and np.array_equal(orc.pbwt_decode(r, f, 60, ht.size), back)


# ------------------------------------------------------------------------------------------------ LONGR
@pytest.mark.parametrize("n_reads,mean_len,seed,rev", [(12, 3000, 5, False), (12, 3000, 5, True), (40, 500, 6, True), (3, 20, 7, False), (200, 150, 8, True)])
def test_longr_against_reference(n_reads, mean_len, seed, rev):
    seq, qual, lens = longread_vb(n_reads, mean_len, seed)
    n = int(lens.sum())
    txt = np.concatenate([seq, qual])
    seq_off = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64); qual_off = seq_off + np.uint64(n)
    is_rev = (np.random.default_rng(seed).random(n_reads) < 0.5).astype(np.uint8) if rev else None
    v2b_ref, values_ref, lens_ref = orc.ref_longr_encode(txt, seq_off, qual_off, lens, is_rev)
    v2b = orc.longr_bins(qual)
    values, lens_be = orc.longr_encode(txt, seq_off, qual_off, lens, is_rev, v2b)
    assert np.array_equal(v2b, v2b_ref), "value_to_bin (codec_longr_segconf_calculate_bins)"
    assert np.array_equal(values, values_ref) and np.array_equal(lens_be, lens_ref)
    back = orc.ref_longr_decode(txt, seq_off, lens, is_rev, v2b_ref, values_ref, lens_ref)   # codec_longr_reconstruct
    assert np.array_equal(back, qual), "reference codec_longr_reconstruct does not give back QUAL"
    assert np.array_equal(orc.longr_decode(txt, seq_off, lens, is_rev, v2b, values, lens_be), back), "restated LONGR decoder != reference"
