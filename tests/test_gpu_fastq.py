"""GPU parity tests for the FASTQ-side genozip codecs: ACGT/XCGT packing and DOMQ — CUDA path through the C-ABI vs the
CPU restatement (oracle/gz_port.c), byte-exact both ways, incl. ragged/empty lines and the all-dom / run-length edge cases."""
import numpy as np, pytest
import orc
from datagen import fastq_vb, line_table, ragged_quals

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def test_acgt(eng):
    seq, _ = fastq_vb(3000, 151, 1, lower_frac=0.01, n_frac=0.01)
    for n in (0, 1, 5, 31, 32, 33, 63, 64, 65, 1000, 4097, seq.size):
        s = seq[:n].copy()
        p, x, allz = eng.acgt_pack(s)
        pw, xw, zw = orc.acgt_pack(s)
        assert np.array_equal(p, pw) and np.array_equal(x, xw) and allz == zw, f"n={n}"
        if orc.have_gz_ref() and n:                                         # the reference's own compiled codec_acgt.c
            pr, xr, zr = orc.ref_acgt_pack(s)
            assert np.array_equal(p, pr) and np.array_equal(x, xr) and allz == zr, f"n={n}: GPU != reference codec_acgt.c"
        assert np.array_equal(eng.acgt_unpack(pw, None if zw else xw, n), s)
    # pure ACGT: exception stream all zero -> acgt_no_x (codec_acgt.c:136-140)
    s = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(2).integers(0, 4, 100000)].copy()
    p, x, allz = eng.acgt_pack(s)
    assert allz and not x.any() and np.array_equal(p, orc.acgt_pack(s)[0])
    assert np.array_equal(eng.acgt_unpack(p, None, s.size), s)
    # IUPAC and odd characters
    s = np.frombuffer(b"ACGTNacgtnRYSWKMBDHVUryswkmbdhvu*-.", np.uint8).copy()
    p, x, allz = eng.acgt_pack(s)
    pw, xw, zw = orc.acgt_pack(s)
    assert np.array_equal(p, pw) and np.array_equal(x, xw)
    assert np.array_equal(eng.acgt_unpack(p, x, s.size), s)


def test_acgt_batch(eng):
    """gzb_acgt_pack_batch / gzb_acgt_unpack_batch: ragged VBlocks (incl. empty, all-ACGT) in one launch vs the oracle"""
    seq, _ = fastq_vb(3000, 151, 3, lower_frac=0.01, n_frac=0.01)
    pure = np.frombuffer(b"ACGT", np.uint8)[np.random.default_rng(4).integers(0, 4, 70001)].copy()
    seqs = [seq[:n].copy() for n in (0, 1, 33, 4097, 100000, seq.size)] + [pure]
    got = eng.acgt_pack_batch(seqs)
    items = []
    for s, (p, x, allz) in zip(seqs, got):
        pw, xw, zw = orc.acgt_pack(s)
        assert np.array_equal(p, pw) and np.array_equal(x, xw) and allz == zw, f"n={s.size}"
        items.append((pw, None if zw else xw, s.size))
    for s, o in zip(seqs, eng.acgt_unpack_batch(items)):
        assert np.array_equal(o, s), f"n={s.size}"


def _check_domq(eng, vbs):
    got = eng.domq_encode(vbs)
    for (txt, off, ln), g in zip(vbs, got):
        w = orc.domq_encode(txt, off, ln)
        for k in ("num_norm_qs", "num_doms", "has_diverse"):
            assert g[k] == w[k], k
        for k in ("denorm", "line_dom", "line_diverse", "qual", "runs", "mplx", "divr"):
            assert g[k].size == w[k].size and np.array_equal(g[k], w[k]), f"{k}: GPU != oracle (len {g[k].size} vs {w[k].size})"
        if orc.have_gz_ref() and ln.size:                                   # the reference's own compiled codec_domq.c
            r = orc.ref_domq_encode(txt, off, ln)
            for k in ("denorm", "qual", "runs", "mplx", "divr"):
                assert np.array_equal(g[k], r[k]), f"{k}: GPU != reference codec_domq.c"
    back = eng.domq_decode(got, [v[2] for v in vbs])
    for (txt, off, ln), b, g in zip(vbs, back, got):
        want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(off, ln)]) if ln.sum() else np.zeros(0, np.uint8)
        assert np.array_equal(b, want), "GPU DOMQ reconstruct mismatch"
        assert np.array_equal(orc.domq_decode(g, ln), want)


def test_domq_fastq_batch(eng):
    vbs = []
    for s in range(4):
        _, q = fastq_vb(5000 + 1000 * s, 150, 10 + s)
        off, ln = line_table(5000 + 1000 * s, 150)
        vbs.append((q, off, ln))
    _check_domq(eng, vbs)


def test_domq_ragged(eng):
    _check_domq(eng, [ragged_quals(s) for s in range(6)])


def test_domq_edges(eng):
    vbs = []
    q = np.full(10 * 100, ord("F"), np.uint8)              # all dom -> QUAL.local = 'X', no runs
    vbs.append((q,) + line_table(10, 100))
    for r in (254, 255, 508, 509, 5000):                    # run-length byte boundaries (codec_domq.c:368-377)
        L = max(600, r + 10)
        q = np.concatenate([np.full(r, ord("F"), np.uint8), [ord("#")], np.full(L - r - 1, ord("F"), np.uint8)]).astype(np.uint8)
        vbs.append((q,) + line_table(1, L))
    q = np.concatenate([np.frombuffer(b"#:#:", np.uint8), np.full(96, ord("F"), np.uint8)])   # leading non-doms then only doms
    vbs.append((q.copy(),) + line_table(1, 100))
    q2 = np.tile(np.frombuffer(b"F#", np.uint8), 3000)       # alternating: no run ever
    vbs.append((q2.copy(),) + line_table(40, 150))
    _check_domq(eng, vbs)


def test_domq_tiles_and_long_lines(eng):
    """The split works on 16384-element tiles (64-element chunks per thread) of the non-diverse concatenation and the histogram pass
    gives lines above 4096 qualities to a whole warp: runs that cross one or several tile borders, non-doms on either side of a border, a tile of doms only, a trailing
    run longer than a tile, long lines beside short ones (codec_domq.c:139-178, 421-500)."""
    rng = np.random.default_rng(77)
    vbs = []
    for case in range(6):
        n = 40000
        q = np.full(n, ord("F"), np.uint8)
        if case == 0:   pos = [63, 64, 127, 129, 16383, 16384, 32767, 32769, 39000]   # non-doms at chunk and tile borders
        elif case == 1: pos = [0, 36000]                                        # one run across two tile borders, then a trailing one
        elif case == 2: pos = [16384 * 2 - 1]                                   # a run ending exactly at a tile's last element
        elif case == 3: pos = list(range(16380, 16390)) + list(range(60, 70)) + [32768]   # literal strings across the borders
        elif case == 4: pos = sorted(rng.choice(n, 300, replace=False).tolist())
        else:           pos = [n - 1]                                           # no trailing run
        q[pos] = rng.choice(np.frombuffer(b"#,:", np.uint8), len(pos))
        vbs.append((q,) + line_table(n // 200, 200))
    # long lines (one warp each) between short ones, one of them diverse, one with a different dom
    lens = [150, 5000, 150, 256, 9000, 257, 4096, 4097, 300, 20000, 1000, 1, 255]
    parts = []
    for i, L in enumerate(lens):
        if i == 4:   line = rng.choice(np.frombuffer(b"#,:F", np.uint8), L)                         # diverse
        elif i == 7: line = np.where(rng.random(L) < 0.9, ord(":"), ord("F")).astype(np.uint8)       # dom ':'
        else:        line = np.where(rng.random(L) < 0.93, ord("F"), ord(",")).astype(np.uint8)
        parts.append(line.astype(np.uint8))
    off = np.cumsum([0] + lens[:-1]).astype(np.uint64)
    vbs.append((np.concatenate(parts), off, np.array(lens, np.uint32)))
    _check_domq(eng, vbs)


def test_domq_full_vb_properties(eng):
    """BASELINE-size VBlock (92K reads x 150): round trip + stream-length identities instead of the slow oracle decode"""
    _, q = fastq_vb(92000, 150, 77)
    off, ln = line_table(92000, 150)
    g = eng.domq_encode([(q, off, ln)])[0]
    w = orc.domq_encode(q, off, ln)
    for k in ("qual", "runs", "mplx", "divr", "denorm"):
        assert np.array_equal(g[k], w[k]), k
    back = eng.domq_decode([g], [ln])[0]
    assert np.array_equal(back, q)
