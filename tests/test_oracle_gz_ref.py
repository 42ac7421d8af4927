"""Pins the CPU restatement of DOMQ (oracle/gz_port.c) against the REFERENCE's own compiled codec_domq.c
(oracle/_ref/libgz_ref.so: the unmodified translation unit hosted by oracle/ref_gz_shim.c with a hand-made VBlock):
the four streams, the de-normalisation table and the section parameter must be byte-identical — FASTQ-like VBlocks,
ragged lines, empty lines, all-dominant and all-diverse VBlocks, dom runs across lines, the 254/255 run-length escapes,
ties between qualities (dom choice and qsort order of the rank tables)."""
import numpy as np, pytest
import orc
from datagen import fastq_vb, line_table, ragged_quals

pytestmark = pytest.mark.skipif(not orc.have_gz_ref(), reason="oracle/_ref/libgz_ref.so not built and /root/reference absent")


def check(txt, off, lens):
    r = orc.ref_domq_encode(txt, off, lens)
    w = orc.domq_encode(txt, off, lens)
    for k in ("qual", "runs", "mplx", "divr", "denorm"):
        assert r[k].size == w[k].size and np.array_equal(r[k], w[k]), f"{k}: restatement != reference (len {w[k].size} vs {r[k].size})"
    assert r["num_norm_qs"] == w["num_norm_qs"] and bool(r["has_diverse"]) == bool(w["has_diverse"])


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
