"""Deterministic synthetic streams shared by the CPU and GPU parity tests and by bench.py."""
import numpy as np


def stream(kind, n, seed):
    r = np.random.default_rng(seed)
    if n == 0:
        return np.zeros(0, np.uint8)
    if kind == "skew8":        # 8 symbols, ~2 bit/sym (BASELINE.md probe distribution)
        p = np.array([.55, .20, .10, .06, .04, .03, .01, .01])
        return r.choice(np.arange(8, dtype=np.uint8) + 33, size=n, p=p).astype(np.uint8)
    if kind == "qual":         # binned Illumina-like, Markov
        syms = np.frombuffer(b"F:,#", dtype=np.uint8)
        stay = r.random(n) < 0.9
        pick = r.choice(4, size=n, p=[.88, .07, .04, .01])
        out = np.empty(n, np.uint8)
        cur = 0
        idx = np.where(~stay)[0]
        state = np.zeros(n, np.int64)
        state[idx] = pick[idx]
        # forward-fill last change
        last = np.maximum.accumulate(np.where(~stay, np.arange(n), 0))
        state = np.where(last > 0, pick[last], pick[0] if not stay[0] else cur)
        return syms[state].astype(np.uint8)
    if kind == "uniform256":
        return r.integers(0, 256, size=n, dtype=np.uint8)
    if kind == "all256":       # every byte value present, skewed
        a = r.integers(0, 256, size=n, dtype=np.uint8)
        if n >= 256:
            a[:256] = np.arange(256, dtype=np.uint8)
        m = r.random(n) < 0.7
        a[m & (np.arange(n) >= 256)] = 65
        return a
    if kind == "const":
        return np.full(n, 71, np.uint8)
    if kind == "two":          # 2 symbols -> PACK 8/byte
        return (r.random(n) < 0.2).astype(np.uint8) * 3 + 48
    if kind == "four":         # <=4 symbols -> PACK 4/byte
        return r.choice(np.frombuffer(b"ACGT", np.uint8), size=n, p=[.4, .3, .2, .1]).astype(np.uint8)
    if kind == "sixteen":      # <=16 symbols -> PACK 2/byte
        return (r.integers(0, 16, size=n) * 3 + 40).astype(np.uint8)
    if kind == "seventeen":    # 17 symbols -> PACK dropped
        return (r.integers(0, 17, size=n) * 3 + 40).astype(np.uint8)
    if kind == "u32le":        # little-endian uint32 counters (STRIPE-friendly)
        v = (np.cumsum(r.integers(0, 40, size=(n + 3) // 4)) + 1000).astype("<u4")
        return v.view(np.uint8)[:n].copy()
    if kind == "runs":         # long runs (arith RLE candidate)
        vals = r.integers(0, 6, size=n // 20 + 1).astype(np.uint8) + 60
        lens = r.integers(1, 60, size=vals.size)
        return np.repeat(vals, lens)[:n].astype(np.uint8) if lens.sum() >= n else np.resize(np.repeat(vals, lens), n).astype(np.uint8)
    if kind == "zeros_hi":     # symbol 0 and 255 present
        return r.choice(np.array([0, 1, 2, 254, 255], np.uint8), size=n, p=[.5, .2, .1, .1, .1]).astype(np.uint8)
    if kind == "text":
        words = [b"@A00123:45:HXXXXXXXX:", b"1:", b"2:", b"1101:", b"2204:", b" 1:N:0:ACGT", b"\n"]
        buf = bytearray()
        while len(buf) < n:
            buf += words[r.integers(0, len(words))] + str(int(r.integers(0, 30000))).encode()
        return np.frombuffer(bytes(buf[:n]), np.uint8).copy()
    raise ValueError(kind)


KINDS = ["skew8", "qual", "uniform256", "all256", "const", "two", "four", "sixteen", "seventeen",
         "u32le", "runs", "zeros_hi", "text"]
EDGE_SIZES = [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 19, 20, 21, 22, 23, 31, 32, 33, 63, 64, 65, 100,
              255, 256, 257, 1000, 4095, 4096, 4097, 10000]


# ---------------------------------------------------------------------------------------------------------
# VBlock-shaped synthetic inputs: what the segmenter hands to the codec path
def fastq_vb(n_reads, read_len, seed, diverse_frac=0.03, lower_frac=0.0, n_frac=0.001):
    """Illumina-like FASTQ VBlock: returns (seq, qual) as uint8 arrays of n_reads*read_len bytes (lines are
    fixed length read_len).  QUAL: binned {F,:,,,#} with Markov run structure (P(stay)=0.97) so that DOMQ
    triggers; a few lines are 'diverse'.  SEQ: uniform ACGT with n_frac 'N'."""
    r = np.random.default_rng(seed)
    n = n_reads * read_len
    seq = np.frombuffer(b"ACGT", np.uint8)[r.integers(0, 4, size=n)].copy()
    if n_frac:
        seq[r.random(n) < n_frac] = ord("N")
    if lower_frac:
        m = r.random(n) < lower_frac
        seq[m] = seq[m] + 32
    syms = np.frombuffer(b"F:,#", np.uint8)
    change = r.random(n) > 0.97
    change[0] = True
    pick = r.choice(4, size=n, p=[.88, .07, .04, .01])
    last = np.maximum.accumulate(np.where(change, np.arange(n), 0))
    qual = syms[pick[last]].copy().reshape(n_reads, read_len)
    div = np.where(r.random(n_reads) < diverse_frac)[0]
    for i in div:
        qual[i] = syms[r.choice(4, size=read_len, p=[.4, .3, .2, .1])]
    return seq, qual.reshape(-1)


def line_table(n_reads, read_len, base=0):
    off = (np.arange(n_reads, dtype=np.uint64) * np.uint64(read_len)) + np.uint64(base)
    ln = np.full(n_reads, read_len, dtype=np.uint32)
    return off, ln


def ragged_quals(seed, n_lines=400):
    """ragged lines incl. empty ones, several doms, long cross-line dom runs, all-dom tails"""
    r = np.random.default_rng(seed)
    lens = r.integers(0, 300, size=n_lines).astype(np.uint32)
    lens[r.random(n_lines) < 0.1] = 0
    parts = []
    for L in lens:
        dom = r.choice(np.frombuffer(b"FI?5", np.uint8), p=[.6, .2, .1, .1])
        p = r.random()
        if p < 0.5:
            line = np.full(L, dom, np.uint8)
            k = int(r.integers(0, 6))
            if L and k:
                line[r.integers(0, L, size=k)] = r.choice(np.frombuffer(b"#,:<", np.uint8), size=k)
        elif p < 0.8:
            line = np.where(r.random(L) < 0.9, dom, r.choice(np.frombuffer(b"#,:<AB", np.uint8), size=L)).astype(np.uint8)
        else:
            line = r.integers(33, 75, size=L).astype(np.uint8)
        parts.append(line)
    txt = np.concatenate(parts) if parts else np.zeros(0, np.uint8)
    off = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    return txt, off, lens


def haplotype_matrix(n_lines, n_samples, seed, ploidy=2, multi=False):
    """VCF-like phased genotype matrix: copy model over haplotypes so PBWT runs are long."""
    r = np.random.default_rng(seed)
    w = n_samples * ploidy
    ht = np.empty((n_lines, w), np.uint8)
    cur = (r.random(w) < 0.1).astype(np.uint8)
    for i in range(n_lines):
        af = r.beta(0.3, 2.0)
        col = (r.random(w) < af).astype(np.uint8)
        # neighbours tend to share alleles: smooth with a random block structure
        blocks = np.repeat(r.random(w // 8 + 1) < af, 8)[:w]
        col = np.where(r.random(w) < 0.8, blocks.astype(np.uint8), col)
        row = col + ord("0")
        if multi and i % 7 == 0:
            row[r.random(w) < 0.02] = ord("2")
        if multi and i % 11 == 0:
            row[r.random(w) < 0.01] = ord(".")
        ht[i] = row
    return ht


def longread_vb(n_reads, mean_len, seed):
    """Nanopore-like reads: lengths mean_len +-20%, quals from an AR(1) process over Phred 1..50."""
    r = np.random.default_rng(seed)
    lens = np.maximum(4, (mean_len * (1 + 0.2 * r.standard_normal(n_reads))).astype(np.int64)).astype(np.uint32)
    n = int(lens.sum())
    seq = np.frombuffer(b"ACGT", np.uint8)[r.integers(0, 4, size=n)].copy()
    e = r.standard_normal(n) * 4.0
    q = np.empty(n, np.float64)
    acc = 20.0
    # AR(1) via scipy-free recursion in blocks (vectorised with lfilter-like cumulative trick is overkill here)
    phi = 0.9
    x = np.zeros(n)
    x[0] = e[0]
    for i in range(1, min(n, 200000)):
        x[i] = phi * x[i - 1] + e[i]
    if n > 200000:
        x[200000:] = np.resize(x[:200000], n - 200000)
    q = np.clip(np.rint(20 + x), 1, 50).astype(np.uint8) + 33
    return seq, q, lens
