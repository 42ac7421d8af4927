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
