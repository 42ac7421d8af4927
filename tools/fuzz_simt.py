"""Differential fuzzing of the CUDA codec kernels WITHOUT a GPU: random streams through the product's kernels on the SIMT
emulator (tests/host/simt) against the reference's own compiled htscodecs (oracle/_ref/libhts_ref.so) — compressed bytes
must be identical, the decode bit-exact — and, with --corrupt, damaged streams through both decoders: same verdict, and the
same bytes where the reference accepts the stream.

    python tools/fuzz_simt.py --seconds 120 [--seed 1] [--corrupt] [--max-n 60000]

Test tooling; tests/test_simt_fuzz.py runs a short seeded pass of it in the CPU suite."""
import argparse, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc                                                                  # noqa: E402
from simt_lib import simt_engine_class                                      # noqa: E402

CODECS = ["RANB", "RANW", "RANb", "RANw", "ARTB", "ARTW", "ARTb", "ARTw"]
SIZES = [0, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 19, 20, 21, 22, 23, 31, 32, 33, 63, 64, 65, 100, 255, 256, 257, 1000, 4095, 4096, 4097]


def random_stream(r, max_n):
    n = int(r.choice(SIZES)) if r.random() < 0.25 else int(r.integers(1, max_n) if r.random() < 0.5 else r.integers(1, 3000))
    if n == 0:
        return np.zeros(0, np.uint8), "empty"
    k = int(r.choice([1, 2, 3, 4, 5, 8, 15, 16, 17, 40, 100, 255, 256])) if r.random() < 0.6 else int(r.integers(1, 257))
    syms = r.permutation(256)[:k].astype(np.uint8)
    if r.random() < 0.3:
        syms = np.sort(syms)
    if r.random() < 0.2:
        syms[0] = 0
    if r.random() < 0.2:
        syms[-1] = 255
    conc = float(r.choice([0.02, 0.1, 0.5, 2.0, 50.0]))
    p = r.dirichlet(np.full(k, conc)) + 1e-9; p /= p.sum()
    shape = r.choice(["iid", "markov", "runs", "stripe", "hot", "ramp"])
    if shape == "iid":
        x = syms[r.choice(k, size=n, p=p)]
    elif shape == "markov":                                                 # order-1 structure: stay with probability q
        q = float(r.choice([0.5, 0.9, 0.99, 0.999]))
        change = r.random(n) > q; change[0] = True
        pick = r.choice(k, size=n, p=p)
        last = np.maximum.accumulate(np.where(change, np.arange(n), 0))
        x = syms[pick[last]]
    elif shape == "runs":
        m = n // 8 + 1
        vals = syms[r.choice(k, size=m, p=p)]; lens = r.geometric(float(r.choice([0.02, 0.1, 0.5])), size=m)
        x = np.resize(np.repeat(vals, lens), n)
    elif shape == "stripe":                                                 # 4-byte records with different statistics per byte plane
        x = np.empty(n, np.uint8)
        for j in range(4):
            pj = r.dirichlet(np.full(k, conc)) + 1e-9; pj /= pj.sum()
            x[j::4] = syms[r.choice(k, size=x[j::4].size, p=pj)]
    elif shape == "hot":                                                    # one symbol nearly always: the rANS hot transition, the arithmetic run step
        x = np.full(n, syms[0], np.uint8)
        m = r.random(n) < float(r.choice([0.0, 0.0005, 0.01, 0.05]))
        x[m] = syms[r.choice(k, size=int(m.sum()), p=p)]
    else:
        x = syms[(np.arange(n) * int(r.integers(1, 7)) // int(r.integers(1, 50))) % k]
    return np.ascontiguousarray(x, np.uint8), f"{shape} k={k} conc={conc}"


def kind_of(c):
    return "rans" if c.startswith("RAN") else "arith"


def corrupt(r, comp):
    c = comp.copy()
    how = r.choice(["flip", "flip_head", "truncate", "extend", "zero_tail"])
    if how == "flip" and c.size:
        for _ in range(int(r.integers(1, 4))):
            c[int(r.integers(0, c.size))] ^= np.uint8(1 << int(r.integers(0, 8)))
    elif how == "flip_head" and c.size:
        c[int(r.integers(0, min(c.size, 12)))] = np.uint8(r.integers(0, 256))
    elif how == "truncate" and c.size > 1:
        c = c[:int(r.integers(1, c.size))].copy()
    elif how == "extend":
        c = np.concatenate([c, r.integers(0, 256, size=int(r.integers(1, 9)), dtype=np.uint8)])
    elif c.size > 4:
        c[-int(r.integers(1, min(c.size, 40))):] = 0
    return c, how


def ref_uncompress_port(kind, comp, n):
    try:
        return orc.uncompress("port", kind, comp, n)
    except AssertionError:
        return None


def _varint_len(b, i):
    j = i
    while j < b.size and j - i < 5 and b[j] & 0x80:
        j += 1
    return j + 1 - i


def o1_table_has_empty_context(b):
    """order-1 rANS container (not STRIPE/PACK, table stored uncompressed): is there a context of the alphabet without
    frequencies, or is the start context 0 missing from the alphabet?  The reference skips its table row (rANS_static4x16pr.c:994-997); a damaged stream that enters it reads
    whatever earlier calls left there."""
    try:
        b = [int(v) for v in b]
        flags, p = b[0], 1
        if flags & 0x08:                                                    # STRIPE: flags, size, N, N lengths, N containers (:1366-1439)
            while b[p] & 0x80:
                p += 1
            p += 1
            N = b[p]; p += 1
            clen = []
            for _ in range(N):
                v = 0
                while True:
                    c = b[p]; p += 1; v = (v << 7) | (c & 0x7f)
                    if not c & 0x80:
                        break
                clen.append(v)
            for cl in clen:
                if o1_table_has_empty_context(b[p:]):
                    return True
                p += cl
            return False
        if not flags & 1:
            return False
        if not flags & 0x10:
            while b[p] & 0x80:
                p += 1
            p += 1
        if flags & 0x80:                                                    # PACK meta: symbol count, map, packed length (pack.c:168-201)
            ns = b[p] or 256
            p += 1 + (ns if ns <= 16 else 0)
            while b[p] & 0x80:
                p += 1
            p += 1
        if b[p] & 1:
            return False
        p += 1
        F0, run, j = [0] * 256, 0, b[p]; p += 1
        while True:
            F0[j] = 1
            if not run and j + 1 == b[p]:
                j = b[p]; run = b[p + 1]; p += 2
            elif run:
                run -= 1; j += 1
            else:
                j = b[p]; p += 1
            if j == 0 or j > 255:
                break
        syms = [i for i in range(256) if F0[i]]
        if not F0[0]:                                                       # decoding starts in context 0 (:1029): not in the alphabet, no row either
            return True
        for _ in syms:
            T, zrun = 0, 0
            for _j in syms:
                if zrun:
                    zrun -= 1; continue
                f = 0
                while True:
                    c = b[p]; p += 1; f = (f << 7) | (c & 0x7f)
                    if not c & 0x80:
                        break
                if f == 0:
                    zrun = b[p]; p += 1
                T += f
            if T == 0:
                return True
    except IndexError:
        pass
    return False


def known_stricter(codec, b):
    """damaged streams that the reference decodes (to garbage) and the kernels refuse — by design"""
    if codec.startswith("RAN") and b.size > 1 and (int(b[0]) & 0x48) == 0x40:
        return "rANS container with the RLE flag: never written by genozip (codec_htscodecs.c:17-20), not implemented"
    if codec.startswith("RAN") and b.size > 4:
        flags = int(b[0])
        i = 1
        if flags & 0x08:
            return None
        if not flags & 0x10:
            i += _varint_len(b, i)
        if flags & 0x80:
            return None
        if (flags & 1) and i < b.size and (int(b[i]) >> 4) not in (10, 12):
            return "order-1 rANS table with a frequency shift the encoder never writes (not 10 or 12): the reference's tables overlap"
    return None


HANG = "hang"


def ref_uncompress(kind, comp, n, timeout=10.0, scramble=False):
    """the reference's decoder on a damaged stream, in a child process (it can crash or spin on such input)
    -> bytes, None where it rejects the stream, HANG where it does not come back or dies"""
    import select, signal, warnings
    rd, wr = os.pipe()
    with warnings.catch_warnings():                                         # (the child only calls into the reference's C code and exits;
        warnings.simplefilter("ignore", DeprecationWarning)                 #  a child that does not answer is killed and counted)
        pid = os.fork()
    if pid == 0:
        try:
            os.close(rd)
            try:
                if scramble:                                                # leave other leftovers in the reference's thread-local decode table
                    rr = np.random.default_rng(12345)
                    junk = rr.integers(0, 256, 50000, dtype=np.uint8)
                    orc.uncompress("ref", "rans", orc.compress("ref", "rans", junk, 1), junk.size)
                out = orc.uncompress("ref", kind, comp, n)
                os.write(wr, b"A" + out.tobytes())
            except AssertionError:
                os.write(wr, b"R")
        finally:
            os._exit(0)
    os.close(wr)
    buf = b""
    t_end = time.time() + timeout
    while True:
        left = t_end - time.time()
        ready = select.select([rd], [], [], max(left, 0))[0] if left > 0 else []
        if not ready:
            os.kill(pid, signal.SIGKILL)
            break
        chunk = os.read(rd, 1 << 20)
        if not chunk:
            break
        buf += chunk
    os.close(rd)
    os.waitpid(pid, 0)
    if not buf:
        return HANG
    return None if buf[:1] == b"R" else np.frombuffer(buf[1:], np.uint8)


def run(seconds, seed, do_corrupt, max_n, verbose=False, stricter=None, max_streams=None):
    from genozip_b200 import GzbError
    eng = simt_engine_class()(0)
    r = np.random.default_rng(seed)
    t0, n_cases, n_bytes, n_corrupt, n_rejected = time.time(), 0, 0, 0, 0
    while time.time() - t0 < seconds and (max_streams is None or n_cases < max_streams):
        batch = []
        for _ in range(int(r.integers(1, 24))):
            x, what = random_stream(r, max_n)
            batch.append((str(r.choice(CODECS)), x, what))
        got = eng.compress([(c, x) for c, x, _ in batch])
        want = [orc.compress("ref", kind_of(c), x, orc.ORDER[c]) for c, x, _ in batch]
        for (c, x, what), g, w in zip(batch, got, want):
            if g.size != w.size or not np.array_equal(g, w):
                np.save("/tmp/fuzz_fail.npy", x)
                raise AssertionError(f"COMPRESS MISMATCH codec {c} n={x.size} [{what}] seed={seed}: got {g.size} bytes, reference {w.size}; input saved to /tmp/fuzz_fail.npy")
        dec = [(c, w, x.size) for (c, x, _), w in zip(batch, want) if x.size]
        for (c, w, n), o, (_, x, what) in zip(dec, eng.uncompress(dec), [b for b in batch if b[1].size]):
            if not np.array_equal(o, x):
                np.save("/tmp/fuzz_fail.npy", x)
                raise AssertionError(f"DECODE MISMATCH codec {c} n={n} [{what}] seed={seed}; input saved to /tmp/fuzz_fail.npy")
        n_cases += len(batch); n_bytes += sum(x.size for _, x, _ in batch)
        if do_corrupt:
            for (c, x, what), w in zip(batch, want):
                if not x.size or r.random() < 0.5:
                    continue
                bad, how = corrupt(r, w)
                if os.environ.get("FUZZ_TRACE"):                            # a hang leaves its stream behind
                    np.save("/tmp/fuzz_last.npy", bad); open("/tmp/fuzz_last.txt", "w").write(f"{c} {x.size} {how} {what}\n")
                ref_out = ref_uncompress(kind_of(c), bad, x.size)
                if ref_out is HANG:                                         # nothing to compare with; the kernels must still come back
                    hangs = stricter.setdefault("the reference itself spins or crashes on the stream", 0) if stricter is not None else 0
                    if stricter is not None:
                        stricter["the reference itself spins or crashes on the stream"] = hangs + 1
                    ref_out = None
                    try:
                        eng.uncompress([(c, bad, x.size)])
                    except GzbError:
                        pass
                    continue
                try:
                    out = eng.uncompress([(c, bad, x.size)])[0]
                except GzbError:
                    out = None
                n_corrupt += 1; n_rejected += ref_out is None
                if ref_out is not None and out is None and stricter is not None:
                    # the kernels refuse a damaged stream the reference decodes to SOMETHING: tolerated only for the documented
                    # classes below, and only if the CPU restatement draws the same line
                    why = known_stricter(c, bad)
                    port_rejects = ref_uncompress_port(kind_of(c), bad, x.size) is None
                    if why and port_rejects:
                        stricter[why] = stricter.get(why, 0) + 1
                        continue
                if ref_out is not None and out is not None and not np.array_equal(out, ref_out) and stricter is not None:
                    again = ref_uncompress(kind_of(c), bad, x.size, scramble=True)
                    if again is HANG or again is None or not np.array_equal(again, ref_out) or (c.startswith("RAN") and o1_table_has_empty_context(bad)):
                        k = "the reference's own output depends on what earlier calls left in its tables (a context without frequencies is entered)"
                        stricter[k] = stricter.get(k, 0) + 1
                        continue
                if (ref_out is None) != (out is None) or (out is not None and not np.array_equal(out, ref_out)):
                    np.save("/tmp/fuzz_fail.npy", bad)
                    raise AssertionError(f"CORRUPT-STREAM MISMATCH codec {c} n={x.size} [{what}] damage={how} seed={seed}: reference "
                                         f"{'rejects' if ref_out is None else 'accepts'}, kernels {'reject' if out is None else 'accept'}"
                                         f"{'' if out is None or ref_out is None else ' with different bytes'}; stream saved to /tmp/fuzz_fail.npy")
        if verbose:
            print(f"{time.time() - t0:6.1f}s  {n_cases} streams, {n_bytes / 1e6:.1f} MB, {n_corrupt} damaged ({n_rejected} rejected by the reference)", flush=True)
    eng.close()
    return n_cases, n_bytes, n_corrupt, n_rejected


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60); ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--corrupt", action="store_true"); ap.add_argument("--max-n", type=int, default=60000)
    a = ap.parse_args()
    assert orc.have_ref(), "needs oracle/_ref/libhts_ref.so (make -C oracle)"
    stricter = {}
    print("ok: %d streams, %d bytes, %d damaged streams (%d rejected by the reference)" % run(a.seconds, a.seed, a.corrupt, a.max_n, verbose=True, stricter=stricter))
    for k, v in stricter.items():
        print(f"   refused although the reference decodes it ({v}x): {k}")
