"""Differential fuzzing of the genozip-codec kernels (ACGT/XCGT, DOMQ, PBWT, LONGR) WITHOUT a GPU: random VBlocks through the
product's kernels on the SIMT emulator (tests/host/simt) against the reference's own compiled codec_acgt.c / codec_domq.c /
codec_pbwt.c / codec_longr.c (oracle/_ref/libgz_ref.so) — every stream byte-identical, and back through the kernels' decoders
and the reference's.

    python tools/fuzz_simt_gz.py --seconds 120 [--seed 1]

Test tooling; tests/test_simt_fuzz.py runs a short seeded pass of it in the CPU suite."""
import argparse, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import orc                                                                  # noqa: E402
from simt_lib import simt_engine_class                                      # noqa: E402


def rand_lines(r, max_lines, max_len):
    n_lines = int(r.integers(1, max_lines))
    mode = r.choice(["fixed", "ragged", "ragged0"])
    if mode == "fixed":
        lens = np.full(n_lines, int(r.integers(1, max_len)), np.uint32)
    else:
        lens = r.integers(0 if mode == "ragged0" else 1, max_len, size=n_lines).astype(np.uint32)
        if mode == "ragged0":
            lens[r.random(n_lines) < 0.2] = 0
    if not lens.sum():
        lens[0] = 1
    off = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64)
    return off, lens


def rand_quals(r, off, lens):
    """quality text: a few dominant values with run structure, per-line dom changes, diverse lines, rare values, ties"""
    n = int(lens.sum())
    k = int(r.choice([1, 2, 3, 4, 8, 20, 41, 94]))
    alphabet = (33 + r.permutation(94)[:k]).astype(np.uint8)
    stay = float(r.choice([0.0, 0.5, 0.9, 0.97, 0.999]))
    q = np.empty(n, np.uint8)
    pos = 0
    for L in lens:
        L = int(L)
        if not L:
            continue
        style = r.random()
        p = r.dirichlet(np.full(k, float(r.choice([0.05, 0.3, 2.0])))) + 1e-9; p /= p.sum()
        if style < 0.15:                                                    # diverse line
            line = alphabet[r.integers(0, k, L)]
        elif style < 0.25:                                                  # exact tie between two values
            line = np.resize(alphabet[[0, min(1, k - 1)]], L)
        else:
            change = r.random(L) > stay; change[0] = True
            pick = r.choice(k, size=L, p=p)
            last = np.maximum.accumulate(np.where(change, np.arange(L), 0))
            line = alphabet[pick[last]]
        q[pos:pos + L] = line; pos += L
    return q


def fuzz_domq(r, eng):
    off, lens = rand_lines(r, int(r.choice([3, 40, 400])), int(r.choice([2, 60, 300, 2600, 6000])))   # (above 4096: the histogram pass gives the line to a whole warp)
    q = rand_quals(r, off, lens)
    g = eng.domq_encode([(q, off, lens)])[0]
    ref = orc.ref_domq_encode(q, off, lens)
    for k in ("qual", "runs", "mplx", "divr", "denorm"):
        assert g[k].size == ref[k].size and np.array_equal(g[k], ref[k]), f"DOMQ {k}: kernels != reference codec_domq.c ({g[k].size} vs {ref[k].size} bytes)"
    assert g["num_norm_qs"] == ref["num_norm_qs"]
    want = np.concatenate([q[int(o):int(o) + int(l)] for o, l in zip(off, lens)])
    assert np.array_equal(eng.domq_decode([g], [lens])[0], want), "DOMQ: kernels' reconstruct"
    assert np.array_equal(orc.ref_domq_decode(g, lens), want), "DOMQ: reference's reconstruct on the kernels' streams"
    return q.size


def fuzz_acgt(r, eng):
    n = int(r.choice([1, 5, 31, 32, 33, 64, 65, 1000, 4097])) if r.random() < 0.4 else int(r.integers(1, 40000))
    style = r.random()
    if style < 0.4:
        s = np.frombuffer(b"ACGT", np.uint8)[r.integers(0, 4, n)].copy()
    elif style < 0.8:
        s = np.frombuffer(b"ACGT", np.uint8)[r.integers(0, 4, n)].copy()
        m = r.random(n) < float(r.choice([0.001, 0.05, 0.5]))
        s[m] = np.frombuffer(b"NacgtnRYSWKMBDHVU*-.", np.uint8)[r.integers(0, 20, int(m.sum()))]
    else:
        s = r.integers(2, 256, n).astype(np.uint8)                          # (bytes 0 and 1 do not survive the format: DESIGN.md §5)
    p, x, allz = eng.acgt_pack(s)
    pr, xr, zr = orc.ref_acgt_pack(s)
    assert np.array_equal(p, pr) and np.array_equal(x, xr) and allz == zr, f"ACGT n={n}: kernels != reference codec_acgt.c"
    assert np.array_equal(eng.acgt_unpack(p, None if allz else x, n), s), "ACGT: kernels' unpack"
    assert np.array_equal(orc.ref_acgt_unpack(p, None if allz else x, n), s), "ACGT: reference's unpack"
    got = eng.acgt_pack_batch([s, s[: n // 2].copy()])
    assert np.array_equal(got[0][0], pr) and np.array_equal(got[0][1], xr)
    return n


def fuzz_pbwt(r, eng):
    n_lines, w = int(r.integers(1, 120)), int(r.choice([1, 2, 3, 8, 33, 100, 700]))
    alleles = np.frombuffer(b"01" if r.random() < 0.5 else b"0011122.*%-&345", np.uint8)
    if r.random() < 0.5:                                                    # haplotype-block structure: few founders, rare mutations
        founders = alleles[r.integers(0, alleles.size, (int(r.integers(1, 6)), n_lines))]
        ht = founders[r.integers(0, founders.shape[0], w)].T.copy()
        m = r.random(ht.shape) < 0.01
        ht[m] = alleles[r.integers(0, alleles.size, int(m.sum()))]
    else:
        ht = alleles[r.integers(0, alleles.size, (n_lines, w))]
    ht = np.ascontiguousarray(ht, np.uint8)
    runs, fgrc = eng.pbwt_encode(ht)
    rr, rf = orc.ref_pbwt_encode(ht)
    assert np.array_equal(runs, rr) and np.array_equal(fgrc, rf), f"PBWT {ht.shape}: kernels != reference codec_pbwt.c"
    assert np.array_equal(eng.pbwt_decode(rr, rf, n_lines, ht.size).reshape(ht.shape), ht), "PBWT: kernels' decode"
    assert np.array_equal(orc.ref_pbwt_decode(runs, fgrc, n_lines, ht.size).reshape(ht.shape), ht), "PBWT: reference's decode"
    return ht.size


def fuzz_longr(r, eng):
    n_reads = int(r.integers(1, 40))
    lens = r.integers(1, int(r.choice([4, 100, 4000])), n_reads).astype(np.uint32)
    n = int(lens.sum())
    seq = np.frombuffer(b"ACGTN", np.uint8)[r.choice(5, n, p=[.25, .25, .25, .24, .01])].copy()
    k = int(r.choice([2, 8, 40, 93]))
    base = r.integers(0, k, n)
    qual = (33 + np.clip(base + (r.integers(-3, 4, n) * (r.random(n) < 0.3)), 0, 92)).astype(np.uint8)
    txt = np.concatenate([seq, qual])
    seq_off = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64); qual_off = (seq_off + np.uint64(n)).astype(np.uint64)
    is_rev = (r.random(n_reads) < 0.5).astype(np.uint8) if r.random() < 0.6 else None
    if is_rev is not None:
        is_rev[lens < 3] = 0                 # codec_longr_alg.c:156 tests an UNSIGNED `seq_len-1-i >= 0`: a reversed read shorter than
                                             # B_AHEAD_OF_Q makes the reference read 4 GB past its buffer (kernels and restatement pad with 'T')
    v2b_ref, vals_ref, lb_ref = orc.ref_longr_encode(txt, seq_off, qual_off, lens, is_rev)
    v2b = orc.longr_bins(qual)
    assert np.array_equal(v2b, v2b_ref)
    vb = (txt, seq_off, qual_off, lens, is_rev, v2b)
    vals, lb = eng.longr_encode([vb])[0]
    assert np.array_equal(vals, vals_ref) and np.array_equal(lb, lb_ref), "LONGR: kernels != reference codec_longr.c"
    assert np.array_equal(eng.longr_decode([vb], [vals], [lb])[0], qual), "LONGR: kernels' decode"
    assert np.array_equal(orc.ref_longr_decode(txt, seq_off, lens, is_rev, v2b, vals, lb), qual), "LONGR: reference's decode"
    return n


# ---- round 2: the steps either side of the codecs and the other quality codecs, against the reference's compiled b250.c, codec_oq.c,
# codec_smux.c, codec_pacb.c, codec_homp.c, codec_t0.c
def rand_reads(r, max_lines=200, max_len=400):
    """a text holding, per read, SEQ (with homopolymer runs), QUAL and a second quality-like string, at random places"""
    n_lines = int(r.integers(1, max_lines))
    parts, pos, so, qo, oo, lens = [np.frombuffer(b"@", np.uint8)], 1, [], [], [], []
    for _ in range(n_lines):
        L = int(r.integers(1, max_len))
        seq = np.repeat(r.choice(np.frombuffer(b"ACGTN", np.uint8), L, p=[.24, .24, .24, .24, .04]), r.choice([1, 1, 1, 2, 3, 7], L))[:L].astype(np.uint8)
        k = int(r.choice([2, 4, 40]))
        qual = (33 + r.integers(0, k, L) * (r.random(L) < float(r.choice([0.1, 0.5, 1.0])))).astype(np.uint8)
        if r.random() < 0.5:                                                  # Ultima-like: mirror the qualities inside every run
            i = 0
            while i < L:
                h = 1
                while i + h < L and seq[i + h] == seq[i]: h += 1
                half = qual[i:i + (h + 1) // 2].copy(); qual[i:i + h] = np.concatenate([half, half[:h // 2][::-1]]); i += h
        oq = (33 + np.clip(qual.astype(np.int32) - 33 + r.integers(-1, 2, L), 0, 93)).astype(np.uint8)
        for arr, lst in ((seq, so), (qual, qo), (oq, oo)):
            lst.append(pos); parts.append(arr); pos += L
            fill = r.integers(65, 91, int(r.integers(0, 5))).astype(np.uint8); parts.append(fill); pos += fill.size
        lens.append(L)
    return np.concatenate(parts), np.array(so, np.uint64), np.array(qo, np.uint64), np.array(oo, np.uint64), np.array(lens, np.uint32)


def _off(lens):
    return np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint64)


def fuzz_oq(r, eng):
    txt, so, qo, oo, lens = rand_reads(r)
    g = eng.oq_mux([(txt, qo, lens, oo, None)])[0]
    ref = orc.oq_mux(txt, qo, lens, oo, None, "ref")
    assert all(np.array_equal(a, b) for a, b in zip(g, ref)), "OQ: kernels != reference codec_oq.c"
    cnt = np.where(g[2] != 0, 0, g[1]).astype(np.uint32); at = np.concatenate([[0], np.cumsum(g[1])]).astype(np.int64)
    keep = [g[0][at[q]:at[q + 1]] for q in range(94) if cnt[q]]
    ch = np.concatenate(keep) if keep else np.zeros(0, np.uint8)
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(oo, lens)])
    assert np.array_equal(eng.oq_demux([(txt, qo, lens, _off(lens), int(lens.sum()), 33, ch, cnt, g[2])])[0], want), "OQ: kernels' demux"
    assert np.array_equal(orc.oq_demux(txt, qo, lens, _off(lens), int(lens.sum()), 33, ch, cnt, g[2], "ref"), want), "OQ: reference's reconstruct"
    return int(lens.sum())


def fuzz_smux(r, eng):
    txt, so, qo, oo, lens = rand_reads(r)
    rev = (r.random(lens.size) < 0.4).astype(np.uint8) if r.random() < 0.6 else None
    g = eng.smux_mux([(txt, qo, lens, so, lens, rev)])[0]
    ref = orc.smux_mux(txt, qo, lens, so, lens, rev, "ref")
    assert np.array_equal(g[0], ref[0]) and np.array_equal(g[1], ref[1]) and g[2] == ref[2], "SMUX: kernels != reference codec_smux.c"
    cnt = g[1].copy(); ch = g[0]
    if g[2]:
        ch = ch[:int(cnt[:4].sum())]; cnt[4] = 0
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(qo, lens)])
    assert np.array_equal(eng.smux_demux([(txt, so, lens, rev, _off(lens), int(lens.sum()), ch, cnt, g[2])])[0], want), "SMUX: kernels' demux"
    assert np.array_equal(orc.smux_demux(txt, so, lens, rev, _off(lens), int(lens.sum()), ch, cnt, g[2], "ref"), want), "SMUX: reference's reconstruct"
    return int(lens.sum())


def fuzz_pacb(r, eng):
    txt, so, qo, oo, lens = rand_reads(r)
    max_np = int(r.choice([1, 12]))
    np0 = r.integers(0, max_np, lens.size).astype(np.uint8) if max_np > 1 else None
    g = eng.pacb_mux([(txt, qo, lens, so, np0, max_np)])[0]
    ref = orc.pacb_mux(txt, qo, lens, so, np0, max_np, "ref")
    assert np.array_equal(g[0], ref[0]) and np.array_equal(g[1], ref[1]), "PACB: kernels != reference codec_pacb.c"
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(qo, lens)])
    assert np.array_equal(eng.pacb_demux([(txt, so, lens, np0, max_np, _off(lens), int(lens.sum()), g[0], g[1])])[0], want), "PACB: kernels' demux"
    assert np.array_equal(orc.pacb_demux(txt, so, lens, np0, max_np, _off(lens), int(lens.sum()), g[0], g[1], "ref"), want), "PACB: reference's reconstruct"
    return int(lens.sum())


def fuzz_homp(r, eng):
    txt, so, qo, oo, lens = rand_reads(r)
    mode = int(r.integers(0, 2))
    g = eng.hp_condense(mode, [(txt, qo, lens, so)])[0]
    ref = orc.hp_condense(mode, txt, qo, lens, so, "ref")
    assert np.array_equal(g[0], ref[0]) and np.array_equal(g[1], ref[1]), f"HOMP/T0 mode {mode}: kernels != reference"
    want = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(qo, lens)])
    assert np.array_equal(eng.hp_expand(mode, [(g[0], txt, so, lens)])[0][0], want), "HOMP/T0: kernels' expand"
    assert np.array_equal(orc.hp_expand(mode, g[0], txt, so, lens, "ref")[0], want), "HOMP/T0: reference's reconstruct"
    return int(lens.sum())


def fuzz_b250(r, eng):
    n_words = int(r.choice([0, 1, 2, 30, 31, 33, 500, 5000])); ol = int(r.choice([0, 5, 200, 5000, 3000000])); n_new = int(r.choice([0, 1, 40, 2000]))
    ni2wi = (ol + r.permutation(n_new)).astype(np.int32)
    out, wi = [], 0
    for _ in range(n_words):
        x = r.random()
        if x < 0.3 and ol:   wi = (wi + 1) % ol; v, f4 = wi, False
        elif x < 0.4:        v, f4 = (-3 if r.random() < 0.5 else -4), False
        elif x < 0.6 and n_new: v, f4 = ol + int(r.integers(0, n_new)), True
        elif ol:             wi = int(r.integers(0, ol)); v, f4 = wi, False
        else:                v, f4 = -3, False
        if f4 or v > 2113660: enc, n = (7 << 29) | v, 4
        elif v == -3:  enc, n = 0xBFFE, 2
        elif v == -4:  enc, n = 0xBFFF, 2
        elif v <= 126: enc, n = v, 1
        elif v <= 16508: enc, n = (2 << 14) | (v - 127), 2
        else:          enc, n = (6 << 21) | (v - 16509), 3
        out += [(enc >> (8 * k)) & 0xff for k in range(n)]
    b = np.array(out, np.uint8)
    up = ol + n_new > 1024
    g = eng.b250_generate([(b, ni2wi, ol, up)])[0]
    ref = orc.b250_generate(b, ni2wi, ol, up, "ref")
    assert g[1] == n_words and np.array_equal(g[0], ref[0]), "b250: kernels != reference b250.c"
    return b.size


FUZZERS = {"domq": fuzz_domq, "acgt": fuzz_acgt, "pbwt": fuzz_pbwt, "longr": fuzz_longr,
           "oq": fuzz_oq, "smux": fuzz_smux, "pacb": fuzz_pacb, "homp": fuzz_homp, "b250": fuzz_b250}


def run(seconds, seed, which=None, verbose=False, max_cases=None):
    eng = simt_engine_class()(0)
    r = np.random.default_rng(seed)
    names = list(which or FUZZERS)
    t0, count = time.time(), {k: 0 for k in names}
    i = 0
    while time.time() - t0 < seconds and (max_cases is None or i < max_cases):
        name = names[i % len(names)]; i += 1
        state = r.bit_generator.state
        try:
            FUZZERS[name](r, eng)
        except AssertionError as e:
            raise AssertionError(f"{name} (seed {seed}, case {i}): {e}") from None
        count[name] += 1
        if verbose and i % 20 == 0:
            print(f"{time.time() - t0:6.1f}s  {count}", flush=True)
    eng.close()
    return count


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=60); ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--only", default=None)
    a = ap.parse_args()
    assert orc.have_gz_ref(), "needs oracle/_ref/libgz_ref.so (make -C oracle)"
    print("ok:", run(a.seconds, a.seed, a.only.split(",") if a.only else None, verbose=True))
