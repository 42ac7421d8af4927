"""The drop-in boundary, driven the way genozip drives it: tests/host/plugin_harness.c plays the genozip side (its own VBlock /
Context / Buffer, the two accessor tables, a codec table whose rows point at the library's plug-in entry points, comp_compress
with its soft-fail retry, per-line reconstruct calls) and every gzb_codec_* function runs through the C-ABI.  What the codecs
produced is compared with the reference's compiled objects (oracle/_ref) or the restatement.

The same harness is linked against libgzb200.so on the GPU box (-m gpu) and against the SIMT-emulator build of the same sources
in the CPU suite, where it checks the host-side logic of the plug-in layer (the kernels themselves are the -m gpu tests' job)."""
import ctypes as C
import os, subprocess, sys
import numpy as np, pytest
import orc
from datagen import stream, fastq_vb, line_table, haplotype_matrix, longread_vb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "tests", "host", "_build")
CODEC = {"RANB": 6, "RANW": 7, "RANb": 8, "RANw": 9, "ARTB": 16, "ARTW": 17, "ARTb": 18, "ARTw": 19}
_H = {}


def harness(kind):
    """the harness linked against the CUDA library ("gpu") or the emulator build ("simt")"""
    if kind in _H:
        return _H[kind]
    os.makedirs(BUILD, exist_ok=True)
    if kind == "simt":
        sys.path.insert(0, os.path.join(ROOT, "tests", "host", "simt"))
        import build as simt_build
        lib = simt_build.build()
    else:
        lib = os.path.join(ROOT, "genozip_b200", "libgzb200.so")
    out = os.path.join(BUILD, f"libplugin_harness_{kind}.so")
    src = os.path.join(ROOT, "tests", "host", "plugin_harness.c")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(lib)):
        subprocess.run(["gcc", "-O1", "-g", "-shared", "-fPIC", "-Wall", "-Wno-unused-function", "-I", os.path.join(ROOT, "include"), src, "-o", out,
                        "-L", os.path.dirname(lib), f"-l:{os.path.basename(lib)}", f"-Wl,-rpath,{os.path.dirname(lib)}", "-lpthread"], check=True)
    H = C.CDLL(out)
    H.harness_last_abort.restype = C.c_char_p
    _H[kind] = H
    return H


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _ok(H, rc):
    assert rc == 0, f"harness rc={rc}: {H.harness_last_abort().decode()}"


def check_simple(H, codec, data, lines=None, tight=0):
    data = np.ascontiguousarray(data, np.uint8)
    comp = np.zeros(data.size * 2 + 70000, np.uint8); back = np.zeros(data.size + 16, np.uint8)
    cl, sf = C.c_uint32(), C.c_int()
    ll = None if lines is None else np.ascontiguousarray(lines, np.uint32)
    _ok(H, H.harness_simple(CODEC[codec], _p(data), data.size, None if ll is None else _p(ll), 0 if ll is None else ll.size, tight, _p(comp), C.byref(cl), _p(back), C.byref(sf)))
    kind = "rans" if codec.startswith("RAN") else "arith"
    want = orc.compress("ref" if orc.have_ref() else "port", kind, data, orc.ORDER[codec])
    assert np.array_equal(comp[:cl.value], want), f"{codec}: section differs from the reference's bytes"
    assert np.array_equal(back[:data.size], data)
    assert sf.value == (1 if tight else 0)


def check_acgt(H, seq, lines=None, sub="RANB"):
    seq = np.ascontiguousarray(seq, np.uint8); n = seq.size
    packed = np.zeros(n // 4 + 64, np.uint8); x = np.zeros(n + 16, np.uint8); xc = np.zeros(2 * n + 70000, np.uint8); back = np.zeros(n + 16, np.uint8)
    pl, nox, xcl, sf = C.c_uint32(), C.c_int(), C.c_uint32(), C.c_int()
    ll = None if lines is None else np.ascontiguousarray(lines, np.uint32)
    _ok(H, H.harness_acgt(_p(seq), n, 0 if ll is None else 1, None if ll is None else _p(ll), 0 if ll is None else ll.size, CODEC[sub],
                          _p(packed), C.byref(pl), _p(x), C.byref(nox), _p(xc), C.byref(xcl), _p(back), C.byref(sf)))
    wp, wx, wno = orc.ref_acgt_pack(seq) if orc.have_gz_ref() else orc.acgt_pack(seq)
    assert np.array_equal(packed[:pl.value], wp), "2-bit words differ"
    assert bool(nox.value) == bool(wno)
    if not wno:
        assert np.array_equal(x[:n], wx), "exception stream differs"
        if n >= 50:
            kind = "rans" if sub.startswith("RAN") else "arith"
            assert np.array_equal(xc[:xcl.value], orc.compress("ref" if orc.have_ref() else "port", kind, wx, orc.ORDER[sub])), "NONREF_X section differs"
    assert np.array_equal(back[:n], seq)


def check_domq(H, qual, off, lens, sub="ARTb"):
    qual = np.ascontiguousarray(qual, np.uint8); off = np.ascontiguousarray(off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    tot = int(lens.sum())
    bufs = [np.zeros(2 * tot + 64, np.uint8), np.zeros(tot + 64, np.uint8), np.zeros(lens.size + 64, np.uint8), np.zeros(tot + 64, np.uint8)]
    ls = [C.c_uint32() for _ in range(4)]
    den = np.zeros(95 * 95, np.uint8); dl = C.c_uint32(); prm = C.c_uint8()
    comp = np.zeros(2 * tot + 70000, np.uint8); cl = C.c_uint32(); back = np.zeros(tot + 64, np.uint8); sf = C.c_int(); cnt = np.zeros(4, np.uint64)
    args = []
    for b, l in zip(bufs, ls):
        args += [_p(b), C.byref(l)]
    rc = H.harness_domq(_p(qual), _p(off), _p(lens), lens.size, CODEC[sub], 1, *args, _p(den), C.byref(dl), C.byref(prm), _p(comp), C.byref(cl), _p(back), C.byref(sf), _p(cnt))
    _ok(H, rc)
    want = orc.ref_domq_encode(qual, off, lens) if orc.have_gz_ref() else orc.domq_encode(qual, off, lens)
    for k, b, l in zip(("qual", "runs", "mplx", "divr"), bufs, ls):
        assert l.value == want[k].size and np.array_equal(b[:l.value], want[k]), f"DOMQ stream {k} differs"
    assert prm.value == (want["num_norm_qs"] | 0x80) and np.array_equal(den[:dl.value], want["denorm"])
    if want["qual"].size >= 50:
        kind = "rans" if sub.startswith("RAN") else "arith"
        assert np.array_equal(comp[:cl.value], orc.compress("ref" if orc.have_ref() else "port", kind, want["qual"], orc.ORDER[sub])), "QUAL section differs"
    assert sf.value == 1, "the soft-fail re-entry of codec_domq_compress was not taken"
    exp = np.concatenate([qual[int(o): int(o) + int(l)] for o, l in zip(off, lens)]) if tot else np.zeros(0, np.uint8)
    assert np.array_equal(back[:tot], exp)
    assert int(cnt[0] + cnt[1]) == int((lens > 0).sum())


def pbwt_text(ht, big):
    """codec_pbwt_reconstruct's switch (codec_pbwt.c:416-448) over the cells, '|' / tab appended alternately by the caller"""
    out = bytearray()
    k = 0
    for a in ht.reshape(-1):
        if a == ord("*"):
            continue
        if ord("0") <= a <= ord("9") or a == ord("."):
            out.append(a)
        elif a == ord("-"):
            del out[-1:]
        elif a == ord("%"):
            if out[-1:] in (b"|", b"/"):
                out[-1:] = b"/."
            else:
                out.append(ord("."))
        elif a == ord("&"):
            out += str(big + 245).encode()
        else:
            out += str((int(a) - 48) & 0xff).encode()
        out.append(ord("\t") if k & 1 else ord("|"))
        k += 1
    return bytes(out)


def check_pbwt(H, ht):
    ht = np.ascontiguousarray(ht, np.uint8); nl, w = ht.shape
    runs = np.zeros(2 * ht.size + 64, np.uint32); fgrc = np.zeros(ht.size + 64, np.uint32); nr, nf = C.c_uint32(), C.c_uint32()
    back = np.zeros(ht.size + 64, np.uint8); text = np.zeros(ht.size * 5 + 64, np.uint8); tl = C.c_uint32()
    _ok(H, H.harness_pbwt(_p(ht), nl, w, _p(runs), C.byref(nr), _p(fgrc), C.byref(nf), _p(back), _p(text), C.byref(tl)))
    wr, wf = orc.ref_pbwt_encode(ht) if orc.have_gz_ref() else orc.pbwt_encode(ht)
    assert np.array_equal(runs[:nr.value], wr) and np.array_equal(fgrc[:nf.value], wf)
    assert np.array_equal(back[:ht.size].reshape(ht.shape), ht)
    assert bytes(text[:tl.value]) == pbwt_text(ht, 12)


def check_longr(H, txt, seq_off, qual_off, seq_len, qual_len, is_rev, sub="ARTW"):
    txt = np.ascontiguousarray(txt, np.uint8)
    so, qo = np.ascontiguousarray(seq_off, np.uint64), np.ascontiguousarray(qual_off, np.uint64)
    sl, ql = np.ascontiguousarray(seq_len, np.uint32), np.ascontiguousarray(qual_len, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    if orc.have_gz_ref():
        v2b, wv, wl = orc.ref_longr_encode(txt, so, qo, ql, rv, seq_lens=sl)
    else:
        parts = [txt[int(o): int(o) + int(l)] for o, l in zip(qo, ql) if l and not (l == 1 and txt[int(o)] == 32)]
        v2b = orc.longr_bins(np.concatenate(parts)); wv, wl = orc.longr_encode(txt, so, qo, ql, rv, v2b, seq_lens=sl)
    values = np.zeros(int(ql.sum()) + 64, np.uint8); nv = C.c_uint32(); lens_be = np.zeros(65536, np.uint32)
    comp = np.zeros(65536 * 8 + 70000, np.uint8); cl = C.c_uint32(); back = np.zeros(int(sl.sum()) + 64, np.uint8); nm = C.c_int()
    _ok(H, H.harness_longr(_p(txt), _p(so), _p(qo), _p(sl), _p(ql), None if rv is None else _p(rv), sl.size, _p(v2b), CODEC[sub],
                           _p(values), C.byref(nv), _p(lens_be), _p(comp), C.byref(cl), _p(back), C.byref(nm)))
    assert nv.value == wv.size and np.array_equal(values[:nv.value], wv) and np.array_equal(lens_be, wl)
    kind = "rans" if sub.startswith("RAN") else "arith"
    assert np.array_equal(comp[:cl.value], orc.compress("ref" if orc.have_ref() else "port", kind, wl.view(np.uint8), orc.ORDER[sub])), "LENS section differs"
    exp = bytearray(); missing = 0
    for o, l, q in zip(qo, sl, ql):
        if not l:
            continue
        if q == 1 and txt[int(o)] == 32 and l != 1 or (q == 1 and txt[int(o)] == 32):
            exp += b"*"; missing += 1
        else:
            exp += bytes(txt[int(o): int(o) + int(l)])
    assert bytes(back[:len(exp)]) == bytes(exp) and nm.value == missing


def check_normq(H, seed, n_lines, p_missing, sub="ARTB"):
    """codec_normq_compress / codec_normq_reconstruct through the plug-in layer: strands through the line callback, lines without quality"""
    from test_normq import _vb
    txt, off, zlen, rev, seq_len, missing = _vb(seed=seed, n_lines=n_lines, p_missing=p_missing)
    tot, otot = int(zlen.sum()), int(seq_len.sum())
    local = np.zeros(tot + 64, np.uint8); comp = np.zeros(2 * tot + 70000, np.uint8); back = np.zeros(otot + 64, np.uint8)
    ll, cl, bl, sf, nm, nl = C.c_uint64(), C.c_uint32(), C.c_uint64(), C.c_int(), C.c_int(), C.c_uint64()
    _ok(H, H.harness_normq(_p(txt), _p(off), _p(zlen), _p(seq_len), _p(rev), zlen.size, CODEC[sub], _p(local), C.byref(ll), _p(comp), C.byref(cl),
                           _p(back), C.byref(bl), C.byref(sf), C.byref(nm), C.byref(nl)))
    want = orc.ref_normq_encode(txt, off, zlen, rev) if orc.have_gz_ref() else orc.normq_encode(txt, off, zlen, rev)
    assert ll.value == want.size and np.array_equal(local[:want.size], want), "QUAL.local differs from codec_normq_compress's"
    if want.size >= 50:
        kind = "rans" if sub.startswith("RAN") else "arith"
        assert np.array_equal(comp[:cl.value], orc.compress("ref" if orc.have_ref() else "port", kind, want, orc.ORDER[sub])), "QUAL section differs"
    assert sf.value == 1, "the soft-fail re-entry of codec_normq_compress was not taken"
    text = orc.ref_normq_decode(want, seq_len, rev) if orc.have_gz_ref() else None
    if text is not None:
        assert bl.value == text.size and np.array_equal(back[:text.size], text), "reconstructed text differs from codec_normq_reconstruct's"
    assert nm.value == int(missing.sum()) and nl.value == zlen.size


def check_homp(H, mode, seed, n_lines, sub="ARTB"):
    """codec_homp_compress / codec_t0_compress and their reconstruct through the plug-in layer: the lines condensed in place with their lengths
    updated, the soft-fail re-entry, one reconstruct call per line"""
    from test_homp_t0 import ultima_like
    txt, so, sl, qo = ultima_like(n_lines, seed, mode, noise=0.1)
    tot = int(sl.sum())
    local = np.zeros(tot + 64, np.uint8); newl = np.zeros(sl.size + 1, np.uint32); comp = np.zeros(2 * tot + 70000, np.uint8); back = np.zeros(tot + 64, np.uint8)
    ll, cl, bl, sf, nm, nl = C.c_uint64(), C.c_uint32(), C.c_uint64(), C.c_int(), C.c_int(), C.c_uint64()
    _ok(H, H.harness_homp(mode, _p(txt), C.c_uint64(txt.size), _p(so), _p(qo), _p(sl), sl.size, CODEC[sub], _p(local), C.byref(ll), _p(newl), _p(comp), C.byref(cl),
                          _p(back), C.byref(bl), C.byref(sf), C.byref(nm), C.byref(nl)))
    want, want_len = orc.hp_condense(mode, txt, so, sl, qo, "ref" if orc.have_gz_ref() else "port")
    assert ll.value == want.size and np.array_equal(local[:want.size], want), "the condensed lines differ from the reference's"
    assert np.array_equal(newl[:sl.size], want_len), "the updated line lengths differ"
    if want.size >= 50:
        kind = "rans" if sub.startswith("RAN") else "arith"
        assert np.array_equal(comp[:cl.value], orc.compress("ref" if orc.have_ref() else "port", kind, want, orc.ORDER[sub])), "the section differs"
    assert sf.value == 1, "the soft-fail re-entry was not taken"
    text = np.concatenate([txt[int(o):int(o) + int(l)] for o, l in zip(so, sl)])
    assert bl.value == text.size and np.array_equal(back[:text.size], text), "reconstructed text differs"
    assert nm.value == 0 and nl.value == (sl.size if mode == 0 else 0)


def run_all(H, big):
    r = np.random.default_rng(3)
    # simple codecs: contiguous and line by line, with and without the soft-fail retry
    for i, codec in enumerate(CODEC):
        d = stream(("skew8", "qual", "u32le", "uniform256")[i % 4], 3000 if not big else 300000, 10 + i)
        check_simple(H, codec, d, tight=i & 1)
    lens = r.integers(0, 200, 60).astype(np.uint32)
    check_simple(H, "ARTb", stream("qual", int(lens.sum()), 4), lines=lens, tight=1)
    # ACGT: with exceptions, lower case, pure ACGT (acgt_no_x), line by line
    seq, _ = fastq_vb(300 if not big else 20000, 150, 5)
    check_acgt(H, seq)
    low = seq.copy(); low[::7] |= 0x20
    check_acgt(H, low, sub="ARTB")
    pure = seq.copy(); pure[pure == ord("N")] = ord("A")
    check_acgt(H, pure)
    check_acgt(H, seq[:150 * 40], lines=np.full(40, 150, np.uint32))
    # DOMQ: fixed and ragged lines, empty lines
    _, qual = fastq_vb(400 if not big else 30000, 150, 6)
    off, ln = line_table(qual.size // 150, 150)
    check_domq(H, qual, off, ln)
    ragged = r.integers(0, 260, 120).astype(np.uint32); ragged[::9] = 0
    o2 = np.concatenate([[0], np.cumsum(ragged[:-1], dtype=np.uint64)]).astype(np.uint64)
    check_domq(H, stream("qual", int(ragged.sum()) + 1, 8) + 0, o2, ragged, sub="RANB")
    # PBWT: bi-allelic, multi-allelic with the pseudo alleles
    check_pbwt(H, haplotype_matrix(60 if not big else 3000, 40 if not big else 1000, 7))
    m = haplotype_matrix(50, 30, 8, multi=True)
    m[3, 1::2][:5] = ord("-"); m[5, 4:8] = ord("%"); m[7, 9] = ord("&"); m[9, 0:6] = ord("*"); m[11, 2] = 58 + 5; m[12, 3] = 20
    check_pbwt(H, m)
    # LONGR: forward, reverse-complemented, lines without quality
    seq, q, lens = longread_vb(8 if not big else 60, 1200 if not big else 40000, 9)
    n = int(lens.sum()); txt = np.concatenate([seq, q])
    so = np.concatenate([[0], np.cumsum(lens[:-1], dtype=np.uint64)]).astype(np.uint64); qo = so + np.uint64(n)
    check_longr(H, txt, so, qo, lens, lens, None)
    rev = (np.arange(lens.size) % 2).astype(np.uint8)
    check_longr(H, txt, so, qo, lens, lens, rev, sub="RANW")
    t2 = txt.copy(); ql = lens.copy()
    for li in (1, 4):
        t2[int(qo[li])] = 32; ql[li] = 1
    check_longr(H, t2, so, qo, lens, ql, rev)
    # NORMQ: reverse-complemented reads, lines without quality
    check_normq(H, 21, 200 if not big else 20000, 0.0)
    check_normq(H, 22, 150 if not big else 5000, 0.15, sub="RANB")
    # HOMP and T0: lines condensed in place, lengths updated through the adapter
    check_homp(H, 0, 31, 120 if not big else 3000)
    check_homp(H, 1, 32, 120 if not big else 3000, sub="RANB")


def check_combiner(H):
    secs = [stream("qual", 2000 + 137 * i, 20 + i) for i in range(24)]
    data = np.concatenate(secs); lens = np.array([s.size for s in secs], np.uint32)
    caps = [orc.est_size("arith", s.size, orc.ORDER["ARTb"]) + 1024 for s in secs]
    off = np.concatenate([[0], np.cumsum(caps[:-1])]).astype(np.uint32)
    out = np.zeros(int(sum(caps)) + 64, np.uint8); ol = np.zeros(len(secs), np.uint32); nb = C.c_uint64()
    rc = H.harness_combine(CODEC["ARTb"], _p(data), _p(lens), len(secs), 20000, _p(out), _p(off), _p(ol), C.byref(nb))
    assert rc == 0
    for s, o, l in zip(secs, off, ol):
        assert np.array_equal(out[o:o + l], orc.compress("ref" if orc.have_ref() else "port", "arith", s, orc.ORDER["ARTb"]))
    assert nb.value < len(secs), f"{len(secs)} sections went in {nb.value} batches: nothing was combined"


def test_plugin_layer_on_the_emulator():
    H = harness("simt")
    run_all(H, big=False)
    check_combiner(H)
    H.harness_shutdown()


@pytest.mark.gpu
def test_plugin_layer_on_the_gpu(request):
    if request.config.getoption("--simt") or request.config.getoption("--dry-gpu"):
        pytest.skip("the CPU suite runs the harness on the emulator build (test_plugin_layer_on_the_emulator)")
    H = harness("gpu")
    run_all(H, big=True)
    H.harness_set_combining(1, 300)
    run_all(H, big=False)                                                   # the same, every simple section through the process-wide combiners
    H.harness_set_combining(0, 0)
    check_combiner(H)
    H.harness_shutdown()
