"""ctypes access to the CHECKERS (test infrastructure only):
   - oracle/liboracle.so        : our CPU restatement (oracle/*.c)
   - oracle/_ref/libhts_ref.so  : the reference's own htscodecs objects (built by oracle/Makefile here;
                                  travels prebuilt to the GPU box)
Nothing in genozip_b200/ imports this module."""
import ctypes as C, os, subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")

ORDER = {"RANB": 0x01, "RANW": 0x19, "RANb": 0x81, "RANw": 0x99,
         "ARTB": 0x01, "ARTW": 0x19, "ARTb": 0x81, "ARTw": 0x99}


def _build():
    subprocess.run(["make", "-s", "-C", ODIR, "all"], check=True)


_port = None
_ref = None
u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)


def port():
    global _port
    if _port is None:
        p = os.path.join(ODIR, "liboracle.so")
        if not os.path.exists(p):
            _build()
        L = C.CDLL(p)
        for nm in ("orc_rans_bound", "orc_arith_bound"):
            getattr(L, nm).restype = C.c_uint32
            getattr(L, nm).argtypes = [C.c_uint32, C.c_int]
        for nm in ("orc_rans_compress", "orc_arith_compress"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, u32p, C.c_int]
        for nm in ("orc_rans_uncompress", "orc_arith_uncompress"):
            getattr(L, nm).restype = C.c_int
            getattr(L, nm).argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, u32p]
        L.orc_acgt_packed_len.restype = C.c_uint64
        L.orc_acgt_packed_len.argtypes = [C.c_uint64]
        L.orc_acgt_pack.restype = C.c_int
        L.orc_acgt_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_acgt_unpack.restype = None
        L.orc_acgt_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_domq_prepare.restype = None
        L.orc_domq_prepare.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_domq_split.restype = None
        L.orc_domq_split.argtypes = [C.c_void_p] * 3 + [C.c_uint32] + [C.c_void_p] * 3 + [C.c_void_p, u32p] * 4
        L.orc_domq_reconstruct.restype = C.c_int
        L.orc_domq_reconstruct.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32,
                                           C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint8, C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_pbwt_encode.restype = C.c_int
        L.orc_pbwt_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, u32p, C.c_void_p, u32p]
        L.orc_pbwt_decode.restype = C.c_int
        L.orc_pbwt_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, u64p]
        L.orc_longr_calc_bins.restype = None
        L.orc_longr_calc_bins.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
        L.orc_longr_encode2.restype = C.c_int
        L.orc_longr_encode2.argtypes = [C.c_void_p] * 6 + [C.c_uint32] + [C.c_void_p] * 3
        L.orc_longr_encode.restype = C.c_int
        L.orc_longr_encode.argtypes = [C.c_void_p] * 5 + [C.c_uint32] + [C.c_void_p] * 3
        L.orc_longr_decode.restype = C.c_int
        L.orc_longr_decode.argtypes = [C.c_void_p] * 4 + [C.c_uint32] + [C.c_void_p] * 4
        _port = L
    return _port


def have_ref():
    return os.path.exists(os.path.join(ODIR, "_ref", "libhts_ref.so")) or os.path.isdir("/root/reference/src/htscodecs")


def ref():
    global _ref
    if _ref is None:
        p = os.path.join(ODIR, "_ref", "libhts_ref.so")
        if not os.path.exists(p):
            _build()
        L = C.CDLL(p)
        L.rans_compress_bound_4x16.restype = C.c_uint
        L.rans_compress_bound_4x16.argtypes = [C.c_uint, C.c_int]
        L.arith_compress_bound.restype = C.c_uint
        L.arith_compress_bound.argtypes = [C.c_uint, C.c_int]
        for nm in ("rans_compress_to_4x16", "arith_compress_to"):
            getattr(L, nm).restype = C.c_void_p
            getattr(L, nm).argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, u32p, C.c_int]
        for nm in ("rans_uncompress_to_4x16", "arith_uncompress_to"):
            getattr(L, nm).restype = C.c_void_p
            getattr(L, nm).argtypes = [C.c_void_p, C.c_void_p, C.c_uint, C.c_void_p, u32p]
        _ref = L
    return _ref


_gzref = None


def have_gz_ref():
    return os.path.exists(os.path.join(ODIR, "_ref", "libgz_ref.so")) or os.path.isdir("/root/reference/src/htscodecs")


def gz_ref():
    """the reference's own genozip codec objects (codec_domq.c ...) hosted by oracle/ref_gz_shim.c"""
    global _gzref
    if _gzref is None:
        p = os.path.join(ODIR, "_ref", "libgz_ref.so")
        if not os.path.exists(p):
            _build()
        L = C.CDLL(p)
        L.ref_domq_encode.restype = C.c_int
        L.ref_domq_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32] + [C.c_void_p, u32p] * 5 + [C.POINTER(C.c_uint8)] * 2
        L.ref_gz_last_error.restype = C.c_char_p
        L.ref_acgt_pack.restype = C.c_int
        L.ref_acgt_pack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, u64p, C.c_void_p, C.POINTER(C.c_int)]
        L.ref_pbwt_encode.restype = C.c_int
        L.ref_pbwt_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, u32p, C.c_void_p, u32p]
        L.ref_longr_encode2.restype = C.c_int
        L.ref_longr_encode2.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_longr_encode.restype = C.c_int
        L.ref_longr_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        for f in ("ref_acgt_unpack", "ref_pbwt_decode", "ref_domq_decode", "ref_longr_decode"):
            getattr(L, f).restype = C.c_int
        L.ref_acgt_unpack.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.ref_pbwt_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, u64p]
        L.ref_domq_decode.argtypes = [C.c_void_p, C.c_uint32] * 5 + [C.c_uint8, C.c_void_p, C.c_uint32, C.c_void_p]
        L.ref_longr_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _gzref = L
    return _gzref


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def est_size(kind, n, order):
    """codec_*_est_size (codec_htscodecs.c:26-33) = 1 KB + bound"""
    L = port()
    return 1024 + (L.orc_rans_bound(n, order) if kind == "rans" else L.orc_arith_bound(n, order))


def compress(impl, kind, data, order):
    data = np.ascontiguousarray(data, dtype=np.uint8)
    n = data.size
    cap = est_size(kind, n, order)
    out = np.zeros(cap + 16, dtype=np.uint8)
    ol = C.c_uint32(cap)
    src = data if n else np.zeros(1, np.uint8)
    if impl == "port":
        f = port().orc_rans_compress if kind == "rans" else port().orc_arith_compress
        rc = f(_ptr(src), n, _ptr(out), C.byref(ol), order)
        assert rc == 0
    else:
        f = ref().rans_compress_to_4x16 if kind == "rans" else ref().arith_compress_to
        r = f(None, _ptr(src), n, _ptr(out), C.byref(ol), order)
        assert r
    return out[:ol.value].copy()


def uncompress(impl, kind, comp, n):
    comp = np.ascontiguousarray(comp, dtype=np.uint8)
    out = np.zeros(max(n, 1) + 16, dtype=np.uint8)
    ol = C.c_uint32(n)
    if impl == "port":
        f = port().orc_rans_uncompress if kind == "rans" else port().orc_arith_uncompress
        rc = f(_ptr(comp), comp.size, _ptr(out), C.byref(ol))
        assert rc == 0, "port uncompress failed"
    else:
        f = ref().rans_uncompress_to_4x16 if kind == "rans" else ref().arith_uncompress_to
        r = f(None, _ptr(comp), comp.size, _ptr(out), C.byref(ol))
        assert r, "ref uncompress failed"
    assert ol.value == n
    return out[:n].copy()


# ------------------------------------------------------------------ genozip-specific codecs (oracle/gz_port.c)
class DomqTables(C.Structure):
    _fields_ = [("n_lines", C.c_uint32), ("num_norm_qs", C.c_uint8), ("num_doms", C.c_uint8), ("has_diverse", C.c_uint8),
                ("denorm", C.c_uint8 * (95 * 95)), ("normalize", C.c_uint8 * (95 * 95))]


def acgt_pack(seq):
    seq = np.ascontiguousarray(seq, np.uint8)
    L = port()
    packed = np.zeros(L.orc_acgt_packed_len(seq.size), np.uint8)
    x = np.zeros(max(seq.size, 1), np.uint8)
    allz = L.orc_acgt_pack(_ptr(seq if seq.size else x), seq.size, _ptr(packed if packed.size else x), _ptr(x))
    return packed, x[:seq.size], bool(allz)


def acgt_unpack(packed, x, n):
    out = np.zeros(max(n, 1), np.uint8)
    port().orc_acgt_unpack(_ptr(packed if packed.size else out), None if x is None else _ptr(x), n, _ptr(out))
    return out[:n]


def domq_encode(txt, off, lens):
    """-> dict(tables, line_dom, line_diverse, qual, runs, mplx, divr)"""
    L = port()
    txt = np.ascontiguousarray(txt, np.uint8); off = np.ascontiguousarray(off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    n = lens.size
    t = DomqTables()
    dom = np.zeros(max(n, 1), np.uint8); div = np.zeros(max(n, 1), np.uint8)
    tx = txt if txt.size else np.zeros(1, np.uint8)
    L.orc_domq_prepare(_ptr(tx), _ptr(off), _ptr(lens), n, _ptr(dom), _ptr(div), C.byref(t))
    tot = int(lens.sum())
    qual = np.zeros(2 * tot + 2, np.uint8); runs = np.zeros(tot + 2, np.uint8)
    mplx = np.zeros(n + 1, np.uint8); divr = np.zeros(tot + 1, np.uint8)
    ql, rl, ml, dl = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32()
    L.orc_domq_split(_ptr(tx), _ptr(off), _ptr(lens), n, _ptr(dom), _ptr(div), C.byref(t),
                     _ptr(qual), C.byref(ql), _ptr(runs), C.byref(rl), _ptr(mplx), C.byref(ml), _ptr(divr), C.byref(dl))
    nn, nd = t.num_norm_qs, t.num_doms
    return dict(num_norm_qs=nn, num_doms=nd, has_diverse=t.has_diverse,
                denorm=np.frombuffer(bytes(t.denorm), np.uint8)[:nd * nn].copy(),
                normalize=np.frombuffer(bytes(t.normalize), np.uint8).copy(),
                line_dom=dom[:n].copy(), line_diverse=div[:n].copy(),
                qual=qual[:ql.value].copy(), runs=runs[:rl.value].copy(), mplx=mplx[:ml.value].copy(), divr=divr[:dl.value].copy())


def ref_domq_encode(txt, off, lens):
    """codec_domq_comp_init (forced) + codec_domq_compress of the REFERENCE's compiled codec_domq.c on these lines
    -> dict(qual, runs, mplx, divr, denorm, num_norm_qs, has_diverse)"""
    L = gz_ref()
    txt = np.ascontiguousarray(txt, np.uint8); off = np.ascontiguousarray(off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    tot = int(lens.sum())
    qual = np.zeros(2 * tot + 16, np.uint8); runs = np.zeros(tot + 16, np.uint8); mplx = np.zeros(lens.size + 16, np.uint8)
    divr = np.zeros(tot + 16, np.uint8); den = np.zeros(95 * 95 * 2, np.uint8)
    ql, rl, ml, dl, nl = (C.c_uint32() for _ in range(5)); prm, hd = C.c_uint8(), C.c_uint8()
    tx = txt if txt.size else np.zeros(1, np.uint8)
    rc = L.ref_domq_encode(_ptr(tx), txt.size, _ptr(off), _ptr(lens), lens.size, _ptr(qual), C.byref(ql), _ptr(runs), C.byref(rl),
                           _ptr(mplx), C.byref(ml), _ptr(divr), C.byref(dl), _ptr(den), C.byref(nl), C.byref(prm), C.byref(hd))
    assert rc == 0, f"reference codec_domq aborted ({rc}): {L.ref_gz_last_error().decode()}"
    assert prm.value & 0x80                                                 # MSb set since 14.0.5 (codec_domq.c:234)
    return dict(qual=qual[:ql.value].copy(), runs=runs[:rl.value].copy(), mplx=mplx[:ml.value].copy(), divr=divr[:dl.value].copy(),
                denorm=den[:nl.value].copy(), num_norm_qs=prm.value & 0x7f, has_diverse=hd.value)


def _gz_check(rc, what):
    assert rc == 0, f"reference {what} failed ({rc}): {gz_ref().ref_gz_last_error().decode()}"


def ref_acgt_pack(seq):
    """the REFERENCE's compiled codec_acgt_compress up to its sub-codec call -> (packed LE words, exception stream, acgt_no_x)"""
    seq = np.ascontiguousarray(seq, np.uint8)
    packed = np.zeros(seq.size // 4 + 64, np.uint8); x = np.zeros(seq.size + 8, np.uint8)
    pl, nox = C.c_uint64(), C.c_int()
    _gz_check(gz_ref().ref_acgt_pack(_ptr(seq), seq.size, _ptr(packed), C.byref(pl), _ptr(x), C.byref(nox)), "codec_acgt")
    return packed[:pl.value].copy(), x[:seq.size].copy(), bool(nox.value)


def ref_pbwt_encode(ht):
    """the REFERENCE's compiled codec_pbwt_compress -> (RUNS u32, FGRC u32), host-endian"""
    ht = np.ascontiguousarray(ht, np.uint8)
    n_lines, w = ht.shape
    runs = np.zeros(2 * ht.size + 8, np.uint32); fgrc = np.zeros(ht.size + 8, np.uint32)
    nr, nf = C.c_uint32(), C.c_uint32()
    _gz_check(gz_ref().ref_pbwt_encode(_ptr(ht), n_lines, w, _ptr(runs), C.byref(nr), _ptr(fgrc), C.byref(nf)), "codec_pbwt")
    return runs[:nr.value].copy(), fgrc[:nf.value].copy()


def ref_longr_encode(txt, seq_off, qual_off, lens, is_rev, seq_lens=None):
    """the REFERENCE's compiled codec_longr_segconf_calculate_bins + codec_longr_compress -> (value_to_bin, values, lens_be);
    lens = quality lengths, seq_lens = sequence lengths where they differ (lines without quality)"""
    txt = np.ascontiguousarray(txt, np.uint8); seq_off = np.ascontiguousarray(seq_off, np.uint64)
    qual_off = np.ascontiguousarray(qual_off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    tot = int(lens.sum())
    v2b = np.zeros(256, np.uint8); values = np.zeros(tot + 8, np.uint8); lens_be = np.zeros(65536, np.uint32)
    sl = None if seq_lens is None else np.ascontiguousarray(seq_lens, np.uint32)
    _gz_check(gz_ref().ref_longr_encode2(_ptr(txt), txt.size, _ptr(seq_off), _ptr(qual_off), _ptr(lens), None if sl is None else _ptr(sl),
                                         None if rv is None else _ptr(rv), lens.size, _ptr(v2b), _ptr(values), _ptr(lens_be)), "codec_longr")
    return v2b, values[:tot].copy(), lens_be


def ref_acgt_unpack(packed, x, n):
    """the REFERENCE's compiled codec_acgt_uncompress (+ codec_xcgt_uncompress when x is given) -> the n bases"""
    packed = np.ascontiguousarray(packed, np.uint8)
    xx = None if x is None else np.ascontiguousarray(x, np.uint8)
    seq = np.zeros(n + 8, np.uint8)
    _gz_check(gz_ref().ref_acgt_unpack(_ptr(packed), packed.size, None if xx is None else _ptr(xx), n, _ptr(seq)), "codec_acgt_uncompress")
    return seq[:n].copy()


def ref_pbwt_decode(runs, fgrc, n_lines, size):
    """the REFERENCE's compiled codec_pbwt_uncompress; runs/fgrc host-endian as from pbwt_encode (FGRC is handed over as stored: big-endian)"""
    r = np.ascontiguousarray(runs, np.uint32); f = np.ascontiguousarray(fgrc, np.uint32).byteswap()
    ht = np.zeros(size + 64, np.uint8); hl = C.c_uint64()
    _gz_check(gz_ref().ref_pbwt_decode(_ptr(r), r.size, _ptr(f), f.size, n_lines, _ptr(ht), C.byref(hl)), "codec_pbwt_uncompress")
    assert hl.value == size, (hl.value, size)
    return ht[:size].copy()


def ref_domq_decode(enc, lens):
    """the REFERENCE's compiled codec_domq_reconstruct, line by line -> all quality strings concatenated"""
    lens = np.ascontiguousarray(lens, np.uint32)
    out = np.zeros(int(lens.sum()) + 8, np.uint8)
    a = [np.ascontiguousarray(enc[k], np.uint8) for k in ("qual", "runs", "mplx", "divr", "denorm")]
    z = np.zeros(1, np.uint8)
    args = []
    for v in a:
        args += [_ptr(v if v.size else z), v.size]
    _gz_check(gz_ref().ref_domq_decode(*args, 0x80 | int(enc["num_norm_qs"]), _ptr(lens), lens.size, _ptr(out)), "codec_domq_reconstruct")
    return out[:int(lens.sum())].copy()


def ref_longr_decode(txt, seq_off, lens, is_rev, v2b, values, lens_be):
    """the REFERENCE's compiled codec_longr_reconstruct, read by read"""
    txt = np.ascontiguousarray(txt, np.uint8); seq_off = np.ascontiguousarray(seq_off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    tot = int(lens.sum())
    out = np.zeros(tot + 8, np.uint8)
    v = np.ascontiguousarray(values, np.uint8) if values.size else np.zeros(1, np.uint8)
    lb = np.ascontiguousarray(lens_be, np.uint32); vb = np.ascontiguousarray(v2b, np.uint8)
    _gz_check(gz_ref().ref_longr_decode(_ptr(txt), _ptr(seq_off), _ptr(lens), None if rv is None else _ptr(rv), lens.size, _ptr(vb), _ptr(v), _ptr(lb),
                                        _ptr(out)), "codec_longr_reconstruct")
    return out[:tot].copy()


def domq_decode(enc, lens):
    L = port()
    lens = np.ascontiguousarray(lens, np.uint32)
    out = np.zeros(int(lens.sum()) + 1, np.uint8)
    runs = enc["runs"].copy() if enc["runs"].size else np.zeros(1, np.uint8)
    z = np.zeros(1, np.uint8)
    g = lambda a: _ptr(a if a.size else z)
    rc = L.orc_domq_reconstruct(g(enc["qual"]), enc["qual"].size, _ptr(runs), enc["runs"].size, g(enc["mplx"]), enc["mplx"].size,
                                g(enc["divr"]), enc["divr"].size, g(enc["denorm"]), enc["num_norm_qs"], _ptr(lens), lens.size, _ptr(out))
    assert rc == 0, f"orc_domq_reconstruct rc={rc}"
    return out[:-1]


def pbwt_encode(ht):
    ht = np.ascontiguousarray(ht, np.uint8)
    n_lines, w = ht.shape
    runs = np.zeros(2 * ht.size + 4, np.uint32); fgrc = np.zeros(ht.size + 4, np.uint32)
    nr, nf = C.c_uint32(), C.c_uint32()
    rc = port().orc_pbwt_encode(_ptr(ht), n_lines, w, _ptr(runs), C.byref(nr), _ptr(fgrc), C.byref(nf))
    assert rc == 0
    return runs[:nr.value].copy(), fgrc[:nf.value].copy()


def pbwt_decode(runs, fgrc, n_lines, size):
    ht = np.zeros(size, np.uint8)
    hl = C.c_uint64()
    r2, f2 = runs.copy(), fgrc.copy()
    rc = port().orc_pbwt_decode(_ptr(r2), r2.size, _ptr(f2), f2.size, n_lines, _ptr(ht), C.byref(hl))
    assert rc == 0 and hl.value == size
    return ht


def longr_bins(qual):
    hist = np.bincount(np.asarray(qual, np.uint8) - 33, minlength=256).astype(np.uint32)
    v2b = np.zeros(256, np.uint8)
    port().orc_longr_calc_bins(_ptr(hist), int(hist.sum()), _ptr(v2b))
    return v2b


def longr_encode(txt, seq_off, qual_off, lens, is_rev, v2b, seq_lens=None):
    n = lens.size
    tot = int(lens.sum())
    values = np.zeros(tot + 1, np.uint8); lens_be = np.zeros(65536, np.uint32)
    sl = None if seq_lens is None else np.ascontiguousarray(seq_lens, np.uint32)
    rc = port().orc_longr_encode2(_ptr(txt), _ptr(seq_off), _ptr(qual_off), _ptr(lens), None if sl is None else _ptr(sl),
                                  None if is_rev is None else _ptr(is_rev), n, _ptr(v2b), _ptr(values), _ptr(lens_be))
    assert rc == 0
    return values[:tot].copy(), lens_be


def longr_decode_lines(txt, seq_off, lens, is_rev, v2b, values, lens_be):
    """the decoder with the missing-quality rule -> (qualities, missing flag per line)"""
    L = port()
    L.orc_longr_decode2.restype = C.c_int
    L.orc_longr_decode2.argtypes = [C.c_void_p] * 4 + [C.c_uint32] + [C.c_void_p] * 5
    tot = int(lens.sum())
    out = np.zeros(tot + 1, np.uint8); miss = np.zeros(lens.size + 1, np.uint8)
    v = np.ascontiguousarray(values, np.uint8) if values.size else np.zeros(1, np.uint8)
    rc = L.orc_longr_decode2(_ptr(txt), _ptr(seq_off), _ptr(lens), None if is_rev is None else _ptr(is_rev), lens.size,
                             _ptr(v2b), _ptr(v), _ptr(np.ascontiguousarray(lens_be, np.uint32)), _ptr(out), _ptr(miss))
    assert rc == 0
    return out[:tot], miss[:lens.size]


def longr_decode(txt, seq_off, lens, is_rev, v2b, values, lens_be):
    tot = int(lens.sum())
    out = np.zeros(tot + 1, np.uint8)
    v = values if values.size else np.zeros(1, np.uint8)
    rc = port().orc_longr_decode(_ptr(txt), _ptr(seq_off), _ptr(lens), None if is_rev is None else _ptr(is_rev), lens.size,
                                 _ptr(v2b), _ptr(v), _ptr(lens_be), _ptr(out))
    assert rc == 0
    return out[:tot]


# ---------------------------------------------------------------- NORMQ (src/codec_normq.c)
def normq_encode(txt, off, lens, is_rev):
    """restatement of codec_normq_compress before its sub-codec -> QUAL.local"""
    L = port()
    L.orc_normq_encode.restype = C.c_uint64
    L.orc_normq_encode.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p]
    txt = np.ascontiguousarray(txt, np.uint8); off = np.ascontiguousarray(off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    out = np.zeros(int(lens.sum()) + 8, np.uint8)
    n = L.orc_normq_encode(_ptr(txt), _ptr(off), _ptr(lens), None if rv is None else _ptr(rv), lens.size, _ptr(out))
    return out[:n].copy()


def normq_decode(local, lens, is_rev):
    """restatement of codec_normq_reconstruct for every line -> (len[i] bytes per line, missing flags); None if the stream does not fit the lines"""
    L = port()
    L.orc_normq_decode.restype = C.c_int
    L.orc_normq_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    local = np.ascontiguousarray(local, np.uint8); lens = np.ascontiguousarray(lens, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    out = np.zeros(int(lens.sum()) + 8, np.uint8); miss = np.zeros(lens.size + 1, np.uint8); used = C.c_uint64()
    lo = local if local.size else np.zeros(1, np.uint8)
    rc = L.orc_normq_decode(_ptr(lo), local.size, _ptr(lens), None if rv is None else _ptr(rv), lens.size, _ptr(out), _ptr(miss), C.byref(used))
    if rc != 0 or used.value != local.size:
        return None
    return out[:int(lens.sum())].copy(), miss[:lens.size].copy()


def ref_normq_encode(txt, off, lens, is_rev):
    """the REFERENCE's compiled codec_normq_compress -> what it hands its sub-codec"""
    txt = np.ascontiguousarray(txt, np.uint8); off = np.ascontiguousarray(off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    out = np.zeros(int(lens.sum()) + 1100, np.uint8); n = C.c_uint64()
    G = gz_ref()
    G.ref_normq_encode.restype = C.c_int
    G.ref_normq_encode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    _gz_check(G.ref_normq_encode(_ptr(txt), txt.size, _ptr(off), _ptr(lens), None if rv is None else _ptr(rv), lens.size, _ptr(out), C.byref(n)), "codec_normq_compress")
    return out[:n.value].copy()


def ref_normq_decode(local, lens, is_rev):
    """the REFERENCE's compiled codec_normq_reconstruct, line by line -> txt_data (a line without quality is the one character '*')"""
    local = np.ascontiguousarray(local, np.uint8); lens = np.ascontiguousarray(lens, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    out = np.zeros(int(lens.sum()) + 64, np.uint8); n = C.c_uint64()
    G = gz_ref()
    G.ref_normq_decode.restype = C.c_int
    G.ref_normq_decode.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    lo = local if local.size else np.zeros(1, np.uint8)
    _gz_check(G.ref_normq_decode(_ptr(lo), local.size, _ptr(lens), None if rv is None else _ptr(rv), lens.size, _ptr(out), C.byref(n)), "codec_normq_reconstruct")
    return out[:n.value].copy()


# ---------------------------------------------------------------- zip_generate_local's transforms (src/zip.c:167-213)
LT_OPS = {"swap16": 1, "swap32": 2, "swap64": 3, "interlace8": 4, "interlace16": 5, "interlace32": 6, "interlace64": 7,
          "deinterlace8": 8, "deinterlace16": 9, "deinterlace32": 10, "deinterlace64": 11}


def local_transform(op, a):
    """numpy restatement: byte swap; INTERLACE (n >= 0 -> 2n, n < 0 -> -2n - 1, modulo the width) then big endian; and the inverse"""
    a = np.ascontiguousarray(a).copy()
    w = a.dtype.itemsize
    U = {1: np.uint8, 2: np.uint16, 4: np.uint32, 8: np.uint64}[w]
    u = a.view(U)
    if op.startswith("swap"):
        return u.byteswap().view(a.dtype)
    if op.startswith("interlace"):
        s_ = u.view({1: np.int8, 2: np.int16, 4: np.int32, 8: np.int64}[w])
        with np.errstate(over="ignore"):
            neg = ((U(0) - u) << U(1)) - U(1)
            pos = u << U(1)
        return np.where(s_ < 0, neg, pos).astype(U).byteswap().view(a.dtype)
    x = u.byteswap()
    with np.errstate(over="ignore"):
        odd = U(0) - ((x >> U(1)) + U(1))
        even = x >> U(1)
    return np.where(x & U(1), odd, even).astype(U).view(a.dtype)


def ref_local_transform(op, a):
    """the reference's own INTERLACE / DEINTERLACE / BGEN macros in buffer.c's loops (oracle/ref_gz_shim.c)"""
    b = np.ascontiguousarray(a).copy()
    G = gz_ref()
    G.ref_local_transform.restype = C.c_int
    G.ref_local_transform.argtypes = [C.c_int, C.c_void_p, C.c_uint64]
    if b.size:
        _gz_check(G.ref_local_transform(LT_OPS[op], _ptr(b), b.size), "local transform")
    return b


# ---------------------------------------------------------------- OQ (src/codec_oq.c)
def _oq_args(txt, qoff, qlen):
    return (np.ascontiguousarray(txt, np.uint8), np.ascontiguousarray(qoff, np.uint64), np.ascontiguousarray(qlen, np.uint32))


def oq_mux(txt, qoff, qlen, ooff, seq_len=None, lib="port"):
    """codec_oq_compress before its sub-codec -> (channels back to back, count[94], monochars[94]); None on an invalid QUAL character.
    lib = "port" (the restatement, oracle/gz_port.c) or "ref" (the reference's compiled codec_oq.c, oracle/_ref)"""
    txt, qoff, qlen = _oq_args(txt, qoff, qlen)
    ooff = np.ascontiguousarray(ooff, np.uint64)
    sl = None if seq_len is None else np.ascontiguousarray(seq_len, np.uint32)
    chan = np.zeros(int(qlen.sum()) + 8, np.uint8); count = np.zeros(94, np.uint32); mono = np.zeros(94, np.uint8)
    if lib == "port":
        L = port()
        L.orc_oq_mux.restype = C.c_int
        L.orc_oq_mux.argtypes = [C.c_void_p] * 5 + [C.c_uint32] + [C.c_void_p] * 3
        rc = L.orc_oq_mux(_ptr(txt), _ptr(qoff), _ptr(qlen), _ptr(ooff), None if sl is None else _ptr(sl), qlen.size, _ptr(chan), _ptr(count), _ptr(mono))
    else:
        L = gz_ref()
        L.ref_oq_encode.restype = C.c_int
        L.ref_oq_encode.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 4 + [C.c_uint32] + [C.c_void_p] * 3
        sl2 = qlen if sl is None else sl
        rc = L.ref_oq_encode(_ptr(txt), txt.size, _ptr(qoff), _ptr(qlen), _ptr(ooff), _ptr(sl2), qlen.size, _ptr(chan), _ptr(count), _ptr(mono))
    if rc != 0:
        return None
    return chan[:int(count.sum())].copy(), count, mono


def oq_demux(txt, qoff, qlen, out_off, out_size, key_bias, channels, count, mono, lib="port"):
    """codec_oq_reconstruct for every line -> out; None when a channel runs out of data"""
    txt, qoff, qlen = _oq_args(txt, qoff, qlen)
    ooff = np.ascontiguousarray(out_off, np.uint64)
    ch = np.ascontiguousarray(channels, np.uint8); ch = ch if ch.size else np.zeros(1, np.uint8)
    count = np.ascontiguousarray(count, np.uint32); mono = np.ascontiguousarray(mono, np.uint8)
    out = np.zeros(out_size + 8, np.uint8)
    if lib == "port":
        L = port()
        L.orc_oq_demux.restype = C.c_int
        L.orc_oq_demux.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_uint32] + [C.c_void_p] * 4
        rc = L.orc_oq_demux(_ptr(txt), _ptr(qoff), _ptr(qlen), _ptr(ooff), qlen.size, key_bias, _ptr(ch), _ptr(count), _ptr(mono), _ptr(out))
    else:
        L = gz_ref()
        L.ref_oq_decode.restype = C.c_int
        L.ref_oq_decode.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint32, C.c_uint32] + [C.c_void_p] * 3 + [C.c_void_p, C.c_uint64]
        rc = L.ref_oq_decode(_ptr(txt), txt.size, _ptr(qoff), _ptr(qlen), _ptr(ooff), qlen.size, key_bias, _ptr(ch), _ptr(count), _ptr(mono), _ptr(out), out_size)
    if rc != 0:
        return None
    return out[:out_size].copy()


# ---------------------------------------------------------------- dyn_int_transpose / BGEN_transpose_u*_buf
def local_transpose(a, cols, piz=False):
    """restatement -> (array, transposed flag)"""
    L = port()
    L.orc_local_transpose.restype = C.c_int
    L.orc_local_transpose.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_int]
    a = np.ascontiguousarray(a).copy()
    rc = L.orc_local_transpose(_ptr(a) if a.size else None, a.size, a.dtype.itemsize, cols, 1 if piz else 0)
    assert rc >= 0
    return a, bool(rc)


def ref_dyn_int_transpose(a, cols, cols_vcf=0):
    """the reference's compiled dyn_int_transpose (oracle/_ref) -> (array, transposed flag)"""
    L = gz_ref()
    L.ref_dyn_int_transpose.restype = C.c_int
    L.ref_dyn_int_transpose.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    a = np.ascontiguousarray(a).copy()
    tr = C.c_int(0)
    rc = L.ref_dyn_int_transpose(_ptr(a), a.size, a.dtype.itemsize, cols, cols_vcf, C.byref(tr))
    assert rc == 0, rc
    return a, bool(tr.value)


# ---------------------------------------------------------------- b250_zip_generate (src/b250.c:202-297)
def b250_generate(b250, ni2wi, ol_len, one_up_ok, lib="port"):
    """-> (converted buffer, n_words) or None (the words do not tile the buffer / a word index cannot be encoded).
    lib = "ref": the reference's compiled b250.c — it decides one_up_ok itself (nodes.len + ol_nodes.len > 1024, :247)"""
    b = np.ascontiguousarray(b250, np.uint8); t = np.ascontiguousarray(ni2wi, np.int32)
    out = np.zeros(b.size + 8, np.uint8)
    bp = _ptr(b) if b.size else _ptr(out); tp = _ptr(t) if t.size else _ptr(out)
    if lib == "port":
        L = port()
        L.orc_b250_generate.restype = C.c_int64
        L.orc_b250_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p]
        nw = C.c_uint64()
        n = L.orc_b250_generate(bp, b.size, tp, t.size, ol_len, 1 if one_up_ok else 0, _ptr(out), C.byref(nw))
        if n < 0:
            return None
        return out[b.size - n:b.size].copy(), int(nw.value)
    L = gz_ref()
    L.ref_b250_generate.restype = C.c_int
    L.ref_b250_generate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    assert one_up_ok == (t.size + ol_len > 1024)
    n = C.c_uint64()
    rc = L.ref_b250_generate(bp, b.size, tp, t.size, ol_len, _ptr(out), C.byref(n))
    if rc != 0:
        return None
    return out[:n.value].copy(), None


# ---------------------------------------------------------------- HOMP / T0 (src/codec_homp.c, src/codec_t0.c)
def hp_condense(mode, txt, str_off, str_len, seq_off, lib="port"):
    """-> (condensed strings back to back, new length of every line)"""
    txt = np.ascontiguousarray(txt, np.uint8); so = np.ascontiguousarray(str_off, np.uint64); sl = np.ascontiguousarray(str_len, np.uint32)
    qo = np.ascontiguousarray(seq_off, np.uint64)
    out = np.zeros(int(sl.sum()) + 8, np.uint8); nl = np.zeros(sl.size + 1, np.uint32)
    if lib == "port":
        L = port()
        L.orc_hp_condense.restype = C.c_uint64
        L.orc_hp_condense.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p]
        n = L.orc_hp_condense(mode, _ptr(txt), _ptr(so), _ptr(sl), _ptr(qo), sl.size, _ptr(out), _ptr(nl))
    else:
        L = gz_ref()
        L.ref_hp_condense.restype = C.c_int
        L.ref_hp_condense.argtypes = [C.c_int, C.c_void_p, C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        ln = C.c_uint64()
        rc = L.ref_hp_condense(mode, _ptr(txt), txt.size, _ptr(so), _ptr(sl), _ptr(qo), sl.size, _ptr(out), C.byref(ln), _ptr(nl))
        assert rc == 0, rc
        n = ln.value
    return out[:n].copy(), nl[:sl.size].copy()


def hp_expand(mode, local, txt, seq_off, lens, lib="port"):
    """-> (lens[i] bytes per line back to back, missing flags or None); None when the stream does not match the lines"""
    local = np.ascontiguousarray(local, np.uint8); txt = np.ascontiguousarray(txt, np.uint8)
    qo = np.ascontiguousarray(seq_off, np.uint64); sl = np.ascontiguousarray(lens, np.uint32)
    out = np.zeros(int(sl.sum()) + 8, np.uint8); lo = local if local.size else np.zeros(1, np.uint8)
    if lib == "port":
        L = port()
        L.orc_hp_expand.restype = C.c_int
        L.orc_hp_expand.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        miss = np.zeros(sl.size + 1, np.uint8)
        rc = L.orc_hp_expand(mode, _ptr(lo), local.size, _ptr(txt), _ptr(qo), _ptr(sl), sl.size, _ptr(out), _ptr(miss))
        return None if rc != 0 else (out[:int(sl.sum())].copy(), miss[:sl.size].copy())
    L = gz_ref()
    L.ref_hp_expand.restype = C.c_int
    L.ref_hp_expand.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    n = C.c_uint64()
    rc = L.ref_hp_expand(mode, _ptr(lo), local.size, _ptr(txt), _ptr(qo), _ptr(sl), sl.size, _ptr(out), C.byref(n))
    return None if rc != 0 else (out[:int(sl.sum())].copy(), None)


# ---------------------------------------------------------------- SMUX (src/codec_smux.c)
def smux_mux(txt, qoff, qlen, soff, slen, is_rev=None, lib="port"):
    """codec_smux_compress -> (the 5 channels back to back, count[5], n_param)"""
    txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32)
    soff = np.ascontiguousarray(soff, np.uint64); slen = np.ascontiguousarray(slen, np.uint32)
    rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    chan = np.zeros(int(qlen.sum()) + 8, np.uint8); count = np.zeros(5, np.uint32); par = C.c_uint8(0)
    rvp = None if rv is None else _ptr(rv)
    if lib == "port":
        L = port()
        L.orc_smux_mux.restype = C.c_int
        L.orc_smux_mux.argtypes = [C.c_void_p] * 6 + [C.c_uint32] + [C.c_void_p] * 3
        rc = L.orc_smux_mux(_ptr(txt), _ptr(qoff), _ptr(qlen), _ptr(soff), _ptr(slen), rvp, qlen.size, _ptr(chan), _ptr(count), C.byref(par))
    else:
        L = gz_ref()
        L.ref_smux_mux.restype = C.c_int
        L.ref_smux_mux.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 5 + [C.c_uint32] + [C.c_void_p] * 3
        rc = L.ref_smux_mux(_ptr(txt), txt.size, _ptr(qoff), _ptr(qlen), _ptr(soff), _ptr(slen), rvp, qlen.size, _ptr(chan), _ptr(count), C.byref(par))
    assert rc == 0, rc
    return chan[:int(count.sum())].copy(), count, int(par.value)


def smux_demux(txt, soff, lens, is_rev, out_off, out_size, channels, count, n_param, lib="port"):
    """codec_smux_reconstruct for every line -> out (None when a channel runs out of data)"""
    txt = np.ascontiguousarray(txt, np.uint8); soff = np.ascontiguousarray(soff, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
    ooff = np.ascontiguousarray(out_off, np.uint64); rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
    ch = np.ascontiguousarray(channels, np.uint8); ch = ch if ch.size else np.zeros(1, np.uint8); count = np.ascontiguousarray(count, np.uint32)
    out = np.zeros(out_size + 8, np.uint8); rvp = None if rv is None else _ptr(rv)
    if lib == "port":
        L = port()
        L.orc_smux_demux.restype = C.c_int
        L.orc_smux_demux.argtypes = [C.c_void_p] * 5 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint8, C.c_void_p, C.c_void_p]
        rc = L.orc_smux_demux(_ptr(txt), _ptr(soff), _ptr(lens), rvp, _ptr(ooff), lens.size, _ptr(ch), _ptr(count), n_param, _ptr(out), None)
    else:
        L = gz_ref()
        L.ref_smux_demux.restype = C.c_int
        L.ref_smux_demux.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint8, C.c_void_p, C.c_uint64]
        rc = L.ref_smux_demux(_ptr(txt), txt.size, _ptr(soff), _ptr(lens), rvp, _ptr(ooff), lens.size, _ptr(ch), _ptr(count), n_param, _ptr(out), out_size)
    return None if rc != 0 else out[:out_size].copy()


# ---------------------------------------------------------------- TMPL (src/codec_tmpl.c)
def tmpl_mux(txt, qoff, qlen, tmpl):
    """restatement of codec_tmpl_compress -> (channels 0..93 and the excess back to back, count[95])"""
    L = port()
    L.orc_tmpl_mux.restype = C.c_int
    L.orc_tmpl_mux.argtypes = [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32)
    tmpl = np.ascontiguousarray(tmpl, np.uint8); t = tmpl if tmpl.size else np.zeros(1, np.uint8)
    chan = np.zeros(int(qlen.sum()) + 8, np.uint8); count = np.zeros(95, np.uint32)
    assert L.orc_tmpl_mux(_ptr(txt), _ptr(qoff), _ptr(qlen), qlen.size, _ptr(t), tmpl.size, _ptr(chan), _ptr(count)) == 0
    return chan[:int(count.sum())].copy(), count


def tmpl_demux(lens, out_off, out_size, tmpl, channels, count):
    L = port()
    L.orc_tmpl_demux.restype = C.c_int
    L.orc_tmpl_demux.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    lens = np.ascontiguousarray(lens, np.uint32); ooff = np.ascontiguousarray(out_off, np.uint64); tmpl = np.ascontiguousarray(tmpl, np.uint8)
    t = tmpl if tmpl.size else np.zeros(1, np.uint8)
    ch = np.ascontiguousarray(channels, np.uint8); ch = ch if ch.size else np.zeros(1, np.uint8); count = np.ascontiguousarray(count, np.uint32)
    out = np.zeros(out_size + 8, np.uint8)
    rc = L.orc_tmpl_demux(_ptr(lens), _ptr(ooff), lens.size, _ptr(t), tmpl.size, _ptr(ch), _ptr(count), _ptr(out))
    return None if rc != 0 else out[:out_size].copy()


def ref_tmpl_mux(txt, qoff, qlen, tmpl_len):
    """the reference's compiled codec_tmpl.c: segconf_finalize (finds the template) + compress -> (template, channels, count[95]) or None (template not dominant)"""
    L = gz_ref()
    L.ref_tmpl_mux.restype = C.c_int
    L.ref_tmpl_mux.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32)
    tmpl = np.zeros(tmpl_len + 8, np.uint8); chan = np.zeros(int(qlen.sum()) + 8, np.uint8); count = np.zeros(95, np.uint32)
    rc = L.ref_tmpl_mux(_ptr(txt), txt.size, _ptr(qoff), _ptr(qlen), qlen.size, tmpl_len, _ptr(tmpl), _ptr(chan), _ptr(count))
    assert rc >= 0, rc
    return None if rc == 1 else (tmpl[:tmpl_len].copy(), chan[:int(count.sum())].copy(), count)


# ---------------------------------------------------------------- PACB (src/codec_pacb.c)
def pacb_mux(txt, qoff, qlen, soff, np0, max_np, lib="port"):
    """codec_pacb_compress -> (the 7 * max_np channels back to back, count[84])"""
    txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32); soff = np.ascontiguousarray(soff, np.uint64)
    n0 = None if np0 is None else np.ascontiguousarray(np0, np.uint8)
    chan = np.zeros(int(qlen.sum()) + 8, np.uint8); count = np.zeros(84, np.uint32); n0p = None if n0 is None else _ptr(n0)
    if lib == "port":
        L = port()
        L.orc_pacb_mux.restype = C.c_int
        L.orc_pacb_mux.argtypes = [C.c_void_p] * 5 + [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        rc = L.orc_pacb_mux(_ptr(txt), _ptr(qoff), _ptr(qlen), _ptr(soff), n0p, max_np, qlen.size, _ptr(chan), _ptr(count))
    else:
        L = gz_ref()
        L.ref_pacb_mux.restype = C.c_int
        L.ref_pacb_mux.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 4 + [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        rc = L.ref_pacb_mux(_ptr(txt), txt.size, _ptr(qoff), _ptr(qlen), _ptr(soff), n0p, max_np, qlen.size, _ptr(chan), _ptr(count))
    assert rc == 0, rc
    return chan[:int(count.sum())].copy(), count


def pacb_demux(txt, soff, lens, np0, max_np, out_off, out_size, channels, count, lib="port"):
    txt = np.ascontiguousarray(txt, np.uint8); soff = np.ascontiguousarray(soff, np.uint64); lens = np.ascontiguousarray(lens, np.uint32); ooff = np.ascontiguousarray(out_off, np.uint64)
    n0 = None if np0 is None else np.ascontiguousarray(np0, np.uint8); n0p = None if n0 is None else _ptr(n0)
    ch = np.ascontiguousarray(channels, np.uint8); ch = ch if ch.size else np.zeros(1, np.uint8); count = np.ascontiguousarray(count, np.uint32)
    out = np.zeros(out_size + 8, np.uint8)
    if lib == "port":
        L = port()
        L.orc_pacb_demux.restype = C.c_int
        L.orc_pacb_demux.argtypes = [C.c_void_p] * 4 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = L.orc_pacb_demux(_ptr(txt), _ptr(soff), _ptr(lens), n0p, max_np, _ptr(ooff), lens.size, _ptr(ch), _ptr(count), _ptr(out))
    else:
        L = gz_ref()
        L.ref_pacb_demux.restype = C.c_int
        L.ref_pacb_demux.argtypes = [C.c_void_p, C.c_uint64] + [C.c_void_p] * 3 + [C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        rc = L.ref_pacb_demux(_ptr(txt), txt.size, _ptr(soff), _ptr(lens), n0p, max_np, _ptr(ooff), lens.size, _ptr(ch), _ptr(count), _ptr(out), out_size)
    return None if rc != 0 else out[:out_size].copy()
