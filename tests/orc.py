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
