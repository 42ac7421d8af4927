"""The product's arithmetic-coder chain logic (genozip_b200/csrc/arith_model.cuh) built for the HOST with a one-lane warp
(tests/host/host_arith.cpp) and checked against the reference's own objects (oracle/_ref) / the restatement:
bodies byte-identical, decode bit-exact, for O0/O1 with and without the RLE models.  No GPU needed; the device-only
pieces (float-reciprocal division, 8-entries-per-lane warp search) are covered by the -m gpu parity tests."""
import ctypes as C, itertools, os, subprocess
import numpy as np
import pytest
import orc
from datagen import stream

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def har(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("har") / "libhost_arith.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(HERE, "host", "host_arith.cpp")], check=True)
    L = C.CDLL(so)
    L.har_encode.restype = C.c_uint32
    L.har_encode.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    L.har_decode.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_int, C.c_void_p, C.c_uint32]
    return L


def _varint_len(n):
    k = 1
    while n >= 128:
        n >>= 7; k += 1
    return k


@pytest.mark.parametrize("kind", ["qual", "skew8", "text", "u32le", "runs", "uniform256", "all256", "const", "zeros_hi", "two"])
def test_host_arith_bodies_match_reference(har, kind):
    impl = "ref" if orc.have_ref() else "port"
    for n in (1, 2, 5, 50, 1000, 30000, 150000):
        data = stream(kind, n, 7)
        for o1, rle in itertools.product((0, 1), (0, 1)):
            order = o1 | (0x40 if rle else 0)
            want = orc.compress(impl, "arith", data, order)
            out = np.zeros(2 * n + 64, np.uint8)
            ln = har.har_encode(data.ctypes.data, n, o1, rle, out.ctypes.data)
            if not (want[0] & 0x20):                                       # not stored raw (X_CAT): flags, varint n, body
                hdr = 1 + _varint_len(n)
                assert (want[0] & 0x43) == order
                assert np.array_equal(want[hdr:], out[:ln]), (kind, n, o1, rle)
            if ln <= n:                                                     # (an expanded body is abandoned early: n + 1)
                dec = np.zeros(n + 8, np.uint8)
                body = np.ascontiguousarray(out[:ln])
                har.har_decode(body.ctypes.data, ln, o1, rle, dec.ctypes.data + 1, n)   # misaligned output on purpose
                assert np.array_equal(dec[1:n + 1], data) and dec[0] == 0 and dec[n + 1] == 0, (kind, n, o1, rle)


@pytest.mark.skipif(not orc.have_ref(), reason="needs the reference objects (oracle/_ref)")
def test_host_arith_truncated_and_corrupt_streams_decode_like_the_reference(har):
    """RC_GetFreq / decodeSymbol error returns and the dry-input rule (c_range_coder.h:111-126, c_simple_model.h:153-161)"""
    R = orc.ref()
    for kind, o1 in itertools.product(("text", "qual", "u32le"), (0, 1)):
        data = stream(kind, 20000, 3); n = data.size
        want = orc.compress("ref", "arith", data, o1)
        hdr = 1 + _varint_len(n)
        for cut in (hdr + 6, hdr + (want.size - hdr) // 2, want.size - 3):
            c = np.ascontiguousarray(want[:cut])
            c2 = c.copy(); c2[hdr + 4 if cut > hdr + 8 else hdr + 2] ^= 0x55
            for cc in (c, c2):
                out = np.zeros(n, np.uint8); ol = C.c_uint32(n)
                assert R.arith_uncompress_to(None, cc.ctypes.data, cc.size, out.ctypes.data, C.byref(ol))
                mine = np.zeros(n, np.uint8)
                body = np.ascontiguousarray(cc[hdr:])
                har.har_decode(body.ctypes.data, body.size, o1, 0, mine.ctypes.data, n)
                assert np.array_equal(out, mine), (kind, o1, cut)
