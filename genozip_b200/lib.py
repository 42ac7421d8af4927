"""ctypes binding of include/gzb200.h.

Mirrors the reference's plug-in vocabulary for this path: codec names are the reference's `Codec` enum names
(src/genozip.h:322-360: RANB, RANW, RANb, RANw, ARTB, ARTW, ARTb, ARTw) and `est_size` is codec_*_est_size
(src/codec_htscodecs.c:26-33).  Error behaviour follows the reference: a too-small output buffer is the only
recoverable condition (status GZB_SOFT_FAIL, reference `return false` under soft_fail, src/compressor.c:90);
everything else raises.
"""
import ctypes as C
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.path.join(HERE, "libgzb200.so")

CODEC = {"NONE": 1, "RANB": 6, "RANW": 7, "RANb": 8, "RANw": 9, "ACGT": 10, "XCGT": 11, "DOMQ": 13, "PBWT": 15,
         "ARTB": 16, "ARTW": 17, "ARTb": 18, "ARTw": 19, "LONGR": 26}
GZB_DEVICE_PTRS = 1
GZB_OK, GZB_SOFT_FAIL = 0, 1


class GzbError(RuntimeError):
    pass


class Section(C.Structure):
    _fields_ = [("codec", C.c_int32), ("status", C.c_int32), ("in_", C.c_void_p), ("out", C.c_void_p),
                ("in_len", C.c_uint32), ("out_cap", C.c_uint32), ("out_len", C.c_uint32), ("reserved", C.c_uint32)]


_lib = None


def load():
    """Load libgzb200.so; there is no fallback — a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise GzbError(f"{LIBPATH} not built: run `python genozip_b200/build.py` (nvcc, sm_100a). No CPU fallback exists.")
    L = C.CDLL(LIBPATH)
    L.gzb_device_count.restype = C.c_int
    L.gzb_engine_create.restype = C.c_int
    L.gzb_engine_create.argtypes = [C.c_int, C.POINTER(C.c_void_p)]
    L.gzb_engine_destroy.argtypes = [C.c_void_p]
    L.gzb_last_error.restype = C.c_char_p
    L.gzb_last_error.argtypes = [C.c_void_p]
    L.gzb_engine_stream.restype = C.c_void_p
    L.gzb_engine_stream.argtypes = [C.c_void_p]
    L.gzb_engine_sync.argtypes = [C.c_void_p]
    L.gzb_vb_device.restype = C.c_int
    L.gzb_vb_device.argtypes = [C.c_uint32, C.c_int]
    L.gzb_kernel_launches.restype = C.c_uint64
    L.gzb_kernel_launches.argtypes = [C.c_void_p]
    L.gzb_last_chain_ms.restype = C.c_float
    L.gzb_last_chain_ms.argtypes = [C.c_void_p]
    L.gzb_est_size.restype = C.c_uint32
    L.gzb_est_size.argtypes = [C.c_int, C.c_uint64]
    for nm in ("gzb_compress_sections", "gzb_uncompress_sections"):
        getattr(L, nm).restype = C.c_int
        getattr(L, nm).argtypes = [C.c_void_p, C.POINTER(Section), C.c_uint32, C.c_uint32]
    _lib = L
    return L


def est_size(codec, n):
    return load().gzb_est_size(CODEC[codec] if isinstance(codec, str) else codec, n)


class Engine:
    """One engine per (process, GPU): a CUDA stream plus a reusable device workspace."""

    def __init__(self, device=0):
        L = load()
        h = C.c_void_p()
        rc = L.gzb_engine_create(device, C.byref(h))
        if rc != 0:
            raise GzbError(f"gzb_engine_create({device}) failed ({rc}): {L.gzb_last_error(None).decode()}")
        self.h, self.L, self.device = h, L, device

    def close(self):
        if self.h:
            self.L.gzb_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        return self.L.gzb_last_error(self.h).decode()

    @property
    def launches(self):
        return int(self.L.gzb_kernel_launches(self.h))

    @property
    def last_chain_ms(self):
        return float(self.L.gzb_last_chain_ms(self.h))

    def sync(self):
        self.L.gzb_engine_sync(self.h)

    # ---- simple codecs, host buffers (numpy uint8 arrays) ----
    def compress(self, items):
        """items: list of (codec_name, np.uint8 array) -> list of np.uint8 arrays (compressed section bodies)."""
        n = len(items)
        secs = (Section * n)()
        outs, keep = [], []
        for i, (codec, data) in enumerate(items):
            data = np.ascontiguousarray(data, dtype=np.uint8)
            cap = est_size(codec, data.size)
            out = np.empty(cap, dtype=np.uint8)
            keep.append(data)
            outs.append(out)
            secs[i].codec = CODEC[codec]
            secs[i].in_ = data.ctypes.data if data.size else out.ctypes.data
            secs[i].in_len = data.size
            secs[i].out = out.ctypes.data
            secs[i].out_cap = cap
        rc = self.L.gzb_compress_sections(self.h, secs, n, 0)
        if rc != 0:
            raise GzbError(f"gzb_compress_sections failed ({rc}): {self._err()}")
        res = []
        for i in range(n):
            if secs[i].status != 0:
                raise GzbError(f"section {i} ({items[i][0]}, n={items[i][1].size}): status {secs[i].status}")
            res.append(outs[i][:secs[i].out_len].copy())
        return res

    def uncompress(self, items):
        """items: list of (codec_name, compressed np.uint8 array, uncompressed_len) -> list of np.uint8 arrays."""
        n = len(items)
        secs = (Section * n)()
        outs, keep = [], []
        for i, (codec, comp, ulen) in enumerate(items):
            comp = np.ascontiguousarray(comp, dtype=np.uint8)
            out = np.empty(ulen, dtype=np.uint8)
            keep.append(comp)
            outs.append(out)
            secs[i].codec = CODEC[codec]
            secs[i].in_ = comp.ctypes.data
            secs[i].in_len = comp.size
            secs[i].out = out.ctypes.data
            secs[i].out_cap = ulen
        rc = self.L.gzb_uncompress_sections(self.h, secs, n, 0)
        if rc != 0:
            raise GzbError(f"gzb_uncompress_sections failed ({rc}): {self._err()}")
        return outs

    # ---- raw access for bench.py (device pointers / prebuilt section arrays) ----
    def compress_raw(self, secs, n, flags=0):
        rc = self.L.gzb_compress_sections(self.h, secs, n, flags)
        if rc != 0:
            raise GzbError(f"gzb_compress_sections failed ({rc}): {self._err()}")

    def uncompress_raw(self, secs, n, flags=0):
        rc = self.L.gzb_uncompress_sections(self.h, secs, n, flags)
        if rc != 0:
            raise GzbError(f"gzb_uncompress_sections failed ({rc}): {self._err()}")
