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
# GZB200_LIB: another BUILD of this same library (an A/B run of a kernel variant, tools/ab_build.py) — never a different implementation
LIBPATH = os.environ.get("GZB200_LIB") or os.path.join(HERE, "libgzb200.so")

CODEC = {"NONE": 1, "RANB": 6, "RANW": 7, "RANb": 8, "RANw": 9, "ACGT": 10, "XCGT": 11, "DOMQ": 13, "PBWT": 15,
         "ARTB": 16, "ARTW": 17, "ARTb": 18, "ARTw": 19, "LONGR": 26}
GZB_DEVICE_PTRS, GZB_OUT_DEVICE, GZB_IN_DEVICE = 1, 2, 4
GZB_SEC_IN_DEVICE, GZB_SEC_OUT_DEVICE = 1, 2
GZB_OK, GZB_SOFT_FAIL = 0, 1


class GzbError(RuntimeError):
    pass


class Section(C.Structure):
    _fields_ = [("codec", C.c_int32), ("status", C.c_int32), ("in_", C.c_void_p), ("out", C.c_void_p),
                ("in_len", C.c_uint32), ("out_cap", C.c_uint32), ("out_len", C.c_uint32), ("sflags", C.c_uint32)]


class DomqVb(C.Structure):          # gzb_domq_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("line_off", C.c_void_p), ("line_len", C.c_void_p),
                ("n_lines", C.c_uint32), ("line_dom", C.c_void_p), ("line_diverse", C.c_void_p),
                ("num_norm_qs", C.c_uint8), ("num_doms", C.c_uint8), ("has_diverse", C.c_uint8), ("pad", C.c_uint8),
                ("denorm", C.c_uint8 * (95 * 95)), ("normalize", C.c_uint8 * (95 * 95)),
                ("qual", C.c_void_p), ("qual_cap", C.c_uint32), ("qual_len", C.c_uint32),
                ("runs", C.c_void_p), ("runs_cap", C.c_uint32), ("runs_len", C.c_uint32),
                ("mplx", C.c_void_p), ("mplx_cap", C.c_uint32), ("mplx_len", C.c_uint32),
                ("divr", C.c_void_p), ("divr_cap", C.c_uint32), ("divr_len", C.c_uint32)]


class DomqPizVb(C.Structure):       # gzb_domq_piz_vb
    _fields_ = [("qual", C.c_void_p), ("qual_len", C.c_uint32), ("runs", C.c_void_p), ("runs_len", C.c_uint32),
                ("mplx", C.c_void_p), ("mplx_len", C.c_uint32), ("divr", C.c_void_p), ("divr_len", C.c_uint32),
                ("denorm", C.c_void_p), ("denorm_len", C.c_uint32), ("num_norm_qs", C.c_uint8),
                ("line_len", C.c_void_p), ("n_lines", C.c_uint32), ("out", C.c_void_p), ("out_cap", C.c_uint64)]


class AcgtVb(C.Structure):         # gzb_acgt_vb
    _fields_ = [("seq", C.c_void_p), ("n_bases", C.c_uint64), ("packed", C.c_void_p), ("x", C.c_void_p),
                ("x_all_zero", C.c_int32), ("reserved", C.c_uint32)]


class Copy(C.Structure):            # gzb_copy
    _fields_ = [("src", C.c_void_p), ("dst", C.c_void_p), ("len", C.c_uint64)]


class NormqVb(C.Structure):         # gzb_normq_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("line_off", C.c_void_p), ("line_len", C.c_void_p), ("is_rev", C.c_void_p),
                ("n_lines", C.c_uint32), ("status", C.c_int32), ("local", C.c_void_p), ("local_cap", C.c_uint64), ("local_len", C.c_uint64),
                ("out", C.c_void_p), ("out_cap", C.c_uint64), ("missing", C.c_void_p)]


class OqVb(C.Structure):            # gzb_oq_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("qual_off", C.c_void_p), ("qual_len", C.c_void_p), ("oq_off", C.c_void_p),
                ("seq_len", C.c_void_p), ("n_lines", C.c_uint32), ("status", C.c_int32), ("key_bias", C.c_uint32), ("reserved", C.c_uint32),
                ("channels", C.c_void_p), ("channels_cap", C.c_uint64), ("count", C.c_uint32 * 94), ("monochars", C.c_uint8 * 94), ("pad", C.c_uint8 * 2),
                ("out", C.c_void_p), ("out_cap", C.c_uint64), ("out_off", C.c_void_p)]


class TransposeItem(C.Structure):   # gzb_transpose_item
    _fields_ = [("data", C.c_void_p), ("n_elems", C.c_uint64), ("cols", C.c_uint32), ("width", C.c_uint8), ("dir", C.c_uint8),
                ("transposed", C.c_uint8), ("pad", C.c_uint8), ("status", C.c_int32), ("reserved", C.c_int32)]


class B250Item(C.Structure):        # gzb_b250_item
    _fields_ = [("b250", C.c_void_p), ("len", C.c_uint64), ("out", C.c_void_p), ("ni2wi", C.c_void_p), ("n_new", C.c_uint32), ("ol_len", C.c_uint32),
                ("one_up_ok", C.c_uint8), ("pad", C.c_uint8 * 3), ("status", C.c_int32), ("out_len", C.c_uint64), ("n_words", C.c_uint64)]


class HompVb(C.Structure):          # gzb_homp_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("str_off", C.c_void_p), ("str_len", C.c_void_p), ("seq_off", C.c_void_p),
                ("n_lines", C.c_uint32), ("status", C.c_int32), ("local", C.c_void_p), ("local_cap", C.c_uint64), ("local_len", C.c_uint64),
                ("new_len", C.c_void_p), ("out", C.c_void_p), ("out_cap", C.c_uint64), ("missing", C.c_void_p)]


class SmuxVb(C.Structure):          # gzb_smux_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("qual_off", C.c_void_p), ("qual_len", C.c_void_p), ("seq_off", C.c_void_p),
                ("seq_len", C.c_void_p), ("is_rev", C.c_void_p), ("n_lines", C.c_uint32), ("status", C.c_int32),
                ("channels", C.c_void_p), ("channels_cap", C.c_uint64), ("count", C.c_uint32 * 5), ("n_param", C.c_uint8), ("pad", C.c_uint8 * 3),
                ("out", C.c_void_p), ("out_cap", C.c_uint64), ("out_off", C.c_void_p)]


class TmplVb(C.Structure):          # gzb_tmpl_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("qual_off", C.c_void_p), ("qual_len", C.c_void_p), ("n_lines", C.c_uint32), ("status", C.c_int32),
                ("tmpl", C.c_void_p), ("tmpl_len", C.c_uint32), ("reserved", C.c_uint32), ("channels", C.c_void_p), ("channels_cap", C.c_uint64),
                ("count", C.c_uint32 * 95), ("pad", C.c_uint32), ("out", C.c_void_p), ("out_cap", C.c_uint64), ("out_off", C.c_void_p)]


class PacbVb(C.Structure):          # gzb_pacb_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("qual_off", C.c_void_p), ("qual_len", C.c_void_p), ("seq_off", C.c_void_p), ("np0", C.c_void_p),
                ("n_lines", C.c_uint32), ("status", C.c_int32), ("max_np", C.c_uint32), ("reserved", C.c_uint32), ("channels", C.c_void_p), ("channels_cap", C.c_uint64),
                ("count", C.c_uint32 * 84), ("out", C.c_void_p), ("out_cap", C.c_uint64), ("out_off", C.c_void_p)]


class LocalItem(C.Structure):       # gzb_local_item
    _fields_ = [("data", C.c_void_p), ("n_elems", C.c_uint64), ("op", C.c_int32), ("status", C.c_int32)]


LT_OPS = {"swap16": 1, "swap32": 2, "swap64": 3, "interlace8": 4, "interlace16": 5, "interlace32": 6, "interlace64": 7,
          "deinterlace8": 8, "deinterlace16": 9, "deinterlace32": 10, "deinterlace64": 11}
LT_WIDTH = {1: 2, 2: 4, 3: 8, 4: 1, 5: 2, 6: 4, 7: 8, 8: 1, 9: 2, 10: 4, 11: 8}


class DigestItem(C.Structure):      # gzb_digest_item
    _fields_ = [("data", C.c_void_p), ("len", C.c_uint64), ("adler", C.c_uint32), ("reserved", C.c_uint32)]


class AssignItem(C.Structure):      # gzb_assign_item
    _fields_ = [("data", C.c_void_p), ("len", C.c_uint64), ("sample_len", C.c_uint32), ("size", C.c_uint32 * 8), ("best", C.c_int32)]


class LongrVb(C.Structure):         # gzb_longr_vb
    _fields_ = [("txt", C.c_void_p), ("txt_len", C.c_uint64), ("seq_off", C.c_void_p), ("qual_off", C.c_void_p),
                ("len", C.c_void_p), ("is_rev", C.c_void_p), ("n_lines", C.c_uint32), ("value_to_bin", C.c_uint8 * 256),
                ("values", C.c_void_p), ("lens_be", C.c_void_p), ("qual_out", C.c_void_p),
                ("missing", C.c_void_p), ("qual_len", C.c_void_p), ("n_bases", C.c_uint64)]


class PbwtVb(C.Structure):          # gzb_pbwt_vb
    _fields_ = [("ht", C.c_void_p), ("ht_cap", C.c_uint64), ("ht_len", C.c_uint64), ("n_lines", C.c_uint32), ("ht_per_line", C.c_uint32),
                ("runs", C.c_void_p), ("runs_cap", C.c_uint32), ("n_runs", C.c_uint32),
                ("fgrc", C.c_void_p), ("fgrc_cap", C.c_uint32), ("n_fgrc", C.c_uint32), ("status", C.c_int32), ("reserved", C.c_uint32)]


_lib = None
_TESTS_MAY_LOAD_EMULATION = False      # set by tests/simt_lib.py and tests/conftest.py (--simt) only


def load():
    """Load libgzb200.so; there is no fallback — a missing library is an error."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIBPATH):
        raise GzbError(f"{LIBPATH} not built: run `python genozip_b200/build.py` (nvcc, sm_100a). No CPU fallback exists.")
    L = C.CDLL(LIBPATH)
    L.gzb_build_is_emulation.restype = C.c_int
    if L.gzb_build_is_emulation() and not _TESTS_MAY_LOAD_EMULATION:
        raise GzbError(f"{LIBPATH} is the test suite's host build of the kernels (tests/host/simt), not the CUDA library: the product never loads it.")
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
    L.gzb_last_kernel_ms.restype = C.c_float
    L.gzb_last_kernel_ms.argtypes = [C.c_void_p, C.c_int]
    L.gzb_est_size.restype = C.c_uint32
    L.gzb_est_size.argtypes = [C.c_int, C.c_uint64]
    for nm in ("gzb_compress_sections", "gzb_uncompress_sections"):
        getattr(L, nm).restype = C.c_int
        getattr(L, nm).argtypes = [C.c_void_p, C.POINTER(Section), C.c_uint32, C.c_uint32]
    L.gzb_acgt_packed_len.restype = C.c_uint64
    L.gzb_acgt_packed_len.argtypes = [C.c_uint64]
    L.gzb_acgt_pack.restype = C.c_int
    L.gzb_acgt_pack.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_uint32]
    L.gzb_acgt_unpack.restype = C.c_int
    L.gzb_acgt_unpack.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    for nm in ("gzb_acgt_pack_batch", "gzb_acgt_unpack_batch"):
        getattr(L, nm).restype = C.c_int
        getattr(L, nm).argtypes = [C.c_void_p, C.POINTER(AcgtVb), C.c_uint32, C.c_uint32]
    for nm in ("gzb_domq_prepare", "gzb_domq_split"):
        getattr(L, nm).restype = C.c_int
        getattr(L, nm).argtypes = [C.c_void_p, C.POINTER(DomqVb), C.c_uint32, C.c_uint32]
    L.gzb_pbwt_encode.restype = C.c_int
    L.gzb_pbwt_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32),
                                  C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_uint32]
    for f in ("gzb_pbwt_encode_batch", "gzb_pbwt_decode_batch"):
        getattr(L, f).restype = C.c_int
        getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.gzb_longr_calculate_bins.restype = C.c_int
    L.gzb_longr_calculate_bins.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
    L.gzb_compress_sections_packed.restype = C.c_int
    L.gzb_compress_sections_packed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32]
    L.gzb_copy_batch.restype = C.c_int
    L.gzb_copy_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    for f in ("gzb_stage_upload", "gzb_stage_fetch"):
        getattr(L, f).restype = C.c_int; getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    L.gzb_stage_wait.restype = C.c_int; L.gzb_stage_wait.argtypes = [C.c_void_p, C.c_int]
    for f in ("gzb_normq_gather", "gzb_normq_reconstruct", "gzb_oq_mux", "gzb_oq_demux", "gzb_smux_mux", "gzb_smux_demux", "gzb_tmpl_mux", "gzb_tmpl_demux", "gzb_pacb_mux", "gzb_pacb_demux"):
        getattr(L, f).restype = C.c_int; getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.gzb_local_transform_batch.restype = C.c_int
    L.gzb_local_transform_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.gzb_adler32_batch.restype = C.c_int
    L.gzb_adler32_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    for f in ("gzb_homp_condense", "gzb_homp_expand"):
        getattr(L, f).restype = C.c_int; getattr(L, f).argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32]
    L.gzb_b250_generate_batch.restype = C.c_int
    L.gzb_b250_generate_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.gzb_local_transpose_batch.restype = C.c_int
    L.gzb_local_transpose_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.gzb_assign_codecs.restype = C.c_int
    L.gzb_assign_codecs.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32]
    L.gzb_pbwt_decode.restype = C.c_int
    L.gzb_pbwt_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint64,
                                  C.POINTER(C.c_uint64), C.c_uint32]
    for nm in ("gzb_longr_encode", "gzb_longr_decode"):
        getattr(L, nm).restype = C.c_int
        getattr(L, nm).argtypes = [C.c_void_p, C.POINTER(LongrVb), C.c_uint32, C.c_uint32]
    L.gzb_domq_reconstruct.restype = C.c_int
    L.gzb_domq_reconstruct.argtypes = [C.c_void_p, C.POINTER(DomqPizVb), C.c_uint32, C.c_uint32]
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

    def trim(self):
        """give the engine's grow-only workspace back (it is re-allocated on demand)"""
        if hasattr(self.L, "gzb_engine_trim"):
            self.L.gzb_engine_trim(self.h)

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

    def compress_packed(self, items, arena_cap=None):
        """gzb_compress_sections_packed on host buffers: the sections appended to ONE buffer (zfile_compress_local_data's z_data
        append) -> (list of compressed bodies, bytes used).  A buffer that is too small is grown and the call repeated, like the
        soft-fail retry of compressor.c:90-110."""
        n = len(items)
        secs = (Section * n)()
        keep = [np.ascontiguousarray(d, dtype=np.uint8) for _, d in items]
        dummy = np.zeros(16, np.uint8)
        for i, ((codec, _), data) in enumerate(zip(items, keep)):
            secs[i].codec = CODEC[codec]; secs[i].in_ = data.ctypes.data if data.size else dummy.ctypes.data; secs[i].in_len = data.size
        cap = arena_cap if arena_cap is not None else max(4096, sum(d.size for d in keep) // 2)
        while True:
            arena = np.empty(max(cap, 16), np.uint8); used = C.c_uint64()
            rc = self.L.gzb_compress_sections_packed(self.h, secs, n, arena.ctypes.data, cap, C.byref(used), 0)
            if rc == 1 and used.value > cap:
                cap = used.value
                continue
            if rc != 0:
                raise GzbError(f"gzb_compress_sections_packed failed ({rc}): {self._err()}")
            base = arena.ctypes.data
            return [arena[secs[i].out - base: secs[i].out - base + secs[i].out_len].copy() for i in range(n)], used.value

    # ---- raw access for bench.py (device pointers / prebuilt section arrays) ----
    def compress_raw(self, secs, n, flags=0):
        rc = self.L.gzb_compress_sections(self.h, secs, n, flags)
        if rc != 0:
            raise GzbError(f"gzb_compress_sections failed ({rc}): {self._err()}")

    def uncompress_raw(self, secs, n, flags=0):
        rc = self.L.gzb_uncompress_sections(self.h, secs, n, flags)
        if rc != 0:
            raise GzbError(f"gzb_uncompress_sections failed ({rc}): {self._err()}")

    # ---- zip_generate_local's transforms (host buffers, in place on copies) ----
    def local_transform(self, items):
        """items: list of (op name, numpy array of the operation's width) -> list of transformed arrays (zip.c:167-213 / buffer.c:337-353, 431-468)"""
        arr = (LocalItem * max(1, len(items)))(); keep = []
        for i, (op, a) in enumerate(items):
            code = LT_OPS[op]
            b = np.ascontiguousarray(a).copy()
            assert b.dtype.itemsize == LT_WIDTH[code], (op, b.dtype)
            keep.append(b)
            arr[i].data = b.ctypes.data if b.size else None; arr[i].n_elems = b.size; arr[i].op = code
        rc = self.L.gzb_local_transform_batch(self.h, arr, len(items), 0)
        if rc != 0:
            raise GzbError(f"gzb_local_transform_batch failed ({rc}): {self._err()}")
        return keep

    # ---- NORMQ (host buffers) ----
    def normq_gather(self, vbs):
        """vbs: list of (txt, line_off, line_len, is_rev or None) -> list of QUAL.local arrays (codec_normq_compress before its sub-codec)"""
        arr = (NormqVb * max(1, len(vbs)))(); keep = []
        for i, (txt, off, ln, rev) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); off = np.ascontiguousarray(off, np.uint64); ln = np.ascontiguousarray(ln, np.uint32)
            rv = None if rev is None else np.ascontiguousarray(rev, np.uint8)
            out = np.zeros(int(ln.sum()) + 16, np.uint8)
            keep.append((txt, off, ln, rv, out))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.line_off = off.ctypes.data if off.size else None
            a.line_len = ln.ctypes.data if ln.size else None; a.is_rev = None if rv is None or not rv.size else rv.ctypes.data; a.n_lines = ln.size
            a.local = out.ctypes.data; a.local_cap = out.size
        rc = self.L.gzb_normq_gather(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_normq_gather failed ({rc}): {self._err()}")
        return [k[4][:int(arr[i].local_len)].copy() for i, k in enumerate(keep)]

    def normq_reconstruct(self, vbs):
        """vbs: list of (local, line_len, is_rev or None) -> list of (out with line_len[i] bytes per line, missing flags); raises on a stream that does not fit"""
        arr = (NormqVb * max(1, len(vbs)))(); keep = []
        for i, (local, ln, rev) in enumerate(vbs):
            local = np.ascontiguousarray(local, np.uint8); ln = np.ascontiguousarray(ln, np.uint32)
            rv = None if rev is None else np.ascontiguousarray(rev, np.uint8)
            out = np.zeros(int(ln.sum()) + 16, np.uint8); miss = np.zeros(ln.size + 1, np.uint8)
            keep.append((local, ln, rv, out, miss))
            a = arr[i]
            a.local = local.ctypes.data if local.size else None; a.local_len = local.size; a.line_len = ln.ctypes.data if ln.size else None
            a.is_rev = None if rv is None or not rv.size else rv.ctypes.data; a.n_lines = ln.size
            a.out = out.ctypes.data; a.out_cap = out.size; a.missing = miss.ctypes.data
        rc = self.L.gzb_normq_reconstruct(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_normq_reconstruct failed ({rc}): {self._err()}")
        return [(k[3][:int(k[1].sum())].copy(), k[4][:k[1].size].copy()) for k in keep]

    # ---- Adler-32 (host buffers; device pointers through adler32_ptrs) ----
    def adler32(self, bufs):
        """adler32 (1, data, len) of every buffer (compressor.c:151,161 z_digest) -> list of ints"""
        arrs = [np.ascontiguousarray(b, dtype=np.uint8) for b in bufs]
        return self.adler32_ptrs([(a.ctypes.data if a.size else 0, a.size) for a in arrs], 0)

    def adler32_ptrs(self, ptr_len, flags):
        items = (DigestItem * max(1, len(ptr_len)))()
        for i, (p, n) in enumerate(ptr_len):
            items[i].data = p; items[i].len = n
        rc = self.L.gzb_adler32_batch(self.h, items, len(ptr_len), flags)
        if rc != 0:
            raise GzbError(f"gzb_adler32_batch failed ({rc}): {self._err()}")
        return [int(items[i].adler) for i in range(len(ptr_len))]

    # ---- SMUX (host buffers) ----
    def smux_mux(self, vbs):
        """vbs: list of (txt, qual_off, qual_len, seq_off, seq_len, is_rev or None) -> list of (5 channels back to back, count[5], n_param)"""
        arr = (SmuxVb * max(1, len(vbs)))(); keep = []
        for i, (txt, qoff, qlen, soff, slen, rev) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32)
            soff = np.ascontiguousarray(soff, np.uint64); slen = np.ascontiguousarray(slen, np.uint32); rv = None if rev is None else np.ascontiguousarray(rev, np.uint8)
            ch = np.zeros(int(qlen.sum()) + 16, np.uint8)
            keep.append((txt, qoff, qlen, soff, slen, rv, ch))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = qlen.size
            a.qual_off = qoff.ctypes.data if qoff.size else None; a.qual_len = qlen.ctypes.data if qlen.size else None
            a.seq_off = soff.ctypes.data if soff.size else None; a.seq_len = slen.ctypes.data if slen.size else None
            a.is_rev = None if rv is None or not rv.size else rv.ctypes.data
            a.channels = ch.ctypes.data; a.channels_cap = ch.size
        rc = self.L.gzb_smux_mux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_smux_mux failed ({rc}): {self._err()}")
        out = []
        for i, k in enumerate(keep):
            cnt = np.array(arr[i].count[:], np.uint32)
            out.append((k[6][:int(cnt.sum())].copy(), cnt, int(arr[i].n_param)))
        return out

    def smux_demux(self, vbs):
        """vbs: list of (txt, seq_off, lens, is_rev or None, out_off, out_size, channels, count[5], n_param) -> list of out arrays"""
        arr = (SmuxVb * max(1, len(vbs)))(); keep = []
        for i, (txt, soff, lens, rev, ooff, out_size, ch, cnt, par) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); soff = np.ascontiguousarray(soff, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
            ooff = np.ascontiguousarray(ooff, np.uint64); rv = None if rev is None else np.ascontiguousarray(rev, np.uint8)
            ch = np.ascontiguousarray(ch, np.uint8); ch = ch if ch.size else np.zeros(1, np.uint8)
            out = np.zeros(out_size + 16, np.uint8)
            keep.append((txt, soff, lens, ooff, rv, ch, out, out_size))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = lens.size
            a.qual_len = lens.ctypes.data if lens.size else None; a.seq_off = soff.ctypes.data if soff.size else None
            a.is_rev = None if rv is None or not rv.size else rv.ctypes.data; a.out_off = ooff.ctypes.data if ooff.size else None
            a.channels = ch.ctypes.data; a.channels_cap = ch.size; a.n_param = par
            for b in range(5):
                a.count[b] = int(cnt[b])
            a.out = out.ctypes.data; a.out_cap = out_size
        rc = self.L.gzb_smux_demux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_smux_demux failed ({rc}): {self._err()}")
        return [k[6][:k[7]].copy() for k in keep]

    # ---- PACB (host buffers) ----
    def pacb_mux(self, vbs):
        """vbs: list of (txt, qual_off, qual_len, seq_off, np0 or None, max_np) -> list of (channels back to back, count[84])"""
        arr = (PacbVb * max(1, len(vbs)))(); keep = []
        for i, (txt, qoff, qlen, soff, np0, max_np) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32); soff = np.ascontiguousarray(soff, np.uint64)
            n0 = None if np0 is None else np.ascontiguousarray(np0, np.uint8); ch = np.zeros(int(qlen.sum()) + 16, np.uint8)
            keep.append((txt, qoff, qlen, soff, n0, ch))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = qlen.size; a.max_np = max_np
            a.qual_off = qoff.ctypes.data if qoff.size else None; a.qual_len = qlen.ctypes.data if qlen.size else None; a.seq_off = soff.ctypes.data if soff.size else None
            a.np0 = None if n0 is None or not n0.size else n0.ctypes.data
            a.channels = ch.ctypes.data; a.channels_cap = ch.size
        rc = self.L.gzb_pacb_mux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_pacb_mux failed ({rc}): {self._err()}")
        out = []
        for i, k in enumerate(keep):
            cnt = np.array(arr[i].count[:], np.uint32)
            out.append((k[5][:int(cnt.sum())].copy(), cnt))
        return out

    def pacb_demux(self, vbs):
        """vbs: list of (txt, seq_off, lens, np0 or None, max_np, out_off, out_size, channels, count[84]) -> list of out arrays"""
        arr = (PacbVb * max(1, len(vbs)))(); keep = []
        for i, (txt, soff, lens, np0, max_np, ooff, out_size, ch, cnt) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); soff = np.ascontiguousarray(soff, np.uint64); lens = np.ascontiguousarray(lens, np.uint32); ooff = np.ascontiguousarray(ooff, np.uint64)
            n0 = None if np0 is None else np.ascontiguousarray(np0, np.uint8)
            ch = np.ascontiguousarray(ch, np.uint8); ch = ch if ch.size else np.zeros(1, np.uint8); out = np.zeros(out_size + 16, np.uint8)
            keep.append((txt, soff, lens, ooff, n0, ch, out, out_size))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = lens.size; a.max_np = max_np
            a.qual_len = lens.ctypes.data if lens.size else None; a.seq_off = soff.ctypes.data if soff.size else None; a.out_off = ooff.ctypes.data if ooff.size else None
            a.np0 = None if n0 is None or not n0.size else n0.ctypes.data
            a.channels = ch.ctypes.data; a.channels_cap = ch.size
            for c in range(84):
                a.count[c] = int(cnt[c])
            a.out = out.ctypes.data; a.out_cap = out_size
        rc = self.L.gzb_pacb_demux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_pacb_demux failed ({rc}): {self._err()}")
        return [k[6][:k[7]].copy() for k in keep]

    # ---- TMPL (host buffers) ----
    def tmpl_mux(self, vbs):
        """vbs: list of (txt, qual_off, qual_len, template) -> list of (channels 0..93 and the excess back to back, count[95])"""
        arr = (TmplVb * max(1, len(vbs)))(); keep = []
        for i, (txt, qoff, qlen, tmpl) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32); tmpl = np.ascontiguousarray(tmpl, np.uint8)
            ch = np.zeros(int(qlen.sum()) + 16, np.uint8)
            keep.append((txt, qoff, qlen, tmpl, ch))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = qlen.size
            a.qual_off = qoff.ctypes.data if qoff.size else None; a.qual_len = qlen.ctypes.data if qlen.size else None
            a.tmpl = tmpl.ctypes.data if tmpl.size else None; a.tmpl_len = tmpl.size
            a.channels = ch.ctypes.data; a.channels_cap = ch.size
        rc = self.L.gzb_tmpl_mux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_tmpl_mux failed ({rc}): {self._err()}")
        out = []
        for i, k in enumerate(keep):
            cnt = np.array(arr[i].count[:], np.uint32)
            out.append((k[4][:int(cnt.sum())].copy(), cnt))
        return out

    def tmpl_demux(self, vbs):
        """vbs: list of (lens, out_off, out_size, template, channels, count[95]) -> list of out arrays"""
        arr = (TmplVb * max(1, len(vbs)))(); keep = []
        for i, (lens, ooff, out_size, tmpl, ch, cnt) in enumerate(vbs):
            lens = np.ascontiguousarray(lens, np.uint32); ooff = np.ascontiguousarray(ooff, np.uint64); tmpl = np.ascontiguousarray(tmpl, np.uint8)
            ch = np.ascontiguousarray(ch, np.uint8); ch = ch if ch.size else np.zeros(1, np.uint8); out = np.zeros(out_size + 16, np.uint8)
            keep.append((lens, ooff, tmpl, ch, out, out_size))
            a = arr[i]
            a.n_lines = lens.size; a.qual_len = lens.ctypes.data if lens.size else None; a.out_off = ooff.ctypes.data if ooff.size else None
            a.tmpl = tmpl.ctypes.data if tmpl.size else None; a.tmpl_len = tmpl.size
            a.channels = ch.ctypes.data; a.channels_cap = ch.size
            for q in range(95):
                a.count[q] = int(cnt[q])
            a.out = out.ctypes.data; a.out_cap = out_size
        rc = self.L.gzb_tmpl_demux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_tmpl_demux failed ({rc}): {self._err()}")
        return [k[4][:k[5]].copy() for k in keep]

    # ---- HOMP / T0 (host buffers) ----
    def hp_condense(self, mode, vbs):
        """vbs: list of (txt, str_off, str_len, seq_off) -> list of (condensed strings back to back, new lengths): the first pass of
        codec_homp_compress (mode 0) / codec_t0_compress (mode 1)"""
        arr = (HompVb * max(1, len(vbs)))(); keep = []
        for i, (txt, so, sl, qo) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); so = np.ascontiguousarray(so, np.uint64); sl = np.ascontiguousarray(sl, np.uint32); qo = np.ascontiguousarray(qo, np.uint64)
            out = np.zeros(int(sl.sum()) + 16, np.uint8); nl = np.zeros(sl.size + 1, np.uint32)
            keep.append((txt, so, sl, qo, out, nl))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = sl.size
            a.str_off = so.ctypes.data if so.size else None; a.str_len = sl.ctypes.data if sl.size else None; a.seq_off = qo.ctypes.data if qo.size else None
            a.local = out.ctypes.data; a.local_cap = out.size; a.new_len = nl.ctypes.data
        rc = self.L.gzb_homp_condense(self.h, arr, len(vbs), mode, 0)
        if rc != 0:
            raise GzbError(f"gzb_homp_condense failed ({rc}): {self._err()}")
        return [(k[4][:int(arr[i].local_len)].copy(), k[5][:k[2].size].copy()) for i, k in enumerate(keep)]

    def hp_expand(self, mode, vbs):
        """vbs: list of (local, txt, seq_off, lens) -> list of (lens[i] bytes per line, missing flags): codec_homp_reconstruct / codec_t0_reconstruct for every line"""
        arr = (HompVb * max(1, len(vbs)))(); keep = []
        for i, (local, txt, qo, sl) in enumerate(vbs):
            local = np.ascontiguousarray(local, np.uint8); txt = np.ascontiguousarray(txt, np.uint8); qo = np.ascontiguousarray(qo, np.uint64); sl = np.ascontiguousarray(sl, np.uint32)
            out = np.zeros(int(sl.sum()) + 16, np.uint8); miss = np.zeros(sl.size + 1, np.uint8)
            keep.append((local, txt, qo, sl, out, miss))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = sl.size
            a.str_len = sl.ctypes.data if sl.size else None; a.seq_off = qo.ctypes.data if qo.size else None
            a.local = local.ctypes.data if local.size else None; a.local_len = local.size
            a.out = out.ctypes.data; a.out_cap = out.size; a.missing = miss.ctypes.data
        rc = self.L.gzb_homp_expand(self.h, arr, len(vbs), mode, 0)
        if rc != 0:
            raise GzbError(f"gzb_homp_expand failed ({rc}): {self._err()}")
        return [(k[4][:int(k[3].sum())].copy(), k[5][:k[3].size].copy()) for k in keep]

    def b250_generate(self, items):
        """items: list of (b250 bytes as the segmenter left them, ni2wi int32 array, ol_len, one_up_ok) -> list of (converted bytes, n_words):
        b250_zip_generate (b250.c:202-297)"""
        arr = (B250Item * max(1, len(items)))(); keep = []
        for i, (b, t, ol, up) in enumerate(items):
            b = np.ascontiguousarray(b, np.uint8); t = np.ascontiguousarray(t, np.int32); out = np.zeros(b.size + 8, np.uint8)
            keep.append((b, t, out))
            a = arr[i]
            a.b250 = b.ctypes.data if b.size else None; a.len = b.size; a.out = out.ctypes.data; a.ni2wi = t.ctypes.data if t.size else None
            a.n_new = t.size; a.ol_len = ol; a.one_up_ok = 1 if up else 0
        rc = self.L.gzb_b250_generate_batch(self.h, arr, len(items), 0)
        if rc != 0:
            raise GzbError(f"gzb_b250_generate_batch failed ({rc}): {self._err()}")
        return [(k[2][k[0].size - int(arr[i].out_len):k[0].size].copy(), int(arr[i].n_words)) for i, k in enumerate(keep)]

    def local_transpose(self, items, piz=False):
        """items: list of (array of uint8/16/32, cols) -> list of (array, transposed flag): dyn_int_transpose (ZIP) or BGEN_transpose_u*_buf (PIZ)"""
        arr = (TransposeItem * max(1, len(items)))(); keep = []
        for i, (a, cols) in enumerate(items):
            a = np.ascontiguousarray(a).copy()
            keep.append(a)
            arr[i].data = a.ctypes.data if a.size else None; arr[i].n_elems = a.size; arr[i].cols = cols; arr[i].width = a.dtype.itemsize; arr[i].dir = 1 if piz else 0
        rc = self.L.gzb_local_transpose_batch(self.h, arr, len(items), 0)
        if rc != 0:
            raise GzbError(f"gzb_local_transpose_batch failed ({rc}): {self._err()}")
        return [(a, bool(arr[i].transposed)) for i, a in enumerate(keep)]

    # ---- OQ (host buffers) ----
    def oq_mux(self, vbs):
        """vbs: list of (txt, qual_off, qual_len, oq_off, seq_len or None) -> list of (channels back to back, count[94], monochars[94])
        (codec_oq_compress before its sub-codec)"""
        arr = (OqVb * max(1, len(vbs)))(); keep = []
        for i, (txt, qoff, qlen, ooff, sl) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32)
            ooff = np.ascontiguousarray(ooff, np.uint64); sl = None if sl is None else np.ascontiguousarray(sl, np.uint32)
            ch = np.zeros(int(qlen.sum()) + 16, np.uint8)
            keep.append((txt, qoff, qlen, ooff, sl, ch))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = qlen.size
            a.qual_off = qoff.ctypes.data if qoff.size else None; a.qual_len = qlen.ctypes.data if qlen.size else None
            a.oq_off = ooff.ctypes.data if ooff.size else None; a.seq_len = None if sl is None or not sl.size else sl.ctypes.data
            a.channels = ch.ctypes.data; a.channels_cap = ch.size
        rc = self.L.gzb_oq_mux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_oq_mux failed ({rc}): {self._err()}")
        out = []
        for i, k in enumerate(keep):
            cnt = np.array(arr[i].count[:], np.uint32)
            out.append((k[5][:int(cnt.sum())].copy(), cnt, np.array(arr[i].monochars[:], np.uint8)))
        return out

    def oq_demux(self, vbs):
        """vbs: list of (txt, qual_off, qual_len, out_off, out_size, key_bias, channels, count[94], monochars[94]) -> list of out arrays
        (codec_oq_reconstruct for every line); raises when a channel is out of data"""
        arr = (OqVb * max(1, len(vbs)))(); keep = []
        for i, (txt, qoff, qlen, ooff, out_size, bias, ch, cnt, mono) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); qoff = np.ascontiguousarray(qoff, np.uint64); qlen = np.ascontiguousarray(qlen, np.uint32)
            ooff = np.ascontiguousarray(ooff, np.uint64); ch = np.ascontiguousarray(ch, np.uint8)
            ch = ch if ch.size else np.zeros(1, np.uint8)
            out = np.zeros(out_size + 16, np.uint8)
            keep.append((txt, qoff, qlen, ooff, ch, out, out_size))
            a = arr[i]
            a.txt = txt.ctypes.data if txt.size else None; a.txt_len = txt.size; a.n_lines = qlen.size; a.key_bias = bias
            a.qual_off = qoff.ctypes.data if qoff.size else None; a.qual_len = qlen.ctypes.data if qlen.size else None
            a.out_off = ooff.ctypes.data if ooff.size else None
            a.channels = ch.ctypes.data; a.channels_cap = ch.size
            for q in range(94):
                a.count[q] = int(cnt[q]); a.monochars[q] = int(mono[q])
            a.out = out.ctypes.data; a.out_cap = out_size
        rc = self.L.gzb_oq_demux(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_oq_demux failed ({rc}): {self._err()}")
        return [k[5][:k[6]].copy() for k in keep]

    def assign_codecs(self, bufs):
        """codec_assign_best_codec's size criterion (codec.c:234-389) for host buffers -> [(best codec name or None, {codec name: body bytes})]"""
        arrs = [np.ascontiguousarray(b, dtype=np.uint8) for b in bufs]
        return self.assign_codecs_ptrs([(a.ctypes.data if a.size else 0, a.size) for a in arrs], 0)

    def assign_codecs_ptrs(self, ptr_len, flags):
        items = (AssignItem * max(1, len(ptr_len)))()
        for i, (p, n) in enumerate(ptr_len):
            items[i].data = p; items[i].len = n
        rc = self.L.gzb_assign_codecs(self.h, items, len(ptr_len), flags)
        if rc != 0:
            raise GzbError(f"gzb_assign_codecs failed ({rc}): {self._err()}")
        names = ("RANB", "RANW", "RANb", "RANw", "ARTB", "ARTW", "ARTb", "ARTw")
        by_id = {v: k for k, v in CODEC.items()}; by_id[1] = "NONE"; by_id[0] = None
        return [(by_id[int(items[i].best)], {nm: int(items[i].size[k]) for k, nm in enumerate(names)} if items[i].best else {}) for i in range(len(ptr_len))]

    # ---- ACGT (host buffers) ----
    def acgt_pack(self, seq):
        """codec_acgt_compress before its sub-codec: -> (packed LE 2-bit words, exception stream x, x_all_zero)"""
        seq = np.ascontiguousarray(seq, dtype=np.uint8)
        n = seq.size
        packed = np.zeros(int(self.L.gzb_acgt_packed_len(n)) or 1, np.uint8)
        x = np.zeros(max(n, 1), np.uint8)
        allz = C.c_int(0)
        rc = self.L.gzb_acgt_pack(self.h, seq.ctypes.data if n else x.ctypes.data, n, packed.ctypes.data, x.ctypes.data, C.byref(allz), 0)
        if rc != 0:
            raise GzbError(f"gzb_acgt_pack failed ({rc}): {self._err()}")
        return packed[:int(self.L.gzb_acgt_packed_len(n))], x[:n], bool(allz.value)

    def acgt_unpack(self, packed, x, n):
        packed = np.ascontiguousarray(packed, dtype=np.uint8)
        out = np.zeros(max(n, 1), np.uint8)
        xp = None if x is None else np.ascontiguousarray(x, dtype=np.uint8).ctypes.data
        rc = self.L.gzb_acgt_unpack(self.h, packed.ctypes.data if packed.size else out.ctypes.data, xp, n, out.ctypes.data, 0)
        if rc != 0:
            raise GzbError(f"gzb_acgt_unpack failed ({rc}): {self._err()}")
        return out[:n]

    def acgt_pack_batch(self, seqs):
        """gzb_acgt_pack_batch on host buffers: list of uint8 arrays -> list of (packed, x, x_all_zero)"""
        n = len(seqs)
        arr = (AcgtVb * n)()
        keep = []
        for i, seq in enumerate(seqs):
            seq = np.ascontiguousarray(seq, dtype=np.uint8)
            packed = np.zeros(int(self.L.gzb_acgt_packed_len(seq.size)) or 1, np.uint8)
            x = np.full(max(seq.size, 1), 0xAA, np.uint8)
            keep.append((seq, packed, x))
            arr[i].seq = seq.ctypes.data if seq.size else x.ctypes.data; arr[i].n_bases = seq.size
            arr[i].packed = packed.ctypes.data; arr[i].x = x.ctypes.data
        rc = self.L.gzb_acgt_pack_batch(self.h, arr, n, 0)
        if rc != 0:
            raise GzbError(f"gzb_acgt_pack_batch failed ({rc}): {self._err()}")
        return [(p[:int(self.L.gzb_acgt_packed_len(s.size))], x[:s.size], bool(arr[i].x_all_zero)) for i, (s, p, x) in enumerate(keep)]

    def acgt_unpack_batch(self, items):
        """gzb_acgt_unpack_batch on host buffers: list of (packed, x or None, n) -> list of uint8 arrays"""
        n = len(items)
        arr = (AcgtVb * n)()
        keep = []
        for i, (packed, x, nb) in enumerate(items):
            packed = np.ascontiguousarray(packed, dtype=np.uint8)
            x = None if x is None else np.ascontiguousarray(x, dtype=np.uint8)
            out = np.zeros(max(nb, 1), np.uint8)
            keep.append((packed, x, out))
            arr[i].seq = out.ctypes.data; arr[i].n_bases = nb
            arr[i].packed = packed.ctypes.data if packed.size else out.ctypes.data
            arr[i].x = None if x is None else x.ctypes.data
        rc = self.L.gzb_acgt_unpack_batch(self.h, arr, n, 0)
        if rc != 0:
            raise GzbError(f"gzb_acgt_unpack_batch failed ({rc}): {self._err()}")
        return [o[:items[i][2]] for i, (_, _, o) in enumerate(keep)]

    # ---- DOMQ (host buffers): batch of VBlocks, each (txt, line_off, line_len) ----
    def domq_encode(self, vbs):
        """codec_domq_prepare_normalize + codec_domq_compress (before the sub-codec) for a batch of VBlocks.
        vbs: list of (txt u8, line_off u64, line_len u32).  Returns a list of dicts with the four streams,
        the per-line dom/diverse and the tables."""
        n = len(vbs)
        arr = (DomqVb * n)()
        keep = []
        for i, (txt, off, ln) in enumerate(vbs):
            txt = np.ascontiguousarray(txt, np.uint8); off = np.ascontiguousarray(off, np.uint64); ln = np.ascontiguousarray(ln, np.uint32)
            tot = int(txt.size)
            bufs = dict(txt=txt if txt.size else np.zeros(1, np.uint8), off=off, ln=ln,
                        dom=np.zeros(max(ln.size, 1), np.uint8), div=np.zeros(max(ln.size, 1), np.uint8),
                        qual=np.zeros(2 * tot + 2, np.uint8), runs=np.zeros(tot + 2, np.uint8),
                        mplx=np.zeros(ln.size + 1, np.uint8), divr=np.zeros(tot + 1, np.uint8))
            keep.append(bufs)
            a = arr[i]
            a.txt = bufs["txt"].ctypes.data; a.txt_len = txt.size
            a.line_off = off.ctypes.data; a.line_len = ln.ctypes.data; a.n_lines = ln.size
            a.line_dom = bufs["dom"].ctypes.data; a.line_diverse = bufs["div"].ctypes.data
            for k in ("qual", "runs", "mplx", "divr"):
                setattr(a, k, bufs[k].ctypes.data); setattr(a, k + "_cap", bufs[k].size)
        rc = self.L.gzb_domq_prepare(self.h, arr, n, 0)
        if rc != 0:
            raise GzbError(f"gzb_domq_prepare failed ({rc}): {self._err()}")
        rc = self.L.gzb_domq_split(self.h, arr, n, 0)
        if rc != 0:
            raise GzbError(f"gzb_domq_split failed ({rc}): {self._err()}")
        res = []
        for i in range(n):
            a, b = arr[i], keep[i]
            nn, nd = a.num_norm_qs, a.num_doms
            res.append(dict(num_norm_qs=nn, num_doms=nd, has_diverse=a.has_diverse,
                            denorm=np.frombuffer(bytes(a.denorm), np.uint8)[:nn * nd].copy(),
                            normalize=np.frombuffer(bytes(a.normalize), np.uint8).copy(),
                            line_dom=b["dom"][:a.n_lines].copy(), line_diverse=b["div"][:a.n_lines].copy(),
                            qual=b["qual"][:a.qual_len].copy(), runs=b["runs"][:a.runs_len].copy(),
                            mplx=b["mplx"][:a.mplx_len].copy(), divr=b["divr"][:a.divr_len].copy()))
        return res

    def domq_decode(self, encs, line_lens):
        """codec_domq_reconstruct for all lines of each VBlock.  encs: dicts as returned by domq_encode."""
        n = len(encs)
        arr = (DomqPizVb * n)()
        keep, outs = [], []
        z = np.zeros(1, np.uint8)
        for i, (enc, ln) in enumerate(zip(encs, line_lens)):
            ln = np.ascontiguousarray(ln, np.uint32)
            tot = int(ln.sum())
            out = np.zeros(tot + 1, np.uint8)
            b = {k: np.ascontiguousarray(enc[k], np.uint8) for k in ("qual", "runs", "mplx", "divr", "denorm")}
            keep.append((b, ln)); outs.append(out)
            a = arr[i]
            for k in ("qual", "runs", "mplx", "divr"):
                setattr(a, k, (b[k] if b[k].size else z).ctypes.data); setattr(a, k + "_len", b[k].size)
            a.denorm = b["denorm"].ctypes.data; a.denorm_len = b["denorm"].size; a.num_norm_qs = enc["num_norm_qs"]
            a.line_len = ln.ctypes.data; a.n_lines = ln.size; a.out = out.ctypes.data; a.out_cap = tot
        rc = self.L.gzb_domq_reconstruct(self.h, arr, n, 0)
        if rc != 0:
            raise GzbError(f"gzb_domq_reconstruct failed ({rc}): {self._err()}")
        return [o[:-1] for o in outs]

    # ---- PBWT (host buffers) ----
    def pbwt_encode(self, ht):
        """codec_pbwt_compress: uint8 matrix [n_lines, ht_per_line] -> (RUNS u32, FGRC u32), host-endian"""
        ht = np.ascontiguousarray(ht, dtype=np.uint8)
        n_lines, w = ht.shape
        runs = np.zeros(2 * ht.size + 8, np.uint32); fgrc = np.zeros(ht.size + 8, np.uint32)
        nr, nf = C.c_uint32(), C.c_uint32()
        rc = self.L.gzb_pbwt_encode(self.h, ht.ctypes.data, n_lines, w, runs.ctypes.data, runs.size, C.byref(nr),
                                    fgrc.ctypes.data, fgrc.size, C.byref(nf), 0)
        if rc != 0:
            raise GzbError(f"gzb_pbwt_encode failed ({rc}): {self._err()}")
        return runs[:nr.value].copy(), fgrc[:nf.value].copy()

    def pbwt_decode(self, runs, fgrc, n_lines, size):
        runs = np.ascontiguousarray(runs, np.uint32); fgrc = np.ascontiguousarray(fgrc, np.uint32)
        ht = np.zeros(size, np.uint8)
        hl = C.c_uint64()
        rc = self.L.gzb_pbwt_decode(self.h, runs.ctypes.data, runs.size, fgrc.ctypes.data, fgrc.size, n_lines, ht.ctypes.data, size, C.byref(hl), 0)
        if rc != 0:
            raise GzbError(f"gzb_pbwt_decode failed ({rc}): {self._err()}")
        return ht[:hl.value]

    def pbwt_encode_batch(self, hts, runs_cap=None, fgrc_cap=None):
        """codec_pbwt_compress for the matrices of a batch of VBlocks in one call -> [(RUNS, FGRC)]"""
        hts = [np.ascontiguousarray(h, dtype=np.uint8) for h in hts]
        arr = (PbwtVb * len(hts))()
        keep = []
        for a, ht in zip(arr, hts):
            runs = np.zeros(runs_cap or (2 * ht.size + 8), np.uint32); fgrc = np.zeros(fgrc_cap or (ht.size + 8), np.uint32)
            keep.append((runs, fgrc))
            a.ht = ht.ctypes.data; a.n_lines, a.ht_per_line = ht.shape
            a.runs = runs.ctypes.data; a.runs_cap = runs.size; a.fgrc = fgrc.ctypes.data; a.fgrc_cap = fgrc.size
        rc = self.L.gzb_pbwt_encode_batch(self.h, arr, len(hts), 0)
        if rc != 0:
            raise GzbError(f"gzb_pbwt_encode_batch failed ({rc}): {self._err()}")
        return [(r[:a.n_runs].copy(), f[:a.n_fgrc].copy()) for a, (r, f) in zip(arr, keep)]

    def pbwt_decode_batch(self, streams, n_lines, sizes):
        """codec_pbwt_uncompress for a batch: streams = [(RUNS, FGRC)], n_lines / sizes per VBlock -> [matrix bytes]"""
        arr = (PbwtVb * len(streams))()
        keep = []
        for a, (runs, fgrc), nl, size in zip(arr, streams, n_lines, sizes):
            runs = np.ascontiguousarray(runs, np.uint32); fgrc = np.ascontiguousarray(fgrc, np.uint32); ht = np.zeros(size, np.uint8)
            keep.append((runs, fgrc, ht))
            a.ht = ht.ctypes.data; a.ht_cap = size; a.n_lines = nl
            a.runs = runs.ctypes.data; a.n_runs = runs.size; a.fgrc = fgrc.ctypes.data; a.n_fgrc = fgrc.size
        rc = self.L.gzb_pbwt_decode_batch(self.h, arr, len(streams), 0)
        if rc != 0:
            raise GzbError(f"gzb_pbwt_decode_batch failed ({rc}): {self._err()}")
        return [k[2][:a.ht_len] for a, k in zip(arr, keep)]

    # ---- LONGR (host buffers), batch of VBlocks: each (txt, seq_off, qual_off, lens, is_rev|None, value_to_bin) ----
    def _longr_arr(self, vbs, values=None, lens_be=None, decode=False):
        n = len(vbs)
        arr = (LongrVb * n)()
        keep = []
        for i, vb in enumerate(vbs):
            txt, seq_off, qual_off, lens, is_rev, v2b = vb[:6]
            qlens = vb[6] if len(vb) > 6 else None                           # quality lengths where they differ from lens (lines without quality)
            txt = np.ascontiguousarray(txt, np.uint8); seq_off = np.ascontiguousarray(seq_off, np.uint64)
            qual_off = np.ascontiguousarray(qual_off, np.uint64); lens = np.ascontiguousarray(lens, np.uint32)
            ql = None if qlens is None else np.ascontiguousarray(qlens, np.uint32)
            tot = int(lens.sum())
            vals = np.zeros(tot + 1, np.uint8) if values is None else np.ascontiguousarray(values[i], np.uint8)
            lb = np.zeros(65536, np.uint32) if lens_be is None else np.ascontiguousarray(lens_be[i], np.uint32)
            qo = np.zeros(tot + 1, np.uint8)
            rv = None if is_rev is None else np.ascontiguousarray(is_rev, np.uint8)
            ms = np.zeros(lens.size + 1, np.uint8)
            keep.append((txt, seq_off, qual_off, lens, rv, vals, lb, qo, ms, ql))
            a = arr[i]
            a.txt = txt.ctypes.data; a.txt_len = txt.size; a.seq_off = seq_off.ctypes.data; a.qual_off = qual_off.ctypes.data
            a.len = lens.ctypes.data; a.is_rev = None if rv is None else rv.ctypes.data; a.n_lines = lens.size
            C.memmove(a.value_to_bin, np.ascontiguousarray(v2b, np.uint8).ctypes.data, 256)
            a.values = vals.ctypes.data; a.lens_be = lb.ctypes.data; a.qual_out = qo.ctypes.data
            a.missing = ms.ctypes.data if decode else None
            a.n_bases = vals.size if (decode and values is not None) else 0
            a.qual_len = None if (ql is None or decode) else ql.ctypes.data
        return arr, keep

    def longr_encode(self, vbs):
        arr, keep = self._longr_arr(vbs)
        rc = self.L.gzb_longr_encode(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_longr_encode failed ({rc}): {self._err()}")
        return [(k[5][:int((k[3] if k[9] is None else k[9]).sum())].copy(), k[6].copy()) for k in keep]

    def longr_decode(self, vbs, values, lens_be):
        arr, keep = self._longr_arr(vbs, values, lens_be, decode=True)
        rc = self.L.gzb_longr_decode(self.h, arr, len(vbs), 0)
        if rc != 0:
            raise GzbError(f"gzb_longr_decode failed ({rc}): {self._err()}")
        self.longr_missing = [k[8][:k[3].size].copy() for k in keep]
        return [k[7][:int(k[3].sum())].copy() for k in keep]

    def longr_calculate_bins(self, vb):
        """codec_longr_segconf_calculate_bins -> value_to_bin (256 bytes), or None when the VBlock has no quality"""
        arr, keep = self._longr_arr([tuple(vb[:5]) + (np.zeros(256, np.uint8),) + tuple(vb[6:])])
        v2b = np.zeros(256, np.uint8)
        rc = self.L.gzb_longr_calculate_bins(self.h, arr, 0, v2b.ctypes.data)
        if rc == 1:
            return None
        if rc != 0:
            raise GzbError(f"gzb_longr_calculate_bins failed ({rc}): {self._err()}")
        return v2b
