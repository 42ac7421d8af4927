"""Host-side driver of the per-VBlock codec path for a batch of FASTQ VBlocks — the Python mirror of what genozip's
compute thread does between segmentation and z_data assembly for the contexts on this path
(zip_compress_all_contexts_local → comp_compress → codec_args[].compress, src/zip.c:291, src/compressor.c:18-182;
and piz_uncompress_all_ctxs → comp_uncompress, src/piz.c:247, src/compressor.c:211-255):

  ZIP   SEQ  (NONREF.local)  --codec_acgt_compress-->  2-bit words (sub-codec LZMA stays on the host: out of scope)
                                                       + NONREF_X.local --XCGT sub-codec--> section
        QUAL (QUAL.local)    --codec_domq_compress-->  QUAL.local / DOMQRUNS / QUALMPLX / DIVRQUAL --sub-codecs--> sections
        read-name contexts   --simple codecs-->        sections
  PIZ   the inverse.

Everything numeric happens in libgzb200.so (CUDA); torch only owns the device / pinned host buffers.  Independent
VBlocks are sharded round-robin over the GPUs of the box by vblock_i (gzb_vb_device), one process per GPU.
"""
import ctypes as C
import threading
import numpy as np
import torch

from concurrent.futures import ThreadPoolExecutor

from .lib import (load, Engine, Section, DomqVb, DomqPizVb, AcgtVb, Copy, CODEC, est_size, GzbError,
                  GZB_DEVICE_PTRS, GZB_OUT_DEVICE, GZB_IN_DEVICE, GZB_SEC_IN_DEVICE, GZB_SEC_OUT_DEVICE)

NAME_LEN = 45            # "@A00123:45:HXXXXXXXX:1:1101:12345:12345 1:N:0:ACGT" without the newline ~ 45-50
SIMPLE = ["RANB", "RANW", "RANb", "RANw", "ARTB", "ARTW", "ARTb", "ARTw"]   # ascending Codec enum order (ties -> first)
STREAMS = ["QUAL", "DOMQRUNS", "QUALMPLX", "DIVRQUAL", "NONREF_X", "Q_TILE", "Q_X", "Q_Y", "Q_MISC"]


def txt_bytes_per_vb(n_reads, read_len):
    """FASTQ text a VBlock of n_reads represents: name line, SEQ, '+', QUAL and 4 newlines per record"""
    return n_reads * (NAME_LEN + 1 + read_len + 1 + 1 + 1 + read_len + 1)


def synth_vblocks(V, n_reads, read_len, seed, device):
    """Synthetic Illumina-like VBlocks generated on the device: returns dict of uint8 device tensors
    [V, ...]: seq, qual (fixed-length lines), and the read-name context streams.  Generated in chunks of 8 VBlocks
    (torch's samplers index with 32 bits)."""
    parts = [_synth_chunk(min(8, V - v0), n_reads, read_len, seed * 100003 + v0, device) for v0 in range(0, V, 8)]
    return {k: torch.cat([p[k] for p in parts], 0).contiguous() for k in parts[0]}


def _synth_chunk(V, n_reads, read_len, seed, device):
    g = torch.Generator(device=device); g.manual_seed(seed)
    n = n_reads * read_len
    acgt = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=device)
    seq = acgt[torch.randint(0, 4, (V, n), generator=g, device=device)]
    seq[torch.rand((V, n), generator=g, device=device) < 0.001] = ord("N")
    # QUAL: binned Illumina {F:88%, ':':7%, ',':4%, '#':1%} with Markov run structure, P(stay) = 0.97
    syms = torch.tensor(list(b"F:,#"), dtype=torch.uint8, device=device)
    pick = torch.multinomial(torch.tensor([.88, .07, .04, .01], device=device), V * n, replacement=True, generator=g).view(V, n)
    change = torch.rand((V, n), generator=g, device=device) > 0.97
    change[:, 0] = True
    idx = torch.where(change, torch.arange(n, device=device).expand(V, n), torch.zeros((), dtype=torch.long, device=device))
    last = torch.cummax(idx, dim=1).values
    qual = syms[torch.gather(pick, 1, last)]
    del pick, change, idx, last
    # ~3% diverse lines
    q2 = qual.view(V, n_reads, read_len)
    div = torch.rand((V, n_reads), generator=g, device=device) < 0.03
    nd = int(div.sum().item())
    if nd:
        q2[div] = syms[torch.multinomial(torch.tensor([.4, .3, .2, .1], device=device), nd * read_len, replacement=True, generator=g).view(nd, read_len)]
    # read-name contexts: tile (b250, long runs), x / y (uint32 big-endian locals), misc b250
    tile = (torch.arange(n_reads, device=device) // 977 % 96).to(torch.uint8).expand(V, n_reads).contiguous()
    xs = (torch.cumsum(torch.randint(0, 60, (V, n_reads), generator=g, device=device), 1) % 30000 + 1000).to(torch.int32)
    ys = torch.randint(1000, 30000, (V, n_reads), generator=g, device=device, dtype=torch.int32)

    def be32(t):
        b = t.contiguous().view(torch.uint8).view(V, n_reads, 4)
        return b.flip(2).contiguous().view(V, n_reads * 4)
    misc = torch.multinomial(torch.tensor([.9, .05, .03, .02], device=device), V * n_reads, replacement=True, generator=g).view(V, n_reads).to(torch.uint8)
    return dict(seq=seq.contiguous(), qual=qual.contiguous(), Q_TILE=tile, Q_X=be32(xs), Q_Y=be32(ys), Q_MISC=misc.contiguous())


def _pin(t):
    """pinned host memory where there is a GPU to transfer to"""
    return t.pin_memory() if torch.cuda.is_available() else t


S_IDX = {s: i for i, s in enumerate(STREAMS)}
DQ = ("QUAL", "DOMQRUNS", "QUALMPLX", "DIVRQUAL")                            # the four streams codec_domq_compress hands on
DQ_FLD = ("qual", "runs", "mplx", "divr")
NAMES = ("Q_TILE", "Q_X", "Q_Y", "Q_MISC")


def struct_view(arr):
    """numpy structured view of a ctypes Structure array (shares memory): descriptor arrays are filled column-wise, not field by field"""
    T = arr._type_
    names, formats, offsets = [], [], []
    for name, ct in T._fields_:
        names.append(name); offsets.append(getattr(T, name).offset)
        if issubclass(ct, C.Array):
            formats.append((np.uint8, (ct._length_,)))
        else:
            formats.append("i4" if ct is C.c_int32 else {1: "u1", 2: "u2", 4: "u4", 8: "u8"}[C.sizeof(ct)])
    dt = np.dtype(dict(names=names, formats=formats, offsets=offsets, itemsize=C.sizeof(T)))
    return np.frombuffer(arr, dtype=dt)


class ZipMeta:
    """what a zip pass leaves behind for the piz pass and the section list: per VBlock and stream the uncompressed length, the
    compressed length and where the compressed section lies (address inside the packed output buffer)."""

    def __init__(self, V, streams=None):
        self.V = V
        self.streams = STREAMS = list(streams if streams is not None else globals()["STREAMS"])
        self.len = np.zeros((V, len(STREAMS)), np.int64)
        self.comp_len = np.zeros((V, len(STREAMS)), np.int64)
        self.comp_ptr = np.zeros((V, len(STREAMS)), np.uint64)
        self.dq_off = np.zeros((V, 4), np.int64)                             # offsets of the DOMQ streams inside the compact intermediate buffer
        self.acgt_no_x = np.zeros(V, bool)
        self.num_norm_qs = np.zeros(V, np.uint8); self.num_doms = np.zeros(V, np.uint8)
        self.denorm = np.zeros((V, 95 * 95), np.uint8)

    def __len__(self):
        return self.V

    def __getitem__(self, v):
        return dict(len={s: int(self.len[v, i]) for i, s in enumerate(self.streams)},
                    comp_len={s: int(self.comp_len[v, i]) for i, s in enumerate(self.streams) if self.len[v, i]},
                    acgt_no_x=bool(self.acgt_no_x[v]), num_norm_qs=int(self.num_norm_qs[v]),
                    denorm=bytes(self.denorm[v, :int(self.num_norm_qs[v]) * int(self.num_doms[v])]))

    def __iter__(self):
        return (self[v] for v in range(self.V))


class FastqCodecPath:
    """zip / piz of a batch of V FASTQ VBlocks through libgzb200 on one GPU.

    Memory is what bounds the batch, and the batch is what bounds throughput (the entropy chains are latency-bound: a launch lasts
    as long as its longest leaf whatever the number of VBlocks).  So nothing of worst-case size exists per VBlock:
      * codec_domq_compress writes its four streams into worst-case buffers (QUAL.local 2n, DOMQRUNS n, DIVRQUAL n: codec_domq.c:
        395-410) of a SUB-BATCH of VBlocks; the streams are then moved, at their real lengths, into one compact buffer
        (gzb_copy_batch) — the reference does the same when it truncates ctx->local.len after the codec ran;
      * compressed sections are appended to one packed buffer (gzb_compress_sections_packed), as zfile_compress_local_data
        appends to vb->z_data;
      * piz decodes the intermediate streams into the buffers zip's intermediates occupied (scrub_intermediates() empties them
        first when a test wants to see that piz really produced them).
    Descriptor arrays are filled column-wise through numpy views (no per-section Python work).

    The host-buffer path gives each of the three independent pipelines of a FASTQ VBlock (QUAL, SEQ, read names) its own engine
    (host thread + CUDA stream) so that transfers overlap the entropy chains (zip_host / piz_host)."""

    def __init__(self, eng: Engine, V, n_reads, read_len, n_engines=3, sub_batch=128, fields=None, seq_len=None):
        """fields: {context name: bytes per VBlock} of the b250 / local streams that go through the simple codecs beside QUAL and SEQ
        (default: the four read-name contexts of a FASTQ VBlock); seq_len: bases handed to codec_acgt per VBlock (default: every base
        of the reads, FASTQ's NONREF.local; an aligned BAM VBlock hands over only the bases its reference does not explain)."""
        self.eng, self.L = eng, eng.L
        self.V, self.n_reads, self.read_len = V, n_reads, read_len
        self.n = n = n_reads * read_len
        self.n_seq = n_seq = n if seq_len is None else int(seq_len)
        fields = dict(fields) if fields is not None else {"Q_TILE": n_reads, "Q_X": 4 * n_reads, "Q_Y": 4 * n_reads, "Q_MISC": n_reads}
        self.NAMES = tuple(fields)
        self.STREAMS = list(DQ) + ["NONREF_X"] + list(self.NAMES)
        self.S_IDX = {s: i for i, s in enumerate(self.STREAMS)}
        dev = torch.device(getattr(eng, "torch_device", None) or f"cuda:{eng.device}")   # (the CPU test suite drives this class through a mock engine)
        self.dev = dev
        n_engines = max(1, n_engines)
        self.engs = [eng] + [type(eng)(eng.device) for _ in range(n_engines - 1)]
        self.pool = ThreadPoolExecutor(n_engines) if n_engines > 1 else None
        self.stream = torch.cuda.ExternalStream(self.L.gzb_engine_stream(eng.h), device=dev) if dev.type == "cuda" else None
        self.SB = max(1, min(sub_batch, V))
        self.packed_len = int(self.L.gzb_acgt_packed_len(n_seq))
        u8 = dict(dtype=torch.uint8, device=dev)
        self.line_off_h = _pin(torch.arange(n_reads, dtype=torch.int64) * read_len)
        self.line_len_h = _pin(torch.full((n_reads,), read_len, dtype=torch.int32))
        self.line_off_d, self.line_len_d = self.line_off_h.to(dev), self.line_len_h.to(dev)
        self.packed_d = torch.empty((V, self.packed_len + 32), **u8)
        self.x_d = torch.empty((V, n_seq), **u8)
        self.linedom_d = torch.empty((V, n_reads), **u8)
        self.linediv_d = torch.empty((V, n_reads), **u8)
        self.caps = {"QUAL": 2 * n + 16, "DOMQRUNS": n + 16, "QUALMPLX": n_reads + 16, "DIVRQUAL": n + 16}
        # worst-case scratch of one sub-batch per engine that runs codec_domq_compress (engine 0 only)
        self.dqs = {s: torch.empty((self.SB, c), **u8) for s, c in self.caps.items()}
        self.name_len = {s: int(b) for s, b in fields.items()}
        self.dq_arena = None                                                  # compact DOMQ streams of all V VBlocks
        self.comp_arena = None                                                # packed compressed sections (device): the QUAL pipeline's
        self.comp_arena2 = None                                               #   and the others'
        self.device_pipelines = 2 if n_engines >= 2 else 1
        self.comp_used = {}                                                   # bytes of packed sections in each output buffer after the last zip
        self.codec = {s: "RANB" for s in self.STREAMS}
        self.dvb = (DomqVb * V)(); self.pvb = (DomqPizVb * V)(); self.avb = (AcgtVb * V)()
        self.dvb_np, self.pvb_np, self.avb_np = struct_view(self.dvb), struct_view(self.pvb), struct_view(self.avb)
        self.meta = None
        self.h = {}
        self.kernel_ms = (0.0, 0.0)   # chain kernel durations of the last call: (rANS, arithmetic), max over the engines
        self.names_dec_d = None; self.seq_out_d = None; self.qual_out_d = None

    @property
    def launches(self):
        return sum(e.launches for e in self.engs)

    @property
    def groups(self):
        return [(0, self.V)]

    def close(self):
        """release the engines this path created (not the caller's) and its host threads"""
        if self.pool is not None:
            self.pool.shutdown(wait=True); self.pool = None
        for e in self.engs[1:]:
            e.close()
        self.engs = self.engs[:1]

    def release_device(self):
        """drop the device-resident leg's buffers (the host-buffer leg that follows allocates its own)"""
        self.close()
        for a in ("packed_d", "x_d", "linedom_d", "linediv_d", "dq_arena", "comp_arena", "comp_arena2", "names_dec_d", "seq_out_d", "qual_out_d", "dec_d"):
            setattr(self, a, None)
        self.dqs = {}

    @staticmethod
    def _sub(arr, v0, v1):
        """ctypes view of elements [v0, v1) of a ctypes array (shares memory)"""
        return (arr._type_ * (v1 - v0)).from_buffer(arr, v0 * C.sizeof(arr._type_))

    @staticmethod
    def _rows(t):
        """addresses of the rows of a contiguous 2-D tensor"""
        return np.uint64(t.data_ptr()) + np.arange(t.shape[0], dtype=np.uint64) * np.uint64(t.stride(0) * t.element_size())

    def _kernel_ms(self, engs):
        """(rANS chain kernel, the arithmetic chain kernels — they run side by side: the longest) of the last call, max over the engines;
        kernel_ms_detail: general / order-0 / split-encoder arithmetic kernels separately"""
        ms = lambda w: float(np.max([self.L.gzb_last_kernel_ms(e.h, w) for e in engs]))
        self.kernel_ms = (ms(0), ms(5))
        self.kernel_ms_detail = {"rans": ms(0), "arith_general": ms(1), "arith_o0": ms(3), "arith_split": ms(4), "split_bucket": ms(6), "split_model": ms(7), "split_code": ms(8)}

    # ------------------------------------------------------------------ codec assignment (host policy, run on the GPU)
    def assign_codecs(self, data):
        """codec_assign_best_codec's size criterion (src/codec.c:234-389, sorter :128-173) restricted to the eight
        in-scope simple codecs: compress the first <=99,999 bytes (CODEC_ASSIGN_SAMPLE_SIZE, src/codec.h:154) of VB 1's
        stream with each and keep the smallest, ties to the lower Codec value.  The reference also weighs clock() time
        (timing-dependent, H5) — not reproduced.  Samples are compressed on the GPU (same bytes as the reference)."""
        meta = ZipMeta(self.V, self.STREAMS)
        self._acgt_pack_device(data, meta, 0, 1)
        self._domq_device(lambda v: data["qual"][v].data_ptr(), self.engs[0], meta, GZB_DEVICE_PTRS, 0, 1)
        meta.len[0, [self.S_IDX[s] for s in self.NAMES]] = [self.name_len[s] for s in self.NAMES]
        present = [s for s in self.STREAMS if int(meta.len[0, self.S_IDX[s]])]
        # one gzb_assign_codecs call: the first <= 99,999 bytes of every stream of VB 1, where they are in HBM, with the eight codecs; sizes only
        res = self.eng.assign_codecs_ptrs([(self._stream_tensor(s, 0, data, meta).data_ptr(), int(meta.len[0, self.S_IDX[s]])) for s in present], GZB_DEVICE_PTRS)
        for s, (best, sizes) in zip(present, res):
            # a sample below 50 B would go out as CODEC_NONE (compressor.c:56-58), and so would one no codec shrinks: this path has no
            # uncompressed sections, it keeps the smallest of the eight then
            self.codec[s] = best if best in SIMPLE else (min(SIMPLE, key=lambda c: (sizes[c], SIMPLE.index(c))) if sizes else "RANB")
        return dict(self.codec)

    def _stream_tensor(self, s, v, data, meta):
        """device tensor holding stream s of VBlock v (whole capacity for the inputs; the real length for the DOMQ streams)"""
        if s in DQ:
            o = int(meta.dq_off[v, DQ.index(s)])
            return self.dq_arena[o: o + int(meta.len[v, self.S_IDX[s]])]
        if s == "NONREF_X":
            return self.x_d[v]
        return data[s][v]

    # ------------------------------------------------------------------ the domain codecs of a batch
    def _acgt_pack_device(self, data, meta, v0, v1, eng=None):
        eng = eng or self.eng
        a = self.avb_np
        a["seq"][v0:v1] = self._rows(data["seq"])[v0:v1]; a["n_bases"][v0:v1] = self.n_seq
        a["packed"][v0:v1] = self._rows(self.packed_d)[v0:v1]; a["x"][v0:v1] = self._rows(self.x_d)[v0:v1]
        if self.L.gzb_acgt_pack_batch(eng.h, self._sub(self.avb, v0, v1), v1 - v0, GZB_DEVICE_PTRS):
            raise GzbError(f"gzb_acgt_pack_batch: {eng._err()}")
        meta.acgt_no_x[v0:v1] = a["x_all_zero"][v0:v1] != 0
        meta.len[v0:v1, self.S_IDX["NONREF_X"]] = np.where(meta.acgt_no_x[v0:v1], 0, self.n_seq)

    def _dq_reserve(self, need, used, total_guess):
        """the compact buffer of the DOMQ streams: sized from the first sub-batch, grown (contents kept) if that was too little"""
        if self.dq_arena is None or self.dq_arena.numel() < need:
            new = torch.empty(max(need, int(total_guess)), dtype=torch.uint8, device=self.dev)
            if self.dq_arena is not None and used:
                new[:used] = self.dq_arena[:used]
            self.dq_arena = new

    def _domq_device(self, txt_ptr, eng, meta, flags, v0, v1, line_tables=None):
        """codec_domq_compress for VBlocks [v0, v1) in sub-batches: worst-case scratch -> the compact buffer.  txt_ptr(v) = address
        of the VBlock's quality text (device with GZB_DEVICE_PTRS, host with GZB_OUT_DEVICE)."""
        L, d = self.L, self.dvb_np
        host = not (flags & GZB_DEVICE_PTRS)
        lo, ll = (self.line_off_h, self.line_len_h) if host else (self.line_off_d, self.line_len_d)
        cursor = int(meta._dq_cursor) if hasattr(meta, "_dq_cursor") else 0
        for b0 in range(v0, v1, self.SB):
            b1 = min(v1, b0 + self.SB); k = b1 - b0
            d["txt"][b0:b1] = [txt_ptr(v) for v in range(b0, b1)]; d["txt_len"][b0:b1] = self.n
            d["line_off"][b0:b1] = lo.data_ptr(); d["line_len"][b0:b1] = ll.data_ptr(); d["n_lines"][b0:b1] = self.n_reads
            d["line_dom"][b0:b1] = self._rows(self.h["linedom"] if host else self.linedom_d)[b0:b1]
            d["line_diverse"][b0:b1] = self._rows(self.h["linediv"] if host else self.linediv_d)[b0:b1]
            for fld, s in zip(DQ_FLD, DQ):
                d[fld][b0:b1] = self._rows(self.dqs[s])[:k]; d[fld + "_cap"][b0:b1] = self.caps[s]
            dv = self._sub(self.dvb, b0, b1)
            if L.gzb_domq_prepare(eng.h, dv, k, flags) or L.gzb_domq_split(eng.h, dv, k, flags):
                raise GzbError(f"gzb_domq: {eng._err()}")
            lens = np.stack([d[f + "_len"][b0:b1].astype(np.int64) for f in DQ_FLD], 1)          # [k, 4]
            al = (lens + 255) & ~255
            offs = cursor + np.concatenate([[0], np.cumsum(al.reshape(-1))[:-1]]).reshape(k, 4)
            need = cursor + int(al.sum())
            self._dq_reserve(need, cursor, need * (1.0 + 1.1 * (self.V - b1) / max(1, b1 - v0)) if b0 == v0 else need * 1.25)
            cp = (Copy * (4 * k))(); c = struct_view(cp)
            c["src"] = np.stack([d[f][b0:b1] for f in DQ_FLD], 1).reshape(-1)
            c["dst"] = (np.uint64(self.dq_arena.data_ptr()) + offs.astype(np.uint64)).reshape(-1)
            c["len"] = lens.reshape(-1)
            if L.gzb_copy_batch(eng.h, cp, 4 * k):
                raise GzbError(f"gzb_copy_batch: {eng._err()}")
            cursor = need
            meta.dq_off[b0:b1] = offs
            for j, s in enumerate(DQ):
                meta.len[b0:b1, self.S_IDX[s]] = lens[:, j]
            meta.num_norm_qs[b0:b1] = d["num_norm_qs"][b0:b1]; meta.num_doms[b0:b1] = d["num_doms"][b0:b1]
            meta.denorm[b0:b1] = d["denorm"][b0:b1]
        meta._dq_cursor = cursor

    def _section_array(self, meta, in_ptr, names, sflags=0):
        """gzb_section descriptors of the non-empty streams `names` of every VBlock, VBlock-major; returns (array, view, v index, stream index)"""
        cols = [self.S_IDX[s] for s in names]
        ln = meta.len[:, cols]
        vv, jj = np.nonzero(ln)
        ss = np.asarray(cols)[jj]
        secs = (Section * max(1, vv.size))(); a = struct_view(secs)
        a["codec"][:vv.size] = np.asarray([CODEC[self.codec[s]] for s in self.STREAMS], np.int32)[ss]
        a["in_"][:vv.size] = in_ptr[vv, ss]; a["in_len"][:vv.size] = ln[vv, jj]
        a["sflags"][:vv.size] = np.asarray(sflags, np.uint32)[ss] if not np.isscalar(sflags) else sflags
        return secs, a, vv, ss

    def _in_ptrs(self, meta, name_rows, dq_base):
        """[V, 9] addresses of the uncompressed streams: DOMQ streams in the compact buffer, the exception stream, the read-name contexts"""
        p = np.zeros((self.V, len(self.STREAMS)), np.uint64)
        for j, s in enumerate(DQ):
            p[:, self.S_IDX[s]] = np.uint64(dq_base) + meta.dq_off[:, j].astype(np.uint64)
        p[:, self.S_IDX["NONREF_X"]] = self._rows(self.x_d)
        for s in self.NAMES:
            p[:, self.S_IDX[s]] = name_rows[s]
        return p

    def _compress_packed(self, eng, secs, a, n, flags, arena_attr, host):
        """gzb_compress_sections_packed into the buffer self.<arena_attr> (device or pinned host), grown and repeated when too small"""
        if not n:
            return
        while True:
            arena = getattr(self, arena_attr, None) if not host else self.h.get(arena_attr)
            if arena is None:
                guess = int(a["in_len"][:n].sum() // 8) + (1 << 20)
                arena = _pin(torch.empty(guess, dtype=torch.uint8)) if host else torch.empty(guess, dtype=torch.uint8, device=self.dev)
                if host: self.h[arena_attr] = arena
                else: setattr(self, arena_attr, arena)
            used = C.c_uint64()
            rc = self.L.gzb_compress_sections_packed(eng.h, secs, n, arena.data_ptr(), arena.numel(), C.byref(used), flags)
            if rc == 1 and used.value > arena.numel():
                bigger = int(used.value * 1.1) + 4096
                new = _pin(torch.empty(bigger, dtype=torch.uint8)) if host else torch.empty(bigger, dtype=torch.uint8, device=self.dev)
                if host: self.h[arena_attr] = new
                else: setattr(self, arena_attr, new)
                continue
            if rc:
                raise GzbError(f"gzb_compress_sections_packed failed ({rc}): {eng._err()}")
            if a["status"][:n].any():
                i = int(np.nonzero(a["status"][:n])[0][0])
                raise GzbError(f"section {i}: status {int(a['status'][i])}")
            self.comp_used[arena_attr] = int(used.value)
            return

    # ------------------------------------------------------------------ ZIP, inputs resident in HBM
    def zip_device(self, data):
        """inputs resident in HBM.  With two or more engines the QUAL pipeline (codec_domq_compress + its four sub-codec sections)
        and the rest (codec_acgt_compress + the exception stream, the read-name contexts) run on one engine / host thread each: the
        DOMQ passes of one overlap the entropy chains of the other — as two compute threads of the reference would."""
        V = self.V
        meta = ZipMeta(V, self.STREAMS)
        for s in self.NAMES:
            meta.len[:, self.S_IDX[s]] = self.name_len[s]
        name_rows = {s: self._rows(data[s]) for s in self.NAMES}

        def compress(eng, names, arena):
            inp = self._in_ptrs(meta, name_rows, self.dq_arena.data_ptr() if self.dq_arena is not None else 0)
            secs, a, vv, ss = self._section_array(meta, inp, names)
            self._compress_packed(eng, secs, a, vv.size, GZB_DEVICE_PTRS, arena, False)
            meta.comp_len[vv, ss] = a["out_len"][:vv.size]; meta.comp_ptr[vv, ss] = a["out"][:vv.size]

        # The chain kernels are long-running and fill the SMs' registers and shared memory: a short kernel launched after them (on any
        # stream) only gets in when their CTAs retire.  So every bandwidth-shaped pass of the batch — ACGT pack, the DOMQ passes —
        # runs BEFORE the first chain kernel of either pipeline is launched (measured: a DOMQ sub-batch behind the other pipeline's
        # rANS chains waited 190 ms for its 13 ms of work).
        domq_done = threading.Event()

        def part_qual(eng):
            try:
                self._domq_device(lambda v: data["qual"][v].data_ptr(), eng, meta, GZB_DEVICE_PTRS, 0, V)
            finally:
                domq_done.set()
            compress(eng, DQ, "comp_arena")

        def part_rest(eng):
            self._acgt_pack_device(data, meta, 0, V, eng)
            domq_done.wait()
            compress(eng, ("NONREF_X",) + self.NAMES, "comp_arena2")

        if self.device_pipelines >= 2 and self.pool is not None and len(self.engs) >= 2:
            self._run_parts([part_qual, part_rest])
            self._kernel_ms(self.engs[:2])
        else:
            part_qual(self.eng); part_rest(self.eng)
            self._kernel_ms([self.eng])
        self.meta = meta
        return meta

    def section_bytes(self, meta, v, s, host=False):
        """the compressed section of stream s of VBlock v as a numpy array"""
        i = self.S_IDX[s]
        ln = int(meta.comp_len[v, i])
        arena = self.h["comp_" + self._pipeline_of(s)] if host else (self.comp_arena if s in DQ else self.comp_arena2)
        o = int(meta.comp_ptr[v, i]) - arena.data_ptr()
        return arena[o: o + ln].cpu().numpy()

    @staticmethod
    def _pipeline_of(s):
        return "qual" if s in DQ else "seq" if s == "NONREF_X" else "names"

    # ------------------------------------------------------------------ PIZ, inputs resident in HBM
    def alloc_piz(self, meta):
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.names_dec_d = {s: torch.empty((self.V, self.name_len[s] + 16), **u8) for s in self.NAMES}
        self.seq_out_d = torch.empty((self.V, self.n_seq), **u8)
        self.qual_out_d = torch.empty((self.V, self.n), **u8)
        self.dec_d = self.names_dec_d                                          # (the decoded read-name contexts, by stream)

    def scrub_intermediates(self):
        """empty every buffer piz decodes into that zip had filled with the same bytes (the DOMQ streams, the exception stream), so
        that a round-trip check sees what piz produced"""
        if self.dq_arena is not None: self.dq_arena.zero_()
        self.x_d.zero_()
        if self.names_dec_d is not None:
            for t in self.names_dec_d.values(): t.zero_()
        if self.seq_out_d is not None: self.seq_out_d.zero_(); self.qual_out_d.zero_()

    def _fill_piz_descriptors(self, meta, qual_out_rows, seq_out_rows, packed_rows, line_len):
        p, a = self.pvb_np, self.avb_np
        base = np.uint64(self.dq_arena.data_ptr())
        for j, (fld, s) in enumerate(zip(DQ_FLD, DQ)):
            p[fld] = base + meta.dq_off[:, j].astype(np.uint64); p[fld + "_len"] = meta.len[:, self.S_IDX[s]]
        p["denorm"] = np.uint64(meta.denorm.ctypes.data) + np.arange(self.V, dtype=np.uint64) * np.uint64(95 * 95)
        p["denorm_len"] = meta.num_norm_qs.astype(np.uint32) * meta.num_doms.astype(np.uint32); p["num_norm_qs"] = meta.num_norm_qs
        p["line_len"] = line_len.data_ptr(); p["n_lines"] = self.n_reads
        p["out"] = qual_out_rows; p["out_cap"] = self.n
        a["seq"] = seq_out_rows; a["n_bases"] = self.n_seq; a["packed"] = packed_rows
        a["x"] = np.where(meta.acgt_no_x, np.uint64(0), self._rows(self.x_d))

    def piz_device(self, meta, outs=None):
        """outs: where the reconstructed streams go — {"seq", "qual", "Q_TILE", ...} device tensors [V, ...]; default: the buffers of alloc_piz"""
        L = self.L
        if outs is None:
            outs = dict(self.names_dec_d, seq=self.seq_out_d, qual=self.qual_out_d)
        outp = self._in_ptrs(meta, {s: self._rows(outs[s]) for s in self.NAMES}, self.dq_arena.data_ptr())
        self._fill_piz_descriptors(meta, self._rows(outs["qual"]), self._rows(outs["seq"]), self._rows(self.packed_d), self.line_len_d)

        def uncompress(eng, names):
            secs, a, vv, ss = self._section_array(meta, meta.comp_ptr, names)
            a["in_len"][:vv.size] = meta.comp_len[vv, ss]; a["out"][:vv.size] = outp[vv, ss]; a["out_cap"][:vv.size] = meta.len[vv, ss]
            if vv.size:
                eng.uncompress_raw(secs, vv.size, GZB_DEVICE_PTRS)

        def part_qual(eng):
            uncompress(eng, DQ)
            if L.gzb_domq_reconstruct(eng.h, self.pvb, self.V, GZB_DEVICE_PTRS):
                raise GzbError(f"gzb_domq_reconstruct: {eng._err()}")

        def part_rest(eng):
            uncompress(eng, ("NONREF_X",) + self.NAMES)
            if L.gzb_acgt_unpack_batch(eng.h, self.avb, self.V, GZB_DEVICE_PTRS):
                raise GzbError(f"gzb_acgt_unpack_batch: {eng._err()}")

        if self.device_pipelines >= 2 and self.pool is not None and len(self.engs) >= 2:
            self._run_parts([part_qual, part_rest])
            self._kernel_ms(self.engs[:2])
        else:
            part_qual(self.eng); part_rest(self.eng)
            self._kernel_ms([self.eng])

    # ------------------------------------------------------------------ HOST-buffer path (e2e): what the C host would call
    def alloc_host(self, data):
        """pinned host copies of the inputs and pinned host buffers for every output"""
        pin = lambda t: _pin(t.cpu() if t.is_cuda else t.clone())           # separate host buffers either way
        for e in self.engs:                                                  # the host-buffer mode deals the leaves to three engines: the device mode's single big
            if hasattr(e, "trim"): e.trim()                                  # workspace goes back first
        self.h = {k: pin(v) for k, v in data.items()}
        V, n = self.V, self.n
        hp = lambda *shape: _pin(torch.empty(shape, dtype=torch.uint8))
        self.h["packed"] = hp(V, self.packed_len + 32)
        self.h["linedom"] = hp(V, self.n_reads); self.h["linediv"] = hp(V, self.n_reads)
        self.h["seq_out"] = hp(V, self.n_seq); self.h["qual_out"] = hp(V, n)
        self.h["dec"] = {s: hp(V, self.name_len[s] + 16) for s in self.NAMES}
        if self.meta is not None:                                            # packed section buffers of the three pipelines, sized from the device pass
            for pl, names in (("qual", DQ), ("seq", ("NONREF_X",)), ("names", self.NAMES)):
                need = int(((self.meta.comp_len[:, [self.S_IDX[s] for s in names]] + 15) & ~15).sum())
                self.h["comp_" + pl] = hp(int(need * 1.05) + 65536)

    def _run_parts(self, parts):
        """one engine (host thread + stream) per independent pipeline; with a single engine they run one after the other"""
        if self.pool is None or len(self.engs) < len(parts):
            for p in parts:
                p(self.engs[0])
        else:
            futs = [self.pool.submit(p, self.engs[i]) for i, p in enumerate(parts)]
            errs = []
            for f in futs:
                try:
                    f.result()
                except Exception as ex:                      # collect every part before raising: no thread is left running
                    errs.append(ex)
            if errs:
                raise errs[0]
        self._kernel_ms(self.engs)                                         # (host-buffer mode: every engine took part; the device mode narrows it afterwards)

    def zip_host(self):
        """host buffers in, host buffers out; the DOMQ streams and the exception stream stay on the device between the complex codec
        and its sub-codec (GZB_OUT_DEVICE / GZB_SEC_IN_DEVICE) exactly as they stay inside one compute thread in the reference.
        The three independent pipelines of a FASTQ VBlock — QUAL (DOMQ + its four sub-streams), SEQ (ACGT + its exception
        stream) and the read-name contexts — run on one engine (host thread + stream) each, so the transfers of one
        overlap the entropy chains of another; QUAL's upload goes first because its chains are the longest.
        Returns (meta, h2d_bytes, d2h_bytes)."""
        L, V, n, H = self.L, self.V, self.n, self.h
        meta = ZipMeta(V, self.STREAMS)
        for s in self.NAMES:
            meta.len[:, self.S_IDX[s]] = self.name_len[s]
        dev_in = np.zeros(len(self.STREAMS), np.uint32)
        for s in DQ + ("NONREF_X",):
            dev_in[self.S_IDX[s]] = GZB_SEC_IN_DEVICE                 # intermediate streams stay in HBM until their sub-codec
        qual_up = threading.Event()
        passes_done = [threading.Event(), threading.Event()]        # the DOMQ passes / the ACGT pack of the whole group: chain kernels start after both (see zip_device)
        if self.pool is None or len(self.engs) < 3:                 # (one engine: the parts run one after the other, nothing to wait for)
            passes_done[0].set(); passes_done[1].set(); qual_up.set()
        name_rows = {s: self._rows(H[s]) for s in self.NAMES}

        def compress(eng, names, arena):
            inp = self._in_ptrs(meta, name_rows, self.dq_arena.data_ptr() if self.dq_arena is not None else 0)
            secs, a, vv, ss = self._section_array(meta, inp, names, dev_in)
            self._compress_packed(eng, secs, a, vv.size, 0, arena, True)
            meta.comp_len[vv, ss] = a["out_len"][:vv.size]; meta.comp_ptr[vv, ss] = a["out"][:vv.size]

        def part_qual(eng):
            try:
                first = min(V, self.SB)
                self._domq_device(lambda v: H["qual"][v].data_ptr(), eng, meta, GZB_OUT_DEVICE, 0, first)
            finally:
                qual_up.set()
            try:
                if first < V:
                    self._domq_device(lambda v: H["qual"][v].data_ptr(), eng, meta, GZB_OUT_DEVICE, first, V)
            finally:
                passes_done[0].set()
            passes_done[1].wait()
            compress(eng, DQ, "comp_qual")

        def part_seq(eng):
            qual_up.wait()
            a = self.avb_np
            try:
                a["seq"] = self._rows(H["seq"]); a["n_bases"] = self.n_seq; a["packed"] = self._rows(H["packed"]); a["x"] = self._rows(self.x_d)
                for v0 in range(0, V, 64):                          # bounded staging in the engine workspace
                    v1 = min(V, v0 + 64)
                    if L.gzb_acgt_pack_batch(eng.h, self._sub(self.avb, v0, v1), v1 - v0, GZB_OUT_DEVICE):
                        raise GzbError(f"gzb_acgt_pack_batch: {eng._err()}")
            finally:
                passes_done[1].set()
            meta.acgt_no_x[:] = a["x_all_zero"] != 0
            meta.len[:, self.S_IDX["NONREF_X"]] = np.where(meta.acgt_no_x, 0, self.n_seq)
            passes_done[0].wait()
            compress(eng, ("NONREF_X",), "comp_seq")

        def part_names(eng):
            passes_done[0].wait(); passes_done[1].wait()
            compress(eng, self.NAMES, "comp_names")

        self._run_parts([part_qual, part_seq, part_names])
        on_dev = [self.S_IDX[s] for s in DQ + ("NONREF_X",)]
        host_in = np.ones(len(self.STREAMS), bool); host_in[on_dev] = False
        h2d = V * (n + self.n_seq + 12 * self.n_reads) + int(meta.len[:, host_in].sum())
        d2h = V * (self.packed_len + 2 * self.n_reads) + int(meta.comp_len.sum())
        return meta, h2d, d2h

    def piz_host(self, meta):
        """the inverse of zip_host"""
        L, V, n, H = self.L, self.V, self.n, self.h
        dev_out = np.zeros(len(self.STREAMS), np.uint32)
        for s in DQ + ("NONREF_X",):
            dev_out[self.S_IDX[s]] = GZB_SEC_OUT_DEVICE
        outp = self._in_ptrs(meta, {s: self._rows(H["dec"][s]) for s in self.NAMES}, self.dq_arena.data_ptr())

        def uncompress(eng, names):
            secs, a, vv, ss = self._section_array(meta, meta.comp_ptr, names, dev_out)
            a["in_len"][:vv.size] = meta.comp_len[vv, ss]; a["out"][:vv.size] = outp[vv, ss]; a["out_cap"][:vv.size] = meta.len[vv, ss]
            if vv.size:
                eng.uncompress_raw(secs, vv.size, 0)

        self._fill_piz_descriptors(meta, self._rows(H["qual_out"]), self._rows(H["seq_out"]), self._rows(H["packed"]), self.line_len_h)

        def part_qual(eng):
            uncompress(eng, DQ)
            if L.gzb_domq_reconstruct(eng.h, self.pvb, V, GZB_IN_DEVICE):
                raise GzbError(f"gzb_domq_reconstruct: {eng._err()}")

        def part_seq(eng):
            uncompress(eng, ("NONREF_X",))
            for v0 in range(0, V, 64):
                v1 = min(V, v0 + 64)
                if L.gzb_acgt_unpack_batch(eng.h, self._sub(self.avb, v0, v1), v1 - v0, GZB_IN_DEVICE):
                    raise GzbError(f"gzb_acgt_unpack_batch: {eng._err()}")

        def part_names(eng):
            uncompress(eng, self.NAMES)

        self._run_parts([part_qual, part_seq, part_names])
        on_dev = [self.S_IDX[s] for s in DQ + ("NONREF_X",)]
        host_out = np.ones(len(self.STREAMS), bool); host_out[on_dev] = False
        h2d = V * (4 * self.n_reads + self.packed_len) + int(meta.comp_len.sum())
        d2h = V * (n + self.n_seq) + int(meta.len[:, host_out].sum())
        return h2d, d2h


UPLOADS, FETCHES, ALL = 0, 1, 2                                             # gzb_stage_wait


class PipelinedHost:
    """The host-buffer leg as a host that keeps handing over VBlock batches runs it (what genozip's dispatcher does with VBlocks): text
    and sections cross PCIe on the engine's copy streams (gzb_stage_upload / gzb_stage_fetch) while the PREVIOUS / NEXT batch's
    kernels run on device buffers (GZB_DEVICE_PTRS):

      zip   upload(k+1)  ||  [ACGT pack, DOMQ passes, entropy chains](k)   then fetch(2-bit words, packed sections)(k)
      piz   upload(2-bit words, sections)(k+1), fetch(SEQ, QUAL, names)(k-1)  ||  [entropy chains, DOMQ / ACGT reconstruct](k)

    The entropy chains are latency-bound and leave the copy engines idle; the transfers are bandwidth-bound and need no SM: two
    input buffers (zip) and two output buffers (piz) on the device are all it takes.  Host buffers are page-locked."""

    def __init__(self, path, data_host):
        """data_host: host or device tensors [V, ...] seq, qual, Q_* (copied into page-locked buffers; the read-name buffers get the 16 bytes of
        slack per row that alloc_piz gives their decoded counterparts, so that one pair of device buffers serves as zip's input
        slots and as piz's output slots)"""
        self.path = path
        self.eng, self.L = path.eng, path.L
        dev = path.dev
        self.width = {k: v.shape[1] for k, v in data_host.items()}
        self.H = {}
        for k, v in data_host.items():                                       # (device tensors are copied straight into the page-locked buffers: no pageable copy in between)
            t = _pin(torch.zeros((v.shape[0], v.shape[1] + (16 if k in path.NAMES else 0)), dtype=torch.uint8))
            t[:, :v.shape[1]].copy_(v)
            self.H[k] = t
        self.slots = [{k: torch.empty(v.shape, dtype=torch.uint8, device=dev) for k, v in self.H.items()} for _ in range(2)]
        self.h_packed = _pin(torch.empty((path.V, path.packed_len + 32), dtype=torch.uint8))
        self.h_comp = {}
        self.h_out = {k: _pin(torch.empty(v.shape, dtype=torch.uint8)) for k, v in self.H.items()}
        self.meta = None
        self.zbuf = self.pbuf = None                                          # double-buffered outputs of zip / 2-bit-word inputs of piz

    def _inputs(self, slot):
        return {k: t[:, :self.width[k]] for k, t in slot.items()}

    def _up(self, dst, src, nbytes=None):
        if self.L.gzb_stage_upload(self.eng.h, dst.data_ptr(), src.data_ptr(), dst.numel() * dst.element_size() if nbytes is None else nbytes):
            raise GzbError(f"gzb_stage_upload: {self.eng._err()}")

    def _down(self, dst, src, nbytes=None):
        if self.L.gzb_stage_fetch(self.eng.h, dst.data_ptr(), src.data_ptr(), src.numel() * src.element_size() if nbytes is None else nbytes):
            raise GzbError(f"gzb_stage_fetch: {self.eng._err()}")

    def _wait(self, which=ALL):
        if self.L.gzb_stage_wait(self.eng.h, which):
            raise GzbError(f"gzb_stage_wait: {self.eng._err()}")

    def zip_steps(self, K=1):
        """K zip passes over the host buffers, back to back; returns (h2d_bytes, d2h_bytes) of ONE step.  Step k's 2-bit words and packed
        sections are fetched under step k+1's kernels (two sets of output buffers on the device)."""
        p = self.path
        if self.zbuf is None:
            self.zbuf = [dict(packed_d=p.packed_d, comp_arena=p.comp_arena, comp_arena2=p.comp_arena2),
                         dict(packed_d=torch.empty_like(p.packed_d), comp_arena=None, comp_arena2=None)]
        for k_, t in self.H.items():
            self._up(self.slots[0][k_], t)
        h2d = sum(t.numel() for t in self.H.values())
        d2h = 0
        meta = None
        for k in range(K):
            self._wait(UPLOADS)                                              # step k's text is on the device
            if k + 1 < K:
                for k_, t in self.H.items():
                    self._up(self.slots[(k + 1) & 1][k_], t)                  # crosses PCIe under this step's kernels
            b = self.zbuf[k & 1]
            p.packed_d, p.comp_arena, p.comp_arena2 = b["packed_d"], b["comp_arena"], b["comp_arena2"]
            meta = p.zip_device(self._inputs(self.slots[k & 1]))
            b["comp_arena"], b["comp_arena2"] = p.comp_arena, p.comp_arena2  # (made or grown by the call)
            self._wait(FETCHES)                                              # step k-1's results reached the host under this step's kernels
            self._down(self.h_packed, p.packed_d)
            d2h = self.h_packed.numel()
            for a in ("comp_arena", "comp_arena2"):
                used = p.comp_used.get(a, 0)
                if used:
                    if a not in self.h_comp or self.h_comp[a].numel() < used:
                        self.h_comp[a] = _pin(torch.empty(int(used * 1.05) + 4096, dtype=torch.uint8))
                    self._down(self.h_comp[a], getattr(p, a), used)
                    d2h += used
        self._wait(ALL)
        self.meta = meta                                                     # (its section addresses are inside the buffers the path is left with)
        return h2d, d2h

    def _upload_comp(self):
        p, n = self.path, 0
        for a in ("comp_arena", "comp_arena2"):
            used = p.comp_used.get(a, 0)
            if used:
                self._up(getattr(p, a), self.h_comp[a], used); n += used
        return n

    def piz_steps(self, K=1):
        """K piz passes: the sections and 2-bit words go up, SEQ / QUAL / the read-name contexts come back into the host output buffers.
        Step k+1's 2-bit words go up, and step k-1's text comes back, under step k's kernels."""
        p, meta = self.path, self.meta
        if self.pbuf is None:
            other = [b["packed_d"] for b in (self.zbuf or []) if b["packed_d"] is not p.packed_d]
            self.pbuf = [p.packed_d, other[0] if other else torch.empty_like(p.packed_d)]
        self._up(self.pbuf[0], self.h_packed)
        h2d = self.h_packed.numel() + self._upload_comp()
        d2h = 0
        for k in range(K):
            self._wait(UPLOADS)                                              # step k's sections and 2-bit words are on the device
            if k + 1 < K:
                self._up(self.pbuf[(k + 1) & 1], self.h_packed)              # crosses PCIe under this step's chains
            outs = self.slots[k & 1]
            p.packed_d = self.pbuf[k & 1]
            p.piz_device(meta, outs)
            self._wait(FETCHES)                                              # step k-1's text reached the host under this step's chains
            if k + 1 < K:
                self._upload_comp()                                          # (the kernels of step k are done with the section buffers; a few hundred MB)
            for k_, t in outs.items():
                self._down(self.h_out[k_], t)                                 # crosses PCIe under the next step's chains
            d2h = sum(t.numel() for t in outs.values())
        self._wait(ALL)
        return h2d, d2h

    def check(self):
        return all(torch.equal(self.h_out[k][:, :w], self.H[k][:, :w]) for k, w in self.width.items())

    def scrub(self):
        """empty the host output buffers (a round-trip check then sees what piz_steps produced)"""
        for t in self.h_out.values():
            t.zero_()
