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

from .lib import (load, Engine, Section, DomqVb, DomqPizVb, AcgtVb, CODEC, est_size, GzbError,
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


class FastqCodecPath:
    """zip / piz of a batch of V FASTQ VBlocks through libgzb200 on one GPU.

    The path owns `n_engines` engines (one host thread + CUDA stream + workspace each, the library's unit of concurrency:
    "one engine per host thread and GPU").  The host-buffer path gives each of the three independent pipelines of a FASTQ
    VBlock (QUAL, SEQ, read names) its own engine so that transfers overlap the entropy chains (zip_host / piz_host).
    The device-resident path can deal the batch to `device_groups` engines in contiguous groups of VBlocks, as genozip's
    dispatcher hands VBlocks to compute threads; measured on B200 this does not help (the chain kernels' duration is
    set by their longest leaf, not by the batch size), so the default is one group."""

    def __init__(self, eng: Engine, V, n_reads, read_len, n_engines=3, device_groups=1):
        self.eng, self.L = eng, eng.L
        self.V, self.n_reads, self.read_len = V, n_reads, read_len
        self.n = n_reads * read_len
        dev = torch.device(getattr(eng, "torch_device", None) or f"cuda:{eng.device}")   # (the CPU test suite drives this class through a mock engine)
        self.dev = dev
        n_engines = max(1, n_engines)
        K = max(1, min(n_engines, V, device_groups))
        self.engs = [eng] + [type(eng)(eng.device) for _ in range(n_engines - 1)]
        self.groups = [(g * V // K, (g + 1) * V // K) for g in range(K)]
        self.pool = ThreadPoolExecutor(n_engines) if n_engines > 1 else None
        self.stream = torch.cuda.ExternalStream(self.L.gzb_engine_stream(eng.h), device=dev) if dev.type == "cuda" else None
        n, V = self.n, V
        self.packed_len = int(self.L.gzb_acgt_packed_len(n))
        u8 = dict(dtype=torch.uint8, device=dev)
        self.line_off_h = _pin(torch.arange(n_reads, dtype=torch.int64) * read_len)
        self.line_len_h = _pin(torch.full((n_reads,), read_len, dtype=torch.int32))
        self.line_off_d, self.line_len_d = self.line_off_h.to(dev), self.line_len_h.to(dev)
        # device intermediates / outputs (zip)
        self.packed_d = torch.empty((V, self.packed_len + 32), **u8)
        self.x_d = torch.empty((V, n), **u8)
        self.linedom_d = torch.empty((V, n_reads), **u8)
        self.linediv_d = torch.empty((V, n_reads), **u8)
        self.dq = {k: torch.empty((V, c), **u8) for k, c in (("QUAL", 2 * n + 16), ("DOMQRUNS", n + 16), ("QUALMPLX", n_reads + 16), ("DIVRQUAL", n + 16))}
        self.caps = {"QUAL": 2 * n + 16, "DOMQRUNS": n + 16, "QUALMPLX": n_reads + 16, "DIVRQUAL": n + 16, "NONREF_X": n,
                     "Q_TILE": n_reads, "Q_X": 4 * n_reads, "Q_Y": 4 * n_reads, "Q_MISC": n_reads}
        self.codec = {s: "RANB" for s in STREAMS}
        self.comp_d = {}          # compressed sections on device: stream -> [V, est]
        self.dvb = (DomqVb * V)()
        self.pvb = (DomqPizVb * V)()
        self.avb = (AcgtVb * V)()
        self.meta = None          # per-VB dicts from the last zip (lengths, tables)
        self.h = {}               # pinned host buffers for the host-buffer (e2e) path
        self.kernel_ms = (0.0, 0.0)   # chain kernel durations of the last call: (rANS, arithmetic), mean over the engines

    @property
    def launches(self):
        return sum(e.launches for e in self.engs)

    def close(self):
        """release the engines this path created (not the caller's) and its host threads"""
        if self.pool is not None:
            self.pool.shutdown(wait=True); self.pool = None
        for e in self.engs[1:]:
            e.close()
        self.engs = self.engs[:1]

    def _each_group(self, fn):
        """run fn(g, engine, v0, v1) for every group, one host thread per engine; results in group order"""
        jobs = [(g, self.engs[g], v0, v1) for g, (v0, v1) in enumerate(self.groups)]
        if self.pool is None:
            res = [fn(*j) for j in jobs]
        else:
            res = list(self.pool.map(lambda j: fn(*j), jobs))
        self.kernel_ms = tuple(float(np.mean([self.L.gzb_last_kernel_ms(e.h, w) for e in self.engs[:len(self.groups)]])) for w in (0, 1))
        return res

    @staticmethod
    def _sub(arr, v0, v1):
        """ctypes view of elements [v0, v1) of a ctypes array (shares memory)"""
        return (arr._type_ * (v1 - v0)).from_buffer(arr, v0 * C.sizeof(arr._type_))

    # ------------------------------------------------------------------ codec assignment (host policy, run on the GPU)
    def assign_codecs(self, data):
        """codec_assign_best_codec's size criterion (src/codec.c:234-389, sorter :128-173) restricted to the eight
        in-scope simple codecs: compress the first <=99,999 bytes (CODEC_ASSIGN_SAMPLE_SIZE, src/codec.h:154) of VB 1's
        stream with each and keep the smallest, ties to the lower Codec value.  The reference also weighs clock() time
        (timing-dependent, H5) — not reproduced.  Samples are compressed on the GPU (same bytes as the reference)."""
        self.zip_device(data, only_vb0_streams=True)
        m = self.meta[0]
        samples = {}
        for s in STREAMS:
            ln = m["len"][s]
            if ln == 0:
                continue
            src = self._stream_dev_tensor(s, 0, data)[:min(ln, 99999)]
            samples[s] = src.cpu().numpy().copy()
        items = [(c, samples[s]) for s in samples for c in SIMPLE]
        outs = self.eng.compress(items)
        k = 0
        for s in samples:
            sizes = [outs[k + j].size for j in range(len(SIMPLE))]
            k += len(SIMPLE)
            self.codec[s] = SIMPLE[int(np.argmin(sizes))] if samples[s].size >= 50 else "RANB"   # <50 B would be CODEC_NONE (compressor.c:56-58)
        return dict(self.codec)

    def _stream_dev_tensor(self, s, v, data):
        if s in self.dq:
            return self.dq[s][v]
        if s == "NONREF_X":
            return self.x_d[v]
        return data[s][v]

    def _stream_cap(self, s, meta):
        """capacity for stream s: 25% above the longest instance in this batch (the data of a step does not change)"""
        longest = max(m["len"][s] for m in meta)
        return min(self.caps[s], int(longest * 1.25) + 4096)

    def _alloc_comp(self, meta):
        for s in STREAMS:
            cap = est_size(self.codec[s], self._stream_cap(s, meta))
            if s not in self.comp_d or self.comp_d[s].shape[1] < cap:
                self.comp_d[s] = torch.empty((self.V, cap), dtype=torch.uint8, device=self.dev)

    # ------------------------------------------------------------------ ZIP, inputs resident in HBM
    def zip_device(self, data, only_vb0_streams=False):
        L, V, n = self.L, self.V, self.n
        meta = [dict(len={}, comp_len={}) for _ in range(V)]
        for v in range(V):
            b = self.avb[v]
            b.seq = data["seq"][v].data_ptr(); b.n_bases = n; b.packed = self.packed_d[v].data_ptr(); b.x = self.x_d[v].data_ptr()
            a = self.dvb[v]
            a.txt = data["qual"][v].data_ptr(); a.txt_len = n
            a.line_off = self.line_off_d.data_ptr(); a.line_len = self.line_len_d.data_ptr(); a.n_lines = self.n_reads
            a.line_dom = self.linedom_d[v].data_ptr(); a.line_diverse = self.linediv_d[v].data_ptr()
            for fld, s in (("qual", "QUAL"), ("runs", "DOMQRUNS"), ("mplx", "QUALMPLX"), ("divr", "DIVRQUAL")):
                setattr(a, fld, self.dq[s][v].data_ptr()); setattr(a, fld + "_cap", self.caps[s])

        def domain(g, eng, v0, v1):
            h = eng.h
            if L.gzb_acgt_pack_batch(h, self._sub(self.avb, v0, v1), v1 - v0, GZB_DEVICE_PTRS):
                raise GzbError(f"gzb_acgt_pack_batch: {eng._err()}")
            dv = self._sub(self.dvb, v0, v1)
            if L.gzb_domq_prepare(h, dv, v1 - v0, GZB_DEVICE_PTRS) or L.gzb_domq_split(h, dv, v1 - v0, GZB_DEVICE_PTRS):
                raise GzbError(f"gzb_domq: {eng._err()}")
            for v in range(v0, v1):
                a, m = self.dvb[v], meta[v]
                m["acgt_no_x"] = bool(self.avb[v].x_all_zero)
                m["len"].update(NONREF_X=0 if self.avb[v].x_all_zero else n, QUAL=a.qual_len, DOMQRUNS=a.runs_len, QUALMPLX=a.mplx_len,
                                DIVRQUAL=a.divr_len, Q_TILE=self.n_reads, Q_X=4 * self.n_reads, Q_Y=4 * self.n_reads, Q_MISC=self.n_reads)
                m["num_norm_qs"] = a.num_norm_qs
                m["denorm"] = bytes(a.denorm)[:a.num_norm_qs * a.num_doms]

        def sections(g, eng, v0, v1):
            secs, idx = self._sections(meta, lambda s, v: self._stream_dev_tensor(s, v, data).data_ptr(), lambda s, v: self.comp_d[s][v].data_ptr(), 0, v0, v1)
            eng.compress_raw(secs, len(idx), GZB_DEVICE_PTRS)
            self._collect(secs, idx, meta)

        if only_vb0_streams or not self.comp_d:
            # first call of a batch: the compressed-section buffers are sized from the streams' actual lengths
            self._each_group(domain)
            self.meta = meta
            if only_vb0_streams:
                return meta
            self._alloc_comp(meta)
            self._each_group(sections)
        else:
            self._each_group(lambda g, eng, v0, v1: (domain(g, eng, v0, v1), sections(g, eng, v0, v1)))
            self.meta = meta
        return meta

    def _sections(self, meta, in_ptr, out_ptr, sflags_for, v0=0, v1=None):
        idx = [(v, s) for v in range(v0, self.V if v1 is None else v1) for s in STREAMS if meta[v]["len"][s] > 0]
        secs = (Section * len(idx))()
        for i, (v, s) in enumerate(idx):
            secs[i].codec = CODEC[self.codec[s]]
            secs[i].in_ = in_ptr(s, v); secs[i].in_len = meta[v]["len"][s]
            secs[i].out = out_ptr(s, v); secs[i].out_cap = est_size(self.codec[s], meta[v]["len"][s])
            secs[i].sflags = sflags_for(s) if callable(sflags_for) else sflags_for
        return secs, idx

    @staticmethod
    def _collect(secs, idx, meta):
        for i, (v, s) in enumerate(idx):
            if secs[i].status != 0:
                raise GzbError(f"section {s} of VB {v}: status {secs[i].status}")
            meta[v]["comp_len"][s] = secs[i].out_len

    # ------------------------------------------------------------------ PIZ, inputs resident in HBM
    def alloc_piz(self, meta):
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.dec_d = {s: torch.empty((self.V, self._stream_cap(s, meta) + 16), **u8) for s in STREAMS}
        self.seq_out_d = torch.empty((self.V, self.n), **u8)
        self.qual_out_d = torch.empty((self.V, self.n), **u8)

    def piz_device(self, meta):
        L, n = self.L, self.n
        keep = []

        def group(g, eng, v0, v1):
            h = eng.h
            idx = [(v, s) for v in range(v0, v1) for s in STREAMS if meta[v]["len"][s] > 0]
            secs = (Section * len(idx))()
            for i, (v, s) in enumerate(idx):
                secs[i].codec = CODEC[self.codec[s]]
                secs[i].in_ = self.comp_d[s][v].data_ptr(); secs[i].in_len = meta[v]["comp_len"][s]
                secs[i].out = self.dec_d[s][v].data_ptr(); secs[i].out_cap = meta[v]["len"][s]
            eng.uncompress_raw(secs, len(idx), GZB_DEVICE_PTRS)
            for v in range(v0, v1):
                a, m = self.pvb[v], meta[v]
                a.qual = self.dec_d["QUAL"][v].data_ptr(); a.qual_len = m["len"]["QUAL"]
                a.runs = self.dec_d["DOMQRUNS"][v].data_ptr(); a.runs_len = m["len"]["DOMQRUNS"]
                a.mplx = self.dec_d["QUALMPLX"][v].data_ptr(); a.mplx_len = m["len"]["QUALMPLX"]
                a.divr = self.dec_d["DIVRQUAL"][v].data_ptr(); a.divr_len = m["len"]["DIVRQUAL"]
                dn = np.frombuffer(m["denorm"], np.uint8); keep.append(dn)
                a.denorm = dn.ctypes.data; a.denorm_len = dn.size; a.num_norm_qs = m["num_norm_qs"]
                a.line_len = self.line_len_d.data_ptr(); a.n_lines = self.n_reads
                a.out = self.qual_out_d[v].data_ptr(); a.out_cap = n
                b = self.avb[v]
                b.seq = self.seq_out_d[v].data_ptr(); b.n_bases = n; b.packed = self.packed_d[v].data_ptr()
                b.x = None if meta[v]["acgt_no_x"] else self.dec_d["NONREF_X"][v].data_ptr()
            if L.gzb_domq_reconstruct(h, self._sub(self.pvb, v0, v1), v1 - v0, GZB_DEVICE_PTRS):
                raise GzbError(f"gzb_domq_reconstruct: {eng._err()}")
            if L.gzb_acgt_unpack_batch(h, self._sub(self.avb, v0, v1), v1 - v0, GZB_DEVICE_PTRS):
                raise GzbError(f"gzb_acgt_unpack_batch: {eng._err()}")

        self._each_group(group)

    # ------------------------------------------------------------------ HOST-buffer path (e2e): what the C host would call
    def alloc_host(self, data):
        """pinned host copies of the inputs and pinned host buffers for every output"""
        pin = lambda t: _pin(t.cpu() if t.is_cuda else t.clone())           # separate host buffers either way
        self.h = {k: pin(v) for k, v in data.items()}
        V, n = self.V, self.n
        hp = lambda *shape: _pin(torch.empty(shape, dtype=torch.uint8))
        self.h["packed"] = hp(V, self.packed_len + 32)
        self.h["linedom"] = hp(V, self.n_reads); self.h["linediv"] = hp(V, self.n_reads)
        # compressed-section buffers: est_size of the largest actual stream of each kind (the capacity the C-ABI requires)
        def comp_cap(s):
            lens = [m["len"][s] for m in (self.meta or [])]
            return max([est_size(self.codec[s], l) for l in lens] + [4096])
        self.h["comp"] = {s: hp(V, comp_cap(s)) for s in STREAMS}
        self.h["seq_out"] = hp(V, n); self.h["qual_out"] = hp(V, n)
        self.h["dec"] = {s: hp(V, self.dec_d[s].shape[1]) for s in ("Q_TILE", "Q_X", "Q_Y", "Q_MISC")}

    def zip_host(self):
        """host buffers in, host buffers out; the DOMQ streams stay on the device between codec_domq_compress and
        its sub-codec (GZB_OUT_DEVICE / GZB_SEC_IN_DEVICE) exactly as they stay inside one compute thread in the reference.
        The three independent pipelines of a FASTQ VBlock — QUAL (DOMQ + its four sub-streams), SEQ (ACGT + its exception
        stream) and the read-name contexts — run on one engine (host thread + stream) each, so the transfers of one
        overlap the entropy chains of another; QUAL's upload goes first because its chains are the longest.
        Returns (meta, h2d_bytes, d2h_bytes)."""
        L, V, n, H = self.L, self.V, self.n, self.h
        meta = [dict(len={}, comp_len={}) for _ in range(V)]
        for v in range(V):
            b = self.avb[v]
            b.seq = H["seq"][v].data_ptr(); b.n_bases = n; b.packed = H["packed"][v].data_ptr(); b.x = self.x_d[v].data_ptr()
            a = self.dvb[v]
            a.txt = H["qual"][v].data_ptr(); a.txt_len = n
            a.line_off = self.line_off_h.data_ptr(); a.line_len = self.line_len_h.data_ptr(); a.n_lines = self.n_reads
            a.line_dom = H["linedom"][v].data_ptr(); a.line_diverse = H["linediv"][v].data_ptr()
            for fld, s in (("qual", "QUAL"), ("runs", "DOMQRUNS"), ("mplx", "QUALMPLX"), ("divr", "DIVRQUAL")):
                setattr(a, fld, self.dq[s][v].data_ptr()); setattr(a, fld + "_cap", self.caps[s])
        on_dev = set(self.dq) | {"NONREF_X"}                # intermediate streams stay in HBM until their sub-codec
        qual_up = threading.Event()

        def in_ptr(s, v):
            if s in self.dq: return self.dq[s][v].data_ptr()
            if s == "NONREF_X": return self.x_d[v].data_ptr()
            return H[s][v].data_ptr()

        def compress(eng, names):
            idx = [(v, s) for v in range(V) for s in names if meta[v]["len"][s] > 0]
            secs = (Section * len(idx))()
            for i, (v, s) in enumerate(idx):
                secs[i].codec = CODEC[self.codec[s]]
                secs[i].in_ = in_ptr(s, v); secs[i].in_len = meta[v]["len"][s]
                secs[i].out = H["comp"][s][v].data_ptr(); secs[i].out_cap = est_size(self.codec[s], meta[v]["len"][s])
                secs[i].sflags = GZB_SEC_IN_DEVICE if s in on_dev else 0
            eng.compress_raw(secs, len(idx), 0)
            self._collect(secs, idx, meta)

        def part_qual(eng):
            try:
                if L.gzb_domq_prepare(eng.h, self.dvb, V, GZB_OUT_DEVICE):
                    raise GzbError(f"gzb_domq_prepare: {eng._err()}")
            finally:
                qual_up.set()
            if L.gzb_domq_split(eng.h, self.dvb, V, GZB_OUT_DEVICE):
                raise GzbError(f"gzb_domq_split: {eng._err()}")
            for v in range(V):
                a, m = self.dvb[v], meta[v]
                m["len"].update(QUAL=a.qual_len, DOMQRUNS=a.runs_len, QUALMPLX=a.mplx_len, DIVRQUAL=a.divr_len)
                m["num_norm_qs"] = a.num_norm_qs
                m["denorm"] = bytes(a.denorm)[:a.num_norm_qs * a.num_doms]
            compress(eng, ("QUAL", "DOMQRUNS", "QUALMPLX", "DIVRQUAL"))

        def part_seq(eng):
            qual_up.wait()
            for v0 in range(0, V, 64):                          # bounded staging in the engine workspace
                if L.gzb_acgt_pack_batch(eng.h, self._sub(self.avb, v0, min(V, v0 + 64)), min(V, v0 + 64) - v0, GZB_OUT_DEVICE):
                    raise GzbError(f"gzb_acgt_pack_batch: {eng._err()}")
            for v in range(V):
                meta[v]["acgt_no_x"] = bool(self.avb[v].x_all_zero)
                meta[v]["len"]["NONREF_X"] = 0 if self.avb[v].x_all_zero else n
            compress(eng, ("NONREF_X",))

        def part_names(eng):
            for v in range(V):
                meta[v]["len"].update(Q_TILE=self.n_reads, Q_X=4 * self.n_reads, Q_Y=4 * self.n_reads, Q_MISC=self.n_reads)
            qual_up.wait()
            compress(eng, ("Q_TILE", "Q_X", "Q_Y", "Q_MISC"))

        self._run_parts([part_qual, part_seq, part_names])
        h2d = V * (n + n + 12 * self.n_reads); d2h = V * (self.packed_len + 2 * self.n_reads)
        for m in meta:
            for s, ln in m["len"].items():
                if ln:
                    if s not in on_dev: h2d += ln
                    d2h += m["comp_len"][s]
        return meta, h2d, d2h

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
        self.kernel_ms = tuple(float(np.max([self.L.gzb_last_kernel_ms(e.h, w) for e in self.engs])) for w in (0, 1))

    def piz_host(self, meta):
        L, V, n, H = self.L, self.V, self.n, self.h
        keep = []

        def uncompress(eng, names):
            idx = [(v, s) for v in range(V) for s in names if meta[v]["len"][s] > 0]
            secs = (Section * len(idx))()
            for i, (v, s) in enumerate(idx):
                on_dev = s in self.dq or s == "NONREF_X"
                secs[i].codec = CODEC[self.codec[s]]
                secs[i].in_ = H["comp"][s][v].data_ptr(); secs[i].in_len = meta[v]["comp_len"][s]
                secs[i].out = (self.dec_d[s][v] if on_dev else H["dec"][s][v]).data_ptr(); secs[i].out_cap = meta[v]["len"][s]
                secs[i].sflags = GZB_SEC_OUT_DEVICE if on_dev else 0
            eng.uncompress_raw(secs, len(idx), 0)

        def part_qual(eng):
            uncompress(eng, ("QUAL", "DOMQRUNS", "QUALMPLX", "DIVRQUAL"))
            for v in range(V):
                a, m = self.pvb[v], meta[v]
                a.qual = self.dec_d["QUAL"][v].data_ptr(); a.qual_len = m["len"]["QUAL"]
                a.runs = self.dec_d["DOMQRUNS"][v].data_ptr(); a.runs_len = m["len"]["DOMQRUNS"]
                a.mplx = self.dec_d["QUALMPLX"][v].data_ptr(); a.mplx_len = m["len"]["QUALMPLX"]
                a.divr = self.dec_d["DIVRQUAL"][v].data_ptr(); a.divr_len = m["len"]["DIVRQUAL"]
                dn = np.frombuffer(m["denorm"], np.uint8); keep.append(dn)
                a.denorm = dn.ctypes.data; a.denorm_len = dn.size; a.num_norm_qs = m["num_norm_qs"]
                a.line_len = self.line_len_h.data_ptr(); a.n_lines = self.n_reads
                a.out = H["qual_out"][v].data_ptr(); a.out_cap = n
            if L.gzb_domq_reconstruct(eng.h, self.pvb, V, GZB_IN_DEVICE):
                raise GzbError(f"gzb_domq_reconstruct: {eng._err()}")

        def part_seq(eng):
            uncompress(eng, ("NONREF_X",))
            for v in range(V):
                b = self.avb[v]
                b.seq = H["seq_out"][v].data_ptr(); b.n_bases = n; b.packed = H["packed"][v].data_ptr()
                b.x = None if meta[v]["acgt_no_x"] else self.dec_d["NONREF_X"][v].data_ptr()
            for v0 in range(0, V, 64):
                if L.gzb_acgt_unpack_batch(eng.h, self._sub(self.avb, v0, min(V, v0 + 64)), min(V, v0 + 64) - v0, GZB_IN_DEVICE):
                    raise GzbError(f"gzb_acgt_unpack_batch: {eng._err()}")

        def part_names(eng):
            uncompress(eng, ("Q_TILE", "Q_X", "Q_Y", "Q_MISC"))

        self._run_parts([part_qual, part_seq, part_names])
        on_dev = set(self.dq) | {"NONREF_X"}
        h2d = V * (4 * self.n_reads + self.packed_len); d2h = V * (n + n)
        for m in meta:
            for s, ln in m["len"].items():
                if ln:
                    h2d += m["comp_len"][s]
                    if s not in on_dev: d2h += ln
        return h2d, d2h
