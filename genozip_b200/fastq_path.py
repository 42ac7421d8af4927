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


SEC_DT, DVB_DT, PVB_DT, AVB_DT = (np.dtype(t) for t in (Section, DomqVb, DomqPizVb, AcgtVb))   # numpy views of the C-ABI descriptor structs


def _pin(t):
    """pinned host memory where there is a GPU to transfer to"""
    return t.pin_memory() if torch.cuda.is_available() else t


class FastqCodecPath:
    """zip / piz of a batch of V FASTQ VBlocks through libgzb200 on one GPU.

    The path owns `n_engines` engines (one host thread + CUDA stream + workspace each, the library's unit of concurrency:
    "one engine per host thread and GPU").  The host-buffer mode gives each of the three independent pipelines of a FASTQ
    VBlock (QUAL, SEQ, read names) its own engine so that transfers overlap the entropy chains (zip_host / piz_host); the
    device-resident mode uses one engine (dealing the batch to several engines, by VBlock groups or by pipeline, was
    measured on B200 and does not help: the chain kernels' duration is set by their longest leaf, not by the batch size)."""

    def __init__(self, eng: Engine, V, n_reads, read_len, n_engines=3):
        self.eng, self.L = eng, eng.L
        self.V, self.n_reads, self.read_len = V, n_reads, read_len
        self.n = n_reads * read_len
        dev = torch.device(getattr(eng, "torch_device", None) or f"cuda:{eng.device}")   # (tests drive this class on the CPU through a mock engine)
        self.dev = dev
        n_engines = max(1, n_engines)
        self.engs = [eng] + [type(eng)(eng.device) for _ in range(n_engines - 1)]
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
        self.kernel_ms = (0.0, 0.0)   # chain kernel durations of the last call: (rANS, arithmetic), longest over the engines used

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

    # ------------------------------------------------------------------ descriptor arrays, built column-wise with numpy
    # (a batch has 512 VBlocks x 9 streams: per-element ctypes attribute assignments would cost tens of milliseconds per step)
    def _rows(self, t):
        """addresses of the rows of a 2-D uint8 tensor [V, cap]"""
        assert t.dim() == 2 and t.shape[0] == self.V and t.stride(1) == 1
        return np.uint64(t.data_ptr()) + np.arange(self.V, dtype=np.uint64) * np.uint64(t.stride(0))

    def _section_array(self, names, in_rows, in_len, out_rows, out_cap, sflags):
        """Section descriptors of the streams `names` for every VBlock with a non-empty stream, VBlock-major.
        in_rows / out_rows: s -> (V,) addresses; in_len: s -> (V,) lengths; out_cap: s -> scalar or (V,); sflags: s -> flags.
        Returns (descriptor array, mask[V, len(names)] of the sections present)."""
        a = np.zeros((self.V, len(names)), SEC_DT)
        for j, s in enumerate(names):
            a["codec"][:, j] = CODEC[self.codec[s]]
            a["in_"][:, j] = in_rows[s]; a["in_len"][:, j] = in_len[s]
            a["out"][:, j] = out_rows[s]; a["out_cap"][:, j] = out_cap[s]
            a["sflags"][:, j] = sflags.get(s, 0)
        mask = a["in_len"] > 0
        return np.ascontiguousarray(a[mask]), mask

    @staticmethod
    def _run_sections(call, secs, flags):
        if secs.size:
            call(secs.ctypes.data_as(C.POINTER(Section)), int(secs.size), flags)
            bad = np.flatnonzero(secs["status"] != 0)
            if bad.size:
                raise GzbError(f"section {int(bad[0])} of the batch: status {int(secs['status'][bad[0]])}")

    def _zip_descriptors(self, seq_rows, packed_rows, qual_rows, line_off, line_len, dom_rows, div_rows):
        n = self.n
        av = np.frombuffer(self.avb, dtype=AVB_DT)
        av["seq"] = seq_rows; av["n_bases"] = n; av["packed"] = packed_rows; av["x"] = self._rows(self.x_d)
        dv = np.frombuffer(self.dvb, dtype=DVB_DT)
        dv["txt"] = qual_rows; dv["txt_len"] = n
        dv["line_off"] = line_off.data_ptr(); dv["line_len"] = line_len.data_ptr(); dv["n_lines"] = self.n_reads
        dv["line_dom"] = dom_rows; dv["line_diverse"] = div_rows
        for fld, s in (("qual", "QUAL"), ("runs", "DOMQRUNS"), ("mplx", "QUALMPLX"), ("divr", "DIVRQUAL")):
            dv[fld] = self._rows(self.dq[s]); dv[fld + "_cap"] = self.caps[s]
        return av, dv

    def _stream_lengths(self, av, dv):
        """lengths of the nine streams of every VBlock after ACGT pack and DOMQ split"""
        V, nr = self.V, self.n_reads
        full = lambda k: np.full(V, k, np.uint32)
        return {"QUAL": dv["qual_len"].copy(), "DOMQRUNS": dv["runs_len"].copy(), "QUALMPLX": dv["mplx_len"].copy(), "DIVRQUAL": dv["divr_len"].copy(),
                "NONREF_X": np.where(av["x_all_zero"] != 0, 0, self.n).astype(np.uint32),
                "Q_TILE": full(nr), "Q_X": full(4 * nr), "Q_Y": full(4 * nr), "Q_MISC": full(nr)}

    def _make_meta(self, av, dv, lens):
        nq, nd, dn = dv["num_norm_qs"], dv["num_doms"], dv["denorm"]
        return [dict(len={s: int(lens[s][v]) for s in STREAMS}, comp_len={}, acgt_no_x=bool(av["x_all_zero"][v]),
                     num_norm_qs=int(nq[v]), denorm=dn[v, :int(nq[v]) * int(nd[v])].tobytes()) for v in range(self.V)]

    @staticmethod
    def _store_comp_lens(meta, names, secs, mask):
        vs, js = np.nonzero(mask)
        for v, j, ln in zip(vs.tolist(), js.tolist(), secs["out_len"].tolist()):
            meta[v]["comp_len"][names[j]] = ln

    # ------------------------------------------------------------------ ZIP, inputs resident in HBM
    def zip_device(self, data, only_vb0_streams=False):
        L, V, eng = self.L, self.V, self.eng
        av, dv = self._zip_descriptors(self._rows(data["seq"]), self._rows(self.packed_d), self._rows(data["qual"]),
                                       self.line_off_d, self.line_len_d, self._rows(self.linedom_d), self._rows(self.linediv_d))
        if L.gzb_acgt_pack_batch(eng.h, self.avb, V, GZB_DEVICE_PTRS):
            raise GzbError(f"gzb_acgt_pack_batch: {eng._err()}")
        if L.gzb_domq_prepare(eng.h, self.dvb, V, GZB_DEVICE_PTRS) or L.gzb_domq_split(eng.h, self.dvb, V, GZB_DEVICE_PTRS):
            raise GzbError(f"gzb_domq: {eng._err()}")
        lens = self._stream_lengths(av, dv)
        meta = self.meta = self._make_meta(av, dv, lens)
        if only_vb0_streams:
            return meta
        self._alloc_comp(meta)                              # (sized from the first batch's streams; a no-op afterwards)
        in_rows = {s: self._rows(self.dq[s]) for s in self.dq}
        in_rows["NONREF_X"] = self._rows(self.x_d)
        in_rows.update({s: self._rows(data[s]) for s in ("Q_TILE", "Q_X", "Q_Y", "Q_MISC")})
        secs, mask = self._section_array(STREAMS, in_rows, lens, {s: self._rows(self.comp_d[s]) for s in STREAMS},
                                         {s: self.comp_d[s].shape[1] for s in STREAMS}, {})
        self._run_sections(eng.compress_raw, secs, GZB_DEVICE_PTRS)
        self._store_comp_lens(meta, STREAMS, secs, mask)
        self.kernel_ms = (float(L.gzb_last_kernel_ms(eng.h, 0)), float(L.gzb_last_kernel_ms(eng.h, 1)))
        return meta

    # ------------------------------------------------------------------ PIZ, inputs resident in HBM
    def alloc_piz(self, meta):
        u8 = dict(dtype=torch.uint8, device=self.dev)
        self.dec_d = {s: torch.empty((self.V, self._stream_cap(s, meta) + 16), **u8) for s in STREAMS}
        self.seq_out_d = torch.empty((self.V, self.n), **u8)
        self.qual_out_d = torch.empty((self.V, self.n), **u8)

    def _meta_arrays(self, meta):
        lens = {s: np.fromiter((m["len"][s] for m in meta), np.uint32, self.V) for s in STREAMS}
        clens = {s: np.fromiter((m["comp_len"].get(s, 0) for m in meta), np.uint32, self.V) for s in STREAMS}
        return lens, clens

    def _piz_descriptors(self, meta, lens, line_len, qual_out_rows, seq_out_rows, packed_rows):
        """DOMQ / ACGT reconstruct descriptors; returns the objects that must stay alive during the calls"""
        V = self.V
        pv = np.frombuffer(self.pvb, dtype=PVB_DT)
        for fld, s in (("qual", "QUAL"), ("runs", "DOMQRUNS"), ("mplx", "QUALMPLX"), ("divr", "DIVRQUAL")):
            pv[fld] = self._rows(self.dec_d[s]); pv[fld + "_len"] = lens[s]
        width = max([len(m["denorm"]) for m in meta] + [1])
        dn = np.zeros((V, width), np.uint8)                  # denormalisation tables (host memory), one row per VBlock
        for v, m in enumerate(meta):
            dn[v, :len(m["denorm"])] = np.frombuffer(m["denorm"], np.uint8)
        pv["denorm"] = np.uint64(dn.ctypes.data) + np.arange(V, dtype=np.uint64) * np.uint64(width)
        pv["denorm_len"] = np.fromiter((len(m["denorm"]) for m in meta), np.uint32, V)
        pv["num_norm_qs"] = np.fromiter((m["num_norm_qs"] for m in meta), np.uint8, V)
        pv["line_len"] = line_len.data_ptr(); pv["n_lines"] = self.n_reads
        pv["out"] = qual_out_rows; pv["out_cap"] = self.n
        av = np.frombuffer(self.avb, dtype=AVB_DT)
        av["seq"] = seq_out_rows; av["n_bases"] = self.n; av["packed"] = packed_rows
        no_x = np.fromiter((m["acgt_no_x"] for m in meta), bool, V)
        av["x"] = np.where(no_x, np.uint64(0), self._rows(self.dec_d["NONREF_X"]))
        return dn

    def piz_device(self, meta):
        L, V, eng = self.L, self.V, self.eng
        lens, clens = self._meta_arrays(meta)
        secs, _ = self._section_array(STREAMS, {s: self._rows(self.comp_d[s]) for s in STREAMS}, clens,
                                      {s: self._rows(self.dec_d[s]) for s in STREAMS}, lens, {})
        self._run_sections(eng.uncompress_raw, secs, GZB_DEVICE_PTRS)
        self.kernel_ms = (float(L.gzb_last_kernel_ms(eng.h, 0)), float(L.gzb_last_kernel_ms(eng.h, 1)))
        keep = self._piz_descriptors(meta, lens, self.line_len_d, self._rows(self.qual_out_d), self._rows(self.seq_out_d), self._rows(self.packed_d))
        if L.gzb_domq_reconstruct(eng.h, self.pvb, V, GZB_DEVICE_PTRS):
            raise GzbError(f"gzb_domq_reconstruct: {eng._err()}")
        if L.gzb_acgt_unpack_batch(eng.h, self.avb, V, GZB_DEVICE_PTRS):
            raise GzbError(f"gzb_acgt_unpack_batch: {eng._err()}")
        del keep

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

    QUAL_STREAMS = ("QUAL", "DOMQRUNS", "QUALMPLX", "DIVRQUAL")
    NAME_STREAMS = ("Q_TILE", "Q_X", "Q_Y", "Q_MISC")

    def zip_host(self):
        """host buffers in, host buffers out; the DOMQ streams stay on the device between codec_domq_compress and
        its sub-codec (GZB_OUT_DEVICE / GZB_SEC_IN_DEVICE) exactly as they stay inside one compute thread in the reference.
        The three independent pipelines of a FASTQ VBlock — QUAL (DOMQ + its four sub-streams), SEQ (ACGT + its exception
        stream) and the read-name contexts — run on one engine (host thread + stream) each, so the transfers of one
        overlap the entropy chains of another; QUAL's upload goes first because its chains are the longest.
        Returns (meta, h2d_bytes, d2h_bytes)."""
        L, V, n, H = self.L, self.V, self.n, self.h
        av, dv = self._zip_descriptors(self._rows(H["seq"]), self._rows(H["packed"]), self._rows(H["qual"]),
                                       self.line_off_h, self.line_len_h, self._rows(H["linedom"]), self._rows(H["linediv"]))
        on_dev = set(self.dq) | {"NONREF_X"}                # intermediate streams stay in HBM until their sub-codec
        in_rows = {s: self._rows(self.dq[s]) for s in self.dq}
        in_rows["NONREF_X"] = self._rows(self.x_d)
        in_rows.update({s: self._rows(H[s]) for s in self.NAME_STREAMS})
        out_rows = {s: self._rows(H["comp"][s]) for s in STREAMS}
        out_cap = {s: H["comp"][s].shape[1] for s in STREAMS}
        sflags = {s: GZB_SEC_IN_DEVICE for s in on_dev}
        nr = self.n_reads
        lens = {"Q_TILE": np.full(V, nr, np.uint32), "Q_X": np.full(V, 4 * nr, np.uint32), "Q_Y": np.full(V, 4 * nr, np.uint32), "Q_MISC": np.full(V, nr, np.uint32)}
        done = {}
        qual_up = threading.Event()

        def compress(eng, names):
            secs, mask = self._section_array(names, in_rows, lens, out_rows, out_cap, sflags)
            self._run_sections(eng.compress_raw, secs, 0)
            done[names] = (secs, mask)

        def part_qual(eng):
            try:
                if L.gzb_domq_prepare(eng.h, self.dvb, V, GZB_OUT_DEVICE):
                    raise GzbError(f"gzb_domq_prepare: {eng._err()}")
            finally:
                qual_up.set()
            if L.gzb_domq_split(eng.h, self.dvb, V, GZB_OUT_DEVICE):
                raise GzbError(f"gzb_domq_split: {eng._err()}")
            lens.update(QUAL=dv["qual_len"].copy(), DOMQRUNS=dv["runs_len"].copy(), QUALMPLX=dv["mplx_len"].copy(), DIVRQUAL=dv["divr_len"].copy())
            compress(eng, self.QUAL_STREAMS)

        def part_seq(eng):
            qual_up.wait()
            for v0 in range(0, V, 64):                          # bounded staging in the engine workspace
                v1 = min(V, v0 + 64)
                if L.gzb_acgt_pack_batch(eng.h, self._sub(self.avb, v0, v1), v1 - v0, GZB_OUT_DEVICE):
                    raise GzbError(f"gzb_acgt_pack_batch: {eng._err()}")
            lens["NONREF_X"] = np.where(av["x_all_zero"] != 0, 0, n).astype(np.uint32)
            compress(eng, ("NONREF_X",))

        def part_names(eng):
            qual_up.wait()
            compress(eng, self.NAME_STREAMS)

        self._run_parts([part_qual, part_seq, part_names])
        meta = self._make_meta(av, dv, lens)
        for names, (secs, mask) in done.items():
            self._store_comp_lens(meta, names, secs, mask)
        h2d = V * (n + n + 12 * nr) + int(sum(int(lens[s].sum()) for s in self.NAME_STREAMS))
        d2h = V * (self.packed_len + 2 * nr) + int(sum(int(secs["out_len"].sum()) for secs, _ in done.values()))
        return meta, h2d, d2h

    def piz_host(self, meta):
        L, V, n, H = self.L, self.V, self.n, self.h
        lens, clens = self._meta_arrays(meta)
        on_dev = set(self.dq) | {"NONREF_X"}
        in_rows = {s: self._rows(H["comp"][s]) for s in STREAMS}
        out_rows = {s: self._rows(self.dec_d[s] if s in on_dev else H["dec"][s]) for s in STREAMS}
        sflags = {s: GZB_SEC_OUT_DEVICE for s in on_dev}
        keep = self._piz_descriptors(meta, lens, self.line_len_h, self._rows(H["qual_out"]), self._rows(H["seq_out"]), self._rows(H["packed"]))

        def uncompress(eng, names):
            secs, _ = self._section_array(names, in_rows, clens, out_rows, lens, sflags)
            self._run_sections(eng.uncompress_raw, secs, 0)

        def part_qual(eng):
            uncompress(eng, self.QUAL_STREAMS)
            if L.gzb_domq_reconstruct(eng.h, self.pvb, V, GZB_IN_DEVICE):
                raise GzbError(f"gzb_domq_reconstruct: {eng._err()}")

        def part_seq(eng):
            uncompress(eng, ("NONREF_X",))
            for v0 in range(0, V, 64):
                v1 = min(V, v0 + 64)
                if L.gzb_acgt_unpack_batch(eng.h, self._sub(self.avb, v0, v1), v1 - v0, GZB_IN_DEVICE):
                    raise GzbError(f"gzb_acgt_unpack_batch: {eng._err()}")

        def part_names(eng):
            uncompress(eng, self.NAME_STREAMS)

        self._run_parts([part_qual, part_seq, part_names])
        del keep
        h2d = V * (4 * self.n_reads + self.packed_len) + int(sum(int(clens[s].sum()) for s in STREAMS))
        d2h = V * (n + n) + int(sum(int(lens[s].sum()) for s in self.NAME_STREAMS))
        return h2d, d2h
