"""bench_domain.py — the VCF (codec_pbwt) and long-read (codec_longr) workloads of bench.py (BASELINE.json configs[3] and [4]).

    python bench.py --workload vcf      [--vblocks V] [--steps K --warmup W]      1000-sample phased VCF: 2000 haplotypes per line
    python bench.py --workload longread [--vblocks V] ...                          Nanopore-like FASTQ, 50 kb reads

One step = one zip pass (the domain codec's transform of every VBlock of the batch) + one piz pass (its inverse) with the
inputs resident in HBM; `e2e` is the same through the C-ABI with pinned host buffers.  What is measured is the DOMAIN CODEC
alone — codec_pbwt_compress / _uncompress (src/codec_pbwt.c:244-287, :372-402) and codec_longr_compress before its sub-codec /
codec_longr_reconstruct (src/codec_longr.c:161-247, :342-373); the sub-codec sections they hand on (RUNS, FGRC, LENS, VALUES) are
ordinary simple-codec sections, measured by the FASTQ workload.  The CPU arm runs the reference's own compiled codec_pbwt.c /
codec_longr.c (oracle/_ref/libgz_ref.so) on the box's host cores, one VBlock per process at a time."""
import ctypes as C
import json, os, sys, time
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))

VCF_LINES, VCF_SAMPLES = 38000, 1000           # ~38 K variant lines x 2000 haplotypes per VBlock
LR_READ_LEN, LR_BASES = 50000, 16_000_000      # 50 kb reads, ~16 M qualities (a 32 MB FASTQ VBlock)


# ---------------------------------------------------------------------------------------------- synthetic data (on the device)
def synth_vcf_vb(n_lines, w, seed, dev, founders=48, switch=0.0015):
    """Phased haplotype matrix [n_lines, w] of '0'/'1' (+ a little '2' and '.'), Li-Stephens-like: every haplotype copies one of
    `founders` founder haplotypes and switches founder with probability `switch` per line; a line's founders carry the ALT allele
    with the line's allele frequency (most variants rare).  PBWT keeps long runs on such data, as on real phased panels."""
    import torch
    g = torch.Generator(device=dev); g.manual_seed(seed)
    sw = torch.rand((n_lines, w), generator=g, device=dev) < switch
    sw[0] = True
    idx = torch.where(sw, torch.arange(n_lines, device=dev).unsqueeze(1), torch.zeros((), dtype=torch.long, device=dev))
    last = torch.cummax(idx, dim=0).values
    del idx, sw
    pick = torch.randint(0, founders, (n_lines, w), generator=g, device=dev, dtype=torch.int64)
    f = torch.gather(pick, 0, last)
    del pick, last
    af = torch.rand((n_lines, 1), generator=g, device=dev) ** 3.3 * 0.6              # skewed to rare alleles
    fa = (torch.rand((n_lines, founders), generator=g, device=dev) < af).to(torch.uint8)
    ht = torch.gather(fa, 1, f) + ord("0")
    del f
    multi = torch.rand((n_lines, 1), generator=g, device=dev) < 0.03                 # a few multi-allelic lines and missing calls
    ht = torch.where(multi & (torch.rand((n_lines, w), generator=g, device=dev) < 0.05), torch.full_like(ht, ord("2")), ht)
    miss = torch.rand((n_lines, 1), generator=g, device=dev) < 0.02
    ht = torch.where(miss & (torch.rand((n_lines, w), generator=g, device=dev) < 0.01), torch.full_like(ht, ord(".")), ht)
    return ht.contiguous()


def synth_longread_vb(n_bases, read_len, seed, dev):
    """Nanopore-like VBlock: reads of read_len +-20 %, uniform ACGT, qualities an AR(1) process (phi 0.9) around Phred 20 clipped to
    1..50.  Returns (txt = all SEQ then all QUAL, seq_off u64, qual_off u64, lens u32) as device tensors."""
    import torch
    g = torch.Generator(device=dev); g.manual_seed(seed)
    n_reads = max(1, n_bases // read_len)
    lens = (read_len * (1 + 0.2 * torch.randn(n_reads, generator=g, device=dev))).clamp(min=4).to(torch.int64)
    n = int(lens.sum().item())
    seq = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)[torch.randint(0, 4, (n,), generator=g, device=dev)]
    e = torch.randn(n + 63, generator=g, device=dev) * 4.0
    k = (0.9 ** torch.arange(63, -1, -1, device=dev, dtype=torch.float32)).view(1, 1, 64)
    x = torch.nn.functional.conv1d(e.view(1, 1, -1), k).view(-1)
    q = (20 + x).round().clamp(1, 50).to(torch.uint8) + 33
    off = torch.cumsum(lens, 0) - lens
    txt = torch.cat([seq, q]).contiguous()
    return txt, off.to(torch.int64).contiguous(), (off + n).to(torch.int64).contiguous(), lens.to(torch.int32).contiguous()


# ---------------------------------------------------------------------------------------------- the two paths
class PbwtPath:
    metric = "vcf_gt_matrix_GBps_zip_plus_piz"
    workload = "vcf_1000_samples_phased_38Klines_x_2000ht_per_vblock (BASELINE configs[3] per-GPU share; codec_pbwt)"
    dom_kernel = ("k_pbwt_rows<0>", "k_pbwt_rows<1>")

    def __init__(self, eng, V, dev, args):
        import torch
        from genozip_b200.lib import PbwtVb
        self.eng, self.L, self.V, self.dev = eng, eng.L, V, dev
        self.n_lines, self.w = args.vcf_lines, 2 * args.vcf_samples
        self.len = self.n_lines * self.w
        self.ht = [synth_vcf_vb(self.n_lines, self.w, 7000 + args.seed_base + v, dev) for v in range(V)]
        # capacities from VBlock 0's actual streams (the reference grows its buffers as it goes, codec_pbwt.c:249-263)
        probe = (PbwtVb * 1)()
        cap_r, cap_f = self.len // 2 + 1024, self.len // 4 + 1024
        r0 = torch.empty(cap_r, dtype=torch.int32, device=dev); f0 = torch.empty(cap_f, dtype=torch.int32, device=dev)
        self._fill_enc(probe[0], self.ht[0], r0, f0)
        self._call("gzb_pbwt_encode_batch", probe, 1, 1)
        self.cap_r, self.cap_f = int(probe[0].n_runs * 1.5) + 4096, int(probe[0].n_fgrc * 1.5) + 4096
        del r0, f0
        self.runs = torch.empty((V, self.cap_r), dtype=torch.int32, device=dev)
        self.fgrc = torch.empty((V, self.cap_f), dtype=torch.int32, device=dev)
        self.out = torch.empty((V, self.len), dtype=torch.uint8, device=dev)
        self.enc = (PbwtVb * V)(); self.dec = (PbwtVb * V)()
        self.h = None

    def _fill_enc(self, a, ht, runs, fgrc):
        a.ht = ht.data_ptr(); a.n_lines = self.n_lines; a.ht_per_line = self.w
        a.runs = runs.data_ptr(); a.runs_cap = runs.numel(); a.fgrc = fgrc.data_ptr(); a.fgrc_cap = fgrc.numel()

    def _call(self, fn, arr, n, flags):
        rc = getattr(self.L, fn)(self.eng.h, arr, n, flags)
        if rc:
            from genozip_b200 import GzbError
            raise GzbError(f"{fn} failed ({rc}): {self.eng._err()}")

    def zip_device(self):
        for v in range(self.V):
            self._fill_enc(self.enc[v], self.ht[v], self.runs[v], self.fgrc[v])
        self._call("gzb_pbwt_encode_batch", self.enc, self.V, 1)
        self.kernel_ms = self.L.gzb_last_kernel_ms(self.eng.h, 2)
        return [(self.enc[v].n_runs, self.enc[v].n_fgrc) for v in range(self.V)]

    def piz_device(self, meta):
        for v in range(self.V):
            a = self.dec[v]
            a.ht = self.out[v].data_ptr(); a.ht_cap = self.len; a.n_lines = self.n_lines
            a.runs = self.runs[v].data_ptr(); a.n_runs = meta[v][0]; a.fgrc = self.fgrc[v].data_ptr(); a.n_fgrc = meta[v][1]
        self._call("gzb_pbwt_decode_batch", self.dec, self.V, 1)
        self.kernel_ms = self.L.gzb_last_kernel_ms(self.eng.h, 2)

    def check(self):
        import torch
        assert all(torch.equal(self.out[v].view(self.n_lines, self.w), self.ht[v]) for v in range(self.V)), "PBWT round trip failed"

    def input_bytes(self):
        return self.V * self.len

    def algorithmic_bytes(self, meta):                                    # rows x cols + 4 (#runs + #fgrc), per direction
        return self.V * self.len + 4 * sum(r + f for r, f in meta)

    def compressed_bytes(self, meta):
        return 4 * sum(r + f for r, f in meta)

    # ---- host buffers
    def alloc_host(self, meta):
        import torch
        pin = lambda t: t.pin_memory()
        self.h = dict(ht=[pin(t.cpu()) for t in self.ht], runs=pin(torch.empty((self.V, self.cap_r), dtype=torch.int32)),
                      fgrc=pin(torch.empty((self.V, self.cap_f), dtype=torch.int32)), out=pin(torch.empty((self.V, self.len), dtype=torch.uint8)))

    def zip_host(self):
        H = self.h
        for v in range(self.V):
            self._fill_enc(self.enc[v], H["ht"][v], H["runs"][v], H["fgrc"][v])
        self._call("gzb_pbwt_encode_batch", self.enc, self.V, 0)
        meta = [(self.enc[v].n_runs, self.enc[v].n_fgrc) for v in range(self.V)]
        return meta, self.V * self.len, 4 * sum(r + f for r, f in meta)

    def piz_host(self, meta):
        H = self.h
        for v in range(self.V):
            a = self.dec[v]
            a.ht = H["out"][v].data_ptr(); a.ht_cap = self.len; a.n_lines = self.n_lines
            a.runs = H["runs"][v].data_ptr(); a.n_runs = meta[v][0]; a.fgrc = H["fgrc"][v].data_ptr(); a.n_fgrc = meta[v][1]
        self._call("gzb_pbwt_decode_batch", self.dec, self.V, 0)
        return 4 * sum(r + f for r, f in meta), self.V * self.len

    def check_host(self):
        import torch
        assert all(torch.equal(self.h["out"][v].view(self.n_lines, self.w), self.h["ht"][v]) for v in range(self.V)), "PBWT host round trip failed"

    def release_device(self):
        self.ht = self.runs = self.fgrc = self.out = None

    def cpu_sample(self, n_vb):
        return [self.ht[v].cpu().numpy() for v in range(n_vb)]

    @staticmethod
    def cpu_one(orc, ht, direction, state):
        if direction == "zip":
            return orc.ref_pbwt_encode(ht)
        runs, fgrc = state
        return orc.ref_pbwt_decode(runs, fgrc, ht.shape[0], ht.size)

    def per_vb_hbm(self):
        return 0


class LongrPath:
    metric = "longread_fastq_GBps_zip_plus_piz"
    workload = "fastq_nanopore_50kb_reads_16M_quals_per_vblock (BASELINE configs[4] per-GPU share; codec_longr)"
    dom_kernel = ("k_longr_channels", "k_longr_decode")

    def __init__(self, eng, V, dev, args):
        import torch
        from genozip_b200.lib import LongrVb
        self.eng, self.L, self.V, self.dev = eng, eng.L, V, dev
        self.vbs = [synth_longread_vb(args.lr_bases, args.lr_read_len, 9000 + args.seed_base + v, dev) for v in range(V)]
        self.n = [int(vb[3].sum().item()) for vb in self.vbs]
        self.values = [torch.empty(n + 16, dtype=torch.uint8, device=dev) for n in self.n]
        self.lens_be = torch.empty((V, 65536), dtype=torch.int32, device=dev)
        self.out = [torch.empty(n + 16, dtype=torch.uint8, device=dev) for n in self.n]
        self.arr = (LongrVb * V)()
        # codec_longr_segconf_calculate_bins on the first VBlock (segconf runs on VB 1 only)
        self._fill(self.arr[0], 0, np.zeros(256, np.uint8))
        v2b = np.zeros(256, np.uint8)
        rc = self.L.gzb_longr_calculate_bins(eng.h, self.arr, 1, v2b.ctypes.data)
        assert rc == 0, eng._err()
        self.v2b = v2b
        self.h = None

    def _fill(self, a, v, v2b, host=None):
        txt, so, qo, ln = self.vbs[v] if host is None else host["vbs"][v]
        a.txt = txt.data_ptr(); a.txt_len = txt.numel(); a.seq_off = so.data_ptr(); a.qual_off = qo.data_ptr(); a.len = ln.data_ptr()
        a.is_rev = None; a.n_lines = ln.numel(); a.n_bases = self.n[v]
        C.memmove(a.value_to_bin, v2b.ctypes.data, 256)
        a.values = (self.values[v] if host is None else host["values"][v]).data_ptr()
        a.lens_be = (self.lens_be[v] if host is None else host["lens_be"][v]).data_ptr()
        a.qual_out = (self.out[v] if host is None else host["out"][v]).data_ptr()
        a.missing = None; a.qual_len = None

    def _call(self, fn, flags):
        rc = getattr(self.L, fn)(self.eng.h, self.arr, self.V, flags)
        if rc:
            from genozip_b200 import GzbError
            raise GzbError(f"{fn} failed ({rc}): {self.eng._err()}")
        self.kernel_ms = self.L.gzb_last_kernel_ms(self.eng.h, 2)

    def zip_device(self):
        for v in range(self.V):
            self._fill(self.arr[v], v, self.v2b)
        self._call("gzb_longr_encode", 1)
        return [(n,) for n in self.n]

    def piz_device(self, meta):
        for v in range(self.V):
            self._fill(self.arr[v], v, self.v2b)
        self._call("gzb_longr_decode", 1)

    def check(self):
        import torch
        for v in range(self.V):
            n = self.n[v]
            assert torch.equal(self.out[v][:n], self.vbs[v][0][n:2 * n]), f"LONGR round trip failed (VBlock {v})"

    def input_bytes(self):                                                 # FASTQ text the VBlocks represent: SEQ + QUAL + name line, '+', 4 newlines per read
        return sum(2 * n + 64 * vb[3].numel() for n, vb in zip(self.n, self.vbs))

    def algorithmic_bytes(self, meta):                                    # 2N read (SEQ, QUAL) + N written + 256 KB of lengths, per direction
        return sum(3 * n + 262144 for n in self.n)

    def compressed_bytes(self, meta):
        return sum(n + 262144 for n in self.n)

    def alloc_host(self, meta):
        import torch
        pin = lambda t: t.pin_memory()
        self.h = dict(vbs=[tuple(pin(t.cpu()) for t in vb) for vb in self.vbs], values=[pin(torch.empty(n + 16, dtype=torch.uint8)) for n in self.n],
                      lens_be=pin(torch.empty((self.V, 65536), dtype=torch.int32)), out=[pin(torch.empty(n + 16, dtype=torch.uint8)) for n in self.n])

    def zip_host(self):
        for v in range(self.V):
            self._fill(self.arr[v], v, self.v2b, self.h)
        self._call("gzb_longr_encode", 0)
        h2d = sum(2 * n + 20 * vb[3].numel() for n, vb in zip(self.n, self.vbs)); d2h = sum(n + 262144 for n in self.n)
        return [(n,) for n in self.n], h2d, d2h

    def piz_host(self, meta):
        for v in range(self.V):
            self._fill(self.arr[v], v, self.v2b, self.h)
        self._call("gzb_longr_decode", 0)
        h2d = sum(2 * n + 12 * vb[3].numel() + n + 262144 for n, vb in zip(self.n, self.vbs)); d2h = sum(self.n)
        return h2d, d2h

    def check_host(self):
        import torch
        for v in range(self.V):
            n = self.n[v]
            assert torch.equal(self.h["out"][v][:n], self.h["vbs"][v][0][n:2 * n]), "LONGR host round trip failed"

    def release_device(self):
        import torch
        self.vbs = [(torch.empty(0), torch.empty(0), torch.empty(0), vb[3].clone()) for vb in self.vbs]   # (the line lengths stay: the byte accounting reads them)
        self.values = self.out = None; self.lens_be = None

    def cpu_sample(self, n_vb):
        return [tuple(t.cpu().numpy() for t in self.vbs[v]) for v in range(n_vb)]

    @staticmethod
    def cpu_one(orc, vb, direction, state):
        txt, so, qo, ln = vb
        if direction == "zip":
            return orc.ref_longr_encode(txt, so.astype(np.uint64), qo.astype(np.uint64), ln.astype(np.uint32), None)
        v2b, values, lens_be = state
        return orc.ref_longr_decode(txt, so.astype(np.uint64), ln.astype(np.uint32), None, v2b, values, lens_be)


PATHS = {"vcf": PbwtPath, "longread": LongrPath}


# ---------------------------------------------------------------------------------------------- the reference's CPU path
_CPU = {}


def _cpu_worker(wid, n_workers, bar, q):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    cls, sample = _CPU["cls"], _CPU["sample"]
    mine = list(range(wid, len(sample), n_workers))
    bar.wait()
    st = [cls.cpu_one(orc, sample[v], "zip", None) for v in mine]
    bar.wait()
    for v, s in zip(mine, st):
        cls.cpu_one(orc, sample[v], "piz", s)
    bar.wait()
    q.put(True)


def cpu_time(cls, sample, workers):
    """the reference's compiled codec (oracle/_ref/libgz_ref.so) on `workers` host processes; returns (t_zip, t_piz)"""
    import multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import orc
    assert orc.have_gz_ref(), "oracle/_ref/libgz_ref.so is missing: run `make -C oracle ref` where /root/reference exists"
    orc.gz_ref()
    _CPU.update(cls=cls, sample=sample)
    ctx = mp.get_context("fork")
    workers = max(1, min(workers, len(sample)))
    bar, q = ctx.Barrier(workers + 1), ctx.Queue()
    ps = [ctx.Process(target=_cpu_worker, args=(w, workers, bar, q)) for w in range(workers)]
    [p.start() for p in ps]
    bar.wait(); t0 = time.perf_counter()
    bar.wait(); t1 = time.perf_counter()
    bar.wait(); t2 = time.perf_counter()
    oks = [q.get(timeout=1200) for _ in ps]
    [p.join() for p in ps]
    assert all(oks)
    return t1 - t0, t2 - t1


def synth_numpy(workload, n_vb, args):
    """host-only twin of the synthetic data for the --impl reference arm (no GPU needed): torch on the CPU, same generators"""
    import torch
    dev = torch.device("cpu")
    if workload == "vcf":
        return [synth_vcf_vb(args.vcf_lines, 2 * args.vcf_samples, 7000 + args.seed_base + v, dev).numpy() for v in range(n_vb)]
    return [tuple(t.numpy() for t in synth_longread_vb(args.lr_bases, args.lr_read_len, 9000 + args.seed_base + v, dev)) for v in range(n_vb)]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cls = PATHS[args.workload]
    cores = os.cpu_count() or 1
    n_vb = cores                                               # one VBlock per host process per step
    sample = synth_numpy(args.workload, n_vb, args)
    nbytes = sum(s.size for s in sample) if args.workload == "vcf" else sum(2 * int(s[3].sum()) + 64 * s[3].size for s in sample)
    ts = []
    for i in range(args.warmup + args.steps):
        tz, tp = cpu_time(cls, sample, cores)
        if i >= args.warmup:
            ts.append((tz, tp))
    tz = sum(t[0] for t in ts); tp = sum(t[1] for t in ts)
    val = nbytes * len(ts) / (tz + tp) / 1e9
    print(json.dumps({
        "impl": "reference", "metric": cls.metric, "value": val, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (tz + tp) / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "zip_GBps": nbytes * len(ts) / tz / 1e9, "piz_GBps": nbytes * len(ts) / tp / 1e9,
        "config": {"workload": cls.workload, "vblocks_per_step": n_vb},
        "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "reference", "sample": f"{n_vb} VBlocks per step, one per host process"},
        "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- GPU arm
def run_gpu(args, ClockSampler):
    import torch
    import torch.distributed as dist
    from genozip_b200 import Engine, GzbError
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    eng = Engine(local)
    cls = PATHS[args.workload]
    args.seed_base = 100000 * rank                              # VBlocks are sharded by vblock_i: every rank owns different ones
    V = args.vblocks
    if V <= 0:
        free_b, _ = torch.cuda.mem_get_info(dev)
        if args.workload == "vcf":
            per_vb = 4.6 * args.vcf_lines * 2 * args.vcf_samples    # matrix, decoded matrix, traversal-order alleles, RUNS/FGRC + records, prefixes
            V = int(max(8, min(296, (0.80 * free_b) // per_vb)))    # 296 = two CTAs per SM
        else:
            per_vb = 7.5 * args.lr_bases + (10 << 20)               # text (2N), values, decoded quals, base_chan (2N), tables (9 MB)
            V = int(max(8, min(1184, (0.80 * free_b) // per_vb)))   # one warp per VBlock: 8 per SM
    if world > 1:
        t = torch.tensor([V], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN); V = int(t.item())
    path = cls(eng, V, dev, args)
    meta = path.zip_device(); path.piz_device(meta); torch.cuda.synchronize(); path.check()       # correctness gate before timing
    stream = torch.cuda.ExternalStream(eng.L.gzb_engine_stream(eng.h), device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fz, fp, steps, warmup):
        tz = tp = kz = kp = 0.0
        m = r = None
        for i in range(warmup + steps):
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            with torch.cuda.stream(stream):
                e0.record(stream); m = fz(); a = path.kernel_ms
                e1.record(stream); r = fp(m if isinstance(m, list) else m[0]); b = path.kernel_ms
                e2.record(stream)
            barrier()
            if i >= warmup:
                tz += e0.elapsed_time(e1); tp += e1.elapsed_time(e2); kz += a; kp += b
        t = torch.tensor([tz, tp, kz, kp], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return [x / steps for x in t.tolist()], (m, r)

    clocks = ClockSampler(local); clocks.start()
    l0 = eng.launches
    (zip_ms, piz_ms, kz, kp), (meta, _) = timed(path.zip_device, path.piz_device, args.steps, args.warmup)
    launches = (eng.launches - l0) // (args.steps + args.warmup) * args.steps
    clk = clocks.stop()
    nbytes = path.input_bytes()
    value = world * nbytes / ((zip_ms + piz_ms) * 1e-3) / 1e9

    cpu_sample = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu_sample = path.cpu_sample(min(V, os.cpu_count() or 1))
    e2e = None
    if not args.no_e2e:
        path.alloc_host(meta)
        path.release_device(); eng.trim(); torch.cuda.empty_cache()        # the host-buffer leg stages everything in the engine's workspace: the resident copies go first
        (ez, ep, _, _), (zr, pr) = timed(path.zip_host, path.piz_host, max(2, args.steps // 2), 1)
        path.check_host()
        e2e = {"value": world * nbytes / ((ez + ep) * 1e-3) / 1e9, "unit": "GB/s", "h2d_bytes_per_step": int(world * (zr[1] + pr[0])),
               "d2h_bytes_per_step": int(world * (zr[2] + pr[1])), "zip_ms": ez, "piz_ms": ep}

    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    alg = path.algorithmic_bytes(meta)
    dom = 0 if kz >= kp else 1
    kms = (kz, kp)[dom]
    achieved = alg / (kms * 1e-3) / 1e9 if kms > 0 else 0.0
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tj.get(cls.dom_kernel[dom])
        if ent:                                                 # the committed ncu capture, scaled to this launch's VBlocks (PBWT: same matrix size; LONGR: per base)
            scale = V / ent["vblocks"] * (args.lr_bases / ent.get("bases_per_vblock", args.lr_bases) if args.workload == "longread" else 1.0)
            traffic = int(ent["bytes_per_launch"] * scale)
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": cls.dom_kernel[dom], "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
            "algorithmic_bytes_per_launch": alg, "launch_ms": kms, "kernel_ms_per_step": {cls.dom_kernel[0]: kz, cls.dom_kernel[1]: kp},
            "note": "rows (PBWT) / bases (LONGR) of one VBlock are a serial chain fixed by the format; one CTA / one warp walks a VBlock, throughput = VBlocks in flight / chain latency"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n_vb = min(V, cores)
        sample = cpu_sample
        tz, tp = cpu_time(cls, sample, cores)
        nb = nbytes * n_vb / V
        cpu = {"value": nb / (tz + tp) / 1e9, "unit": "GB/s", "cores": cores, "kind": "reference",
               "sample": f"{n_vb} of the step's {V} VBlocks, one per host process (zip {tz:.2f} s, piz {tp:.2f} s); the reference's compiled codec from oracle/_ref/libgz_ref.so",
               "zip_GBps": nb / tz / 1e9, "piz_GBps": nb / tp / 1e9}
    if rank == 0:
        print(json.dumps({
            "metric": cls.metric, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": zip_ms + piz_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "zip_GBps": world * nbytes / (zip_ms * 1e-3) / 1e9, "piz_GBps": world * nbytes / (piz_ms * 1e-3) / 1e9,
            "config": {"workload": cls.workload, "vblocks_per_gpu_per_step": V, "input_bytes_per_step_per_gpu": nbytes,
                       "compressed_bytes_per_step": path.compressed_bytes(meta), "l2": "inputs are larger than L2; no flush needed",
                       "scope": "the domain codec's transform only; the sub-codec sections it hands on are simple-codec sections (fastq workload)",
                       "sharding": "VBlocks round-robin by vblock_i, no data-path collective"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()
