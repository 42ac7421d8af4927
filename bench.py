#!/usr/bin/env python
"""bench.py — FASTQ input GB/s (zip + piz) of the per-VBlock codec path on B200, next to the reference's CPU path.

One "step" = one zip pass + one piz pass of the hot path over one batch of V synthetic Illumina-like FASTQ VBlocks
(BASELINE.json configs[1] scaled to one GPU: 150 bp reads, ~32 MB of FASTQ text per VBlock; codec_domq + codec_acgt
hot, read-name contexts through the simple codecs).  Out of scope and therefore NOT in the timed region: the
segmenter that produces these streams, and LZMA of the 2-bit sequence words (host).

  python bench.py --gpus N --steps K --warmup W            (N>1 via torchrun, one rank per GPU; weak scaling)
  python bench.py --impl reference ...                      the reference's CPU implementation of the same path
"""
import argparse, json, os, subprocess, sys, threading, time
import ctypes as C
import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ALL_CPUS = os.sched_getaffinity(0) if hasattr(os, "sched_getaffinity") else set()
METRIC = "fastq_input_GBps_zip_plus_piz"
UNIT = "GB/s"
WORKLOAD = "fastq_illumina_150bp_paired_vb32MB (BASELINE configs[1] per-GPU share)"
EXCLUDED = "segmenter; LZMA of the 2-bit sequence words (host, out of scope) — in both arms"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="gzb200", choices=["gzb200", "reference"])
    ap.add_argument("--vblocks", type=int, default=int(os.environ.get("GZB_BENCH_VBLOCKS", "0")),
                    help="VBlocks per GPU per step (0 = as many as fit, at most 256: the chain kernels are latency-bound, so throughput grows with the batch)")
    ap.add_argument("--reads", type=int, default=92000, help="reads per VBlock (92,000 x 150 bp ~ 32 MB of FASTQ text)")
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--workload", default="fastq", choices=["fastq", "bam", "vcf", "longread"],
                    help="fastq = BASELINE configs[1] (the headline metric); bam = configs[2] (aligned BAM VBlocks: DOMQ + the field streams, codec_acgt on NONREF only — genozip_b200/bam_path.py); vcf = configs[3] (codec_pbwt); longread = configs[4] (codec_longr) — bench_domain.py")
    ap.add_argument("--vcf-lines", type=int, default=38000); ap.add_argument("--vcf-samples", type=int, default=1000)
    ap.add_argument("--lr-bases", type=int, default=16_000_000); ap.add_argument("--lr-read-len", type=int, default=50000)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu):
        self.gpu, self.p, self.lines = gpu, None, []

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True); self.t.start()
        except Exception:
            self.p = None

    def _read(self):
        for ln in self.p.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.p:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=2)
        except Exception:
            self.p.kill()
        sm, mx, reasons = [], 0, set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx = max(mx, float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- reference / CPU arm
_CPU = {}


class Spec:
    """what differs between the two read workloads that run the same path class (fastq_path.FastqCodecPath / bam_path.BamCodecPath)"""

    def __init__(self, workload):
        self.bam = workload == "bam"
        if self.bam:
            from genozip_b200 import bam_path as B
            self.cls, self.synth, self.bytes_per_vb = B.BamCodecPath, B.synth_bam_vblocks, B.bam_bytes_per_vb
            self.metric = "bam_input_GBps_zip_plus_piz"
            self.workload = "bam_aligned_sorted_150bp_vb92Kreads (BASELINE configs[2] per-GPU share; post-seg streams, SURVEY 8d C3)"
            self.excluded = "segmenter, reference / aligner, BGZF; LZMA of NONREF's 2-bit words (host, out of scope) — in both arms"
            self.table = os.path.join(ROOT, "bench_codecs_bam.json")
            self.accounting = ("input bytes = the uncompressed BAM records the VBlocks represent (block_size + 32 fixed bytes, 40-byte read name, one CIGAR op, "
                               "4-bit SEQ, QUAL, ~28 B of aux fields per read); the segmenter is out of scope: what flows through the path is QUAL (1 B / base), "
                               "NONREF (1.5 % of the bases), SQBITMAP (1 bit / base) and ~24.1 B / read of field, aux and QNAME contexts; the same count in both arms")
        else:
            from genozip_b200 import fastq_path as F
            self.cls, self.synth, self.bytes_per_vb = F.FastqCodecPath, F.synth_vblocks, F.txt_bytes_per_vb
            self.metric, self.workload, self.excluded, self.table = METRIC, WORKLOAD, EXCLUDED, CODEC_TABLE
            self.accounting = ("input bytes = the FASTQ text the VBlocks represent (45-byte name line, SEQ, '+', QUAL, 4 newlines per read); of the name line, "
                               "10 B/read of segmented read-name contexts flow through the path (the segmenter is out of scope); the same count in both arms")

    def codecs(self):
        return json.load(open(self.table))


def _cpu_worker(wid, n_workers, n_vb, bar, q):
    """one host process: zip then piz of its share of the VBlocks, phases separated by barriers so that the parent
    times the whole pool (processes, not threads: the Python glue around the C calls must not serialise on the GIL)"""
    import orc
    data_np, n_reads, read_len, codec, impl = _CPU["data"], _CPU["n_reads"], _CPU["read_len"], _CPU["codec"], _CPU["impl"]
    off = (np.arange(n_reads, dtype=np.uint64) * np.uint64(read_len)); ln = np.full(n_reads, read_len, np.uint32)
    names = tuple(k for k in data_np if k not in ("seq", "qual"))      # the simple-codec contexts of the workload
    mine = list(range(wid, n_vb, n_workers))
    gz = orc.have_gz_ref()                                   # the reference's own compiled codec_domq.c / codec_acgt.c (-O3), else the restatement
    acgt_pack, domq_encode = (orc.ref_acgt_pack, orc.ref_domq_encode) if gz else (orc.acgt_pack, orc.domq_encode)
    acgt_unpack, domq_decode = (orc.ref_acgt_unpack, orc.ref_domq_decode) if gz else (orc.acgt_unpack, orc.domq_decode)
    bar.wait()
    zs = []
    for v in mine:
        packed, x, allz = acgt_pack(data_np["seq"][v])
        enc = domq_encode(data_np["qual"][v], off, ln)
        streams = {"QUAL": enc["qual"], "DOMQRUNS": enc["runs"], "QUALMPLX": enc["mplx"], "DIVRQUAL": enc["divr"]}
        if not allz:
            streams["NONREF_X"] = x
        for k in names:
            streams[k] = data_np[k][v]
        comp = {}
        for s_, d in streams.items():
            if d.size:
                c = codec[s_]
                comp[s_] = (orc.compress(impl, "rans" if c.startswith("RAN") else "arith", d, orc.ORDER[c]), d.size)
        zs.append(dict(packed=packed, allz=allz, enc=enc, comp=comp))
    bar.wait()
    ok = True
    for v, z in zip(mine, zs):
        dec = {}
        for s_, (c, n) in z["comp"].items():
            cc = codec[s_]
            dec[s_] = orc.uncompress(impl, "rans" if cc.startswith("RAN") else "arith", c, n)
        e = dict(z["enc"]); e.update(qual=dec["QUAL"], runs=dec.get("DOMQRUNS", np.zeros(0, np.uint8)), mplx=dec["QUALMPLX"],
                                     divr=dec.get("DIVRQUAL", np.zeros(0, np.uint8)))
        q_ = domq_decode(e, ln)
        s2 = acgt_unpack(z["packed"], None if z["allz"] else dec["NONREF_X"], data_np["seq"][v].size)
        ok = ok and np.array_equal(q_, data_np["qual"][v]) and np.array_equal(s2, data_np["seq"][v])
        ok = ok and all(np.array_equal(dec[k], data_np[k][v]) for k in names)
    bar.wait()
    q.put((ok, sum(len(c) for z in zs for c, _ in z["comp"].values())))


def cpu_path_time(data_np, n_reads, read_len, codec, workers, n_vb):
    """The reference's CPU implementation of the same path on host cores: htscodecs entry points from oracle/_ref
    (the reference's own objects) when present, else the CPU restatement; DOMQ/ACGT through the restatement.
    `workers` host processes (fork), VBlocks dealt round-robin.  Returns (t_zip, t_piz, kind)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import multiprocessing as mp
    import orc                                              # test infrastructure: used here only as the CPU baseline
    kind = "reference" if os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhts_ref.so")) else "port"
    impl = "ref" if kind == "reference" else "port"
    orc.port(); (orc.ref() if impl == "ref" else None)
    _CPU.update(data=data_np, n_reads=n_reads, read_len=read_len, codec=codec, impl=impl)
    ctx = mp.get_context("fork")
    workers = max(1, min(workers, n_vb))
    bar, q = ctx.Barrier(workers + 1), ctx.Queue()
    ps = [ctx.Process(target=_cpu_worker, args=(w, workers, n_vb, bar, q)) for w in range(workers)]
    [p.start() for p in ps]
    bar.wait(); t0 = time.perf_counter()
    bar.wait(); t1 = time.perf_counter()
    bar.wait(); t2 = time.perf_counter()
    res = [q.get(timeout=600) for _ in ps]
    [p.join() for p in ps]
    assert all(r[0] for r in res), "CPU arm: round trip failed"
    _CPU["compressed_bytes"] = int(sum(r[1] for r in res))
    return t1 - t0, t2 - t1, kind


def synth_numpy(V, n_reads, read_len, seed, synth=None):
    """the bytes of the GPU arm's first V VBlocks for the CPU-only arm: the same torch generator, on the GPU when the box has one
    (the GPU arm generates there, and CUDA's random streams differ from the CPU's), else on the CPU (same distribution, other bytes)"""
    import torch
    from genozip_b200.fastq_path import synth_vblocks
    dev = torch.device("cuda", 0) if torch.cuda.is_available() else torch.device("cpu")
    d = (synth or synth_vblocks)(V, n_reads, read_len, seed, dev)
    out = {k: [t[v].cpu().numpy() for v in range(V)] for k, t in d.items()}
    del d
    if dev.type == "cuda":
        torch.cuda.empty_cache()
    return out, dev.type


# Codec of each stream: codec_assign_best_codec's size criterion over the eight in-scope codecs on VB 1 (fastq_path.assign_codecs).
# The table is COMMITTED so that both arms run the same codecs whatever runs first; the GPU arm re-derives it every run and reports
# whether it still agrees (`codecs_rederived_equal`).
CODEC_TABLE = os.path.join(ROOT, "bench_codecs.json")


def load_codecs():
    return json.load(open(CODEC_TABLE))


def run_reference(args):
    """--impl reference: time the reference's CPU implementation on this box's host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_vb = max(2 * cores, 8)                                 # two VBlocks per host process per step
    spec = Spec(args.workload)
    txt_bytes_per_vb = spec.bytes_per_vb
    data, gen = synth_numpy(n_vb, args.reads, args.read_len, 1000, spec.synth)          # = the GPU arm's rank-0 VBlocks 0 .. n_vb-1
    codec = spec.codecs()
    ts = []
    for i in range(args.warmup + args.steps):
        tz, tp, kind = cpu_path_time(data, args.reads, args.read_len, codec, cores, n_vb)
        if i >= args.warmup:
            ts.append((tz, tp))
    tz = sum(t[0] for t in ts); tp = sum(t[1] for t in ts)
    nbytes = n_vb * txt_bytes_per_vb(args.reads, args.read_len) * len(ts)
    val = nbytes / (tz + tp) / 1e9
    sample = f"{n_vb} VBlocks x {args.reads} reads x {args.read_len} bp per step dealt to {cores} host processes"
    print(json.dumps({
        "impl": "reference", "metric": spec.metric, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * (tz + tp) / len(ts), "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "zip_GBps": nbytes / tz / 1e9, "piz_GBps": nbytes / tp / 1e9,
        "config": {"workload": spec.workload, "vblocks_per_step": n_vb,
                   "reads_per_vblock": args.reads, "read_len": args.read_len, "codecs": codec,
                   "compressed_bytes_per_vblock": _CPU.get("compressed_bytes", 0) / n_vb,
                   "data_generator": f"torch-{gen} (the GPU arm's rank-0 VBlocks 0..{n_vb - 1}" + (")" if gen == "cuda" else "; no GPU here: same distribution, other bytes)"),
                   "excluded": spec.excluded},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------- GPU arm
def bind_to_gpu_cpus(gpu):
    """run this rank on the CPUs NVML lists as local to its GPU, so that the pinned host buffers it allocates (first touch) and the
    threads that feed the copy engines sit on the GPU's NUMA node; returns what was done, for the bench line"""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1} & set(os.sched_getaffinity(0))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if cpus and len(cpus) < 6 * world and len(cpus) < len(os.sched_getaffinity(0)):
            # every rank runs ~6 busy host threads (two pipelines, two staging feeders, the interpreter); if the ranks of this box would
            # all crowd onto one NUMA node's cores, spreading out is worth more than local page-locked memory
            return f"not bound: {len(cpus)} cpus local to gpu {gpu} for up to {world} ranks"
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cpus local to gpu {gpu}"
    except Exception as ex:
        return f"not bound ({type(ex).__name__})"
    return "not bound"


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from genozip_b200 import Engine
    spec = Spec(args.workload)
    FastqCodecPath, synth_vblocks, txt_bytes_per_vb = spec.cls, spec.synth, spec.bytes_per_vb

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = bind_to_gpu_cpus(local)                             # page-locked buffers are then first-touched on the GPU's own NUMA node
    eng = Engine(local)
    from genozip_b200 import GzbError
    V = args.vblocks
    if V <= 0:                                                  # the entropy chains are latency-bound: throughput grows with the batch, memory bounds it
        free_b, _ = torch.cuda.mem_get_info(dev)
        n = args.reads * args.read_len
        per_vb = 9.3 * n + 14 * args.reads + (6 << 20)          # inputs 2n, 2-bit words n/4, exception stream n, DOMQ streams ~0.4n, outputs 2n, engine workspace ~3.5n
        V = int(max(8, min(768, (0.86 * free_b) // per_vb)))     # (measured on B200: 512 -> 17.7, 768 -> 22.2, 819 -> 21.7 GB/s: beyond ~768 the chain kernels are issue-bound)
    while True:                                                  # a batch that does not fit is halved (all ranks agree on the size)
        if world > 1:
            t = torch.tensor([V], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN); V = int(t.item())
        path = data = None
        ok = 1
        try:
            path = FastqCodecPath(eng, V, args.reads, args.read_len)
            # VBlocks are sharded round-robin by vblock_i: rank r owns vblock_i = r+1, r+1+world, ...  Seeds follow vblock_i.
            data = synth_vblocks(V, args.reads, args.read_len, 1000 + rank, dev)
            torch.cuda.empty_cache()                           # (the generator's temporaries: the engines allocate with cudaMalloc, outside torch's cache)
            codecs = path.assign_codecs(data) if rank == 0 else None
            if world > 1:
                obj = [codecs]; dist.broadcast_object_list(obj, src=0); codecs = obj[0]
            path.codec = dict(codecs)
            # correctness gate before timing: piz(zip(x)) == x on the device
            meta = path.zip_device(data)
            path.alloc_piz(meta)
            path.scrub_intermediates()                         # piz decodes into the buffers zip's intermediates occupied: empty them for the gate
            path.piz_device(meta)
            torch.cuda.synchronize()
        except (torch.OutOfMemoryError, GzbError) as ex:
            if "memory" not in str(ex).lower():
                raise
            ok = 0
        if world > 1:
            t = torch.tensor([ok], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN); ok = int(t.item())
        if ok:
            break
        if path is not None:
            path.close()
        del path, data
        import gc; gc.collect()
        eng.close(); torch.cuda.empty_cache()
        eng = Engine(local)
        V = max(4, int(V * 0.8))
    committed = spec.codecs()
    rederived_equal = committed == dict(codecs)
    codecs = committed                                         # both arms run the committed table (see CODEC_TABLE)
    path.codec = dict(codecs)
    if not rederived_equal:                                    # (sizes were planned for the re-derived table)
        meta = path.zip_device(data); path.scrub_intermediates(); path.piz_device(meta); torch.cuda.synchronize()
    assert torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"]), "round trip failed"
    STREAMS = list(path.STREAMS)
    for s in path.NAMES:
        assert torch.equal(path.dec_d[s][:, :data[s].shape[1]], data[s]), f"round trip failed: {s}"

    txt_bytes = V * txt_bytes_per_vb(args.reads, args.read_len)
    stream = path.stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def section_list_gather(meta):
        """final section-list gather: every rank's per-section compressed lengths to all ranks over NCCL"""
        if world == 1:
            return
        t = torch.tensor([[m["comp_len"].get(s, 0) for s in STREAMS] for m in meta], dtype=torch.int32, device=dev)
        out = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(out, t)

    detail = {}

    def timed(fn_zip, fn_piz, steps, warmup):
        tz = tp = 0.0
        kern = {"rans_enc": 0.0, "arith_enc": 0.0, "rans_dec": 0.0, "arith_dec": 0.0}
        detail.clear()
        for i in range(warmup + steps):
            barrier()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            with torch.cuda.stream(stream):
                e0.record(stream)
                m = fn_zip()
                section_list_gather(m[0] if isinstance(m, tuple) else m)
                kz = path.kernel_ms; dz = dict(path.kernel_ms_detail)
                e1.record(stream)
                r = fn_piz(m[0] if isinstance(m, tuple) else m)
                kp = path.kernel_ms; dp = dict(path.kernel_ms_detail)
                e2.record(stream)
            barrier()
            if i >= warmup:
                tz += e0.elapsed_time(e1); tp += e1.elapsed_time(e2)
                kern["rans_enc"] += kz[0]; kern["arith_enc"] += kz[1]; kern["rans_dec"] += kp[0]; kern["arith_dec"] += kp[1]
                for k_, v_ in dz.items(): detail["enc_" + k_] = detail.get("enc_" + k_, 0.0) + v_ / steps
                for k_, v_ in dp.items(): detail["dec_" + k_] = detail.get("dec_" + k_, 0.0) + v_ / steps
        t = torch.tensor([tz, tp], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                     # max over ranks
        return t[0].item() / steps, t[1].item() / steps, {k: v / steps for k, v in kern.items()}, (m, r)

    clocks = ClockSampler(local); clocks.start()
    l0 = path.launches
    zip_ms, piz_ms, kern, (meta, _) = timed(lambda: path.zip_device(data), lambda m: path.piz_device(m), args.steps, args.warmup)
    launches = (path.launches - l0) // (args.steps + args.warmup) * args.steps
    clk = clocks.stop()
    value = world * txt_bytes / ((zip_ms + piz_ms) * 1e-3) / 1e9

    # e2e: the same steps with HOST (page-locked) buffers in and out, every byte crossing PCIe inside the timed region, the way a host
    # that keeps handing over VBlock batches drives the C-ABI (fastq_path.PipelinedHost): Ke zip steps back to back, then Ke piz steps;
    # the next step's text is staged (gzb_stage_upload) and the previous step's results are fetched (gzb_stage_fetch) while the
    # current step's kernels run on device buffers.
    e2e = None
    ph, e2e_err = None, ""
    cpu_data = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        n_cpu = min(V, 4 * (os.cpu_count() or 1))
        cpu_data = {k: [data[k][v].cpu().numpy() for v in range(n_cpu)] for k in data}
    Ve = V
    if not args.no_e2e:
        from genozip_b200.fastq_path import PipelinedHost
        # the host leg keeps page-locked inputs and outputs (4.3n bytes per VBlock), and every rank of the box does: its batch is what the
        # box's memory allows (the device-resident leg above is bounded by the GPU's memory alone)
        try:
            import psutil
            Ve = int(max(8, min(V, (0.6 * psutil.virtual_memory().available / world) // (4.4 * args.reads * args.read_len))))
        except Exception:
            pass
        Ve = max(8, min(Ve, int(os.environ.get("GZB_E2E_VBLOCKS", Ve))))   # (to exercise the smaller-batch branch on a box with plenty of memory)
        if world > 1:
            t = torch.tensor([Ve], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN); Ve = int(t.item())
        path.seq_out_d = path.qual_out_d = path.names_dec_d = path.dec_d = None      # (the pipelined leg brings its own double-buffered outputs)
        if Ve < V:                                                        # a smaller batch for this leg: a path of its size
            data = {k: v[:Ve].contiguous() for k, v in data.items()}
            path.close(); path.release_device(); del path
            import gc; gc.collect(); torch.cuda.empty_cache()
            path = FastqCodecPath(eng, Ve, args.reads, args.read_len)
            path.codec = dict(codecs)
            path.alloc_piz(path.zip_device(data)); path.piz_device(path.meta)
            path.seq_out_d = path.qual_out_d = path.names_dec_d = path.dec_d = None
        torch.cuda.empty_cache()
        ph, e2e_err = None, ""
        try:                                                              # (page-locked memory is the box's, not this rank's: if it does not fit, every rank skips the leg)
            ph = PipelinedHost(path, data)
            del data
            torch.cuda.empty_cache()
            ph.zip_steps(1); ph.scrub(); ph.piz_steps(1)                  # warm-up step (buffers grow to their sizes) + correctness gate
            assert ph.check(), "host round trip failed"
            ph.scrub()
        except (RuntimeError, MemoryError, GzbError) as ex:
            if "memory" not in str(ex).lower():
                raise
            ph, e2e_err = None, f"{type(ex).__name__}: {str(ex)[:200]}"
        if world > 1:
            t = torch.tensor([1 if ph is not None else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
            if not int(t.item()):
                ph, e2e_err = None, e2e_err or "another rank ran out of page-locked memory"
        Ke = max(2, args.steps)
    if ph is not None:
        barrier()
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        with torch.cuda.stream(stream):
            e0.record(stream)
            h2d_z, d2h_z = ph.zip_steps(Ke)
            for _ in range(Ke):
                section_list_gather(ph.meta)
            e1.record(stream)
            h2d_p, d2h_p = ph.piz_steps(Ke)
            e2.record(stream)
        barrier()
        assert ph.check(), "host round trip failed"
        t = torch.tensor([e0.elapsed_time(e1) / Ke, e1.elapsed_time(e2) / Ke], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ez, ep = t[0].item(), t[1].item()
        e2e_bytes = Ve * txt_bytes_per_vb(args.reads, args.read_len)
        e2e = {"value": world * e2e_bytes / ((ez + ep) * 1e-3) / 1e9, "unit": UNIT, "vblocks_per_gpu_per_step": Ve, "h2d_bytes_per_step": int(world * (h2d_z + h2d_p)), "d2h_bytes_per_step": int(world * (d2h_z + d2h_p)),
               "zip_ms": ez, "piz_ms": ep, "steps": Ke,
               "how": "Ke zip steps back to back, then Ke piz steps, host buffers in and out; the next step's text is staged and the previous step's results are fetched "
                      "(gzb_stage_upload / gzb_stage_fetch) while the current step's kernels run; the first upload and the last fetch of each run are not hidden and are in the time"}

    # roofline of the dominant kernel: algorithmic bytes (N uncompressed + C compressed) / its launch time
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    which_peak = "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg = {"rans": 0, "arith": 0}
    for m in meta:
        for s, n in m["len"].items():
            if n:
                alg["rans" if codecs[s].startswith("RAN") else "arith"] += n + m["comp_len"][s]
    dom = max(kern, key=lambda k: kern[k])
    dom_bytes = alg["rans" if dom.startswith("rans") else "arith"] // len(path.groups)    # each group launches the chain phase once
    achieved = dom_bytes / (kern[dom] * 1e-3) / 1e9 if kern[dom] > 0 else 0.0
    # the arithmetic chain phase is up to three kernels side by side (general / order-0 / split encoder): name the one that lasts longest
    if dom.startswith("arith"):
        side = "enc_" if dom.endswith("enc") else "dec_"
        sub = max(("arith_general", "arith_o0", "arith_split"), key=lambda k: detail.get(side + k, 0.0))
        kname = {"enc_arith_general": "k_arith_encode_t<0>", "enc_arith_o0": "k_arith_encode_t<1>", "enc_arith_split": "k_ar_split_code",
                 "dec_arith_general": "k_arith_decode_t<0>", "dec_arith_o0": "k_arith_decode_t<1>", "dec_arith_split": "k_arith_decode_t<0>"}[side + sub]
    else:
        kname = {"rans_enc": "k_rans_encode", "rans_dec": "k_rans_decode"}[dom]
    traffic = traffic_note = None
    try:                                                        # dram__bytes of this kernel from the committed ncu --set full capture, scaled to this launch's VBlocks
        tj = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
        ent = tj.get(kname) if not spec.bam else None            # (the capture is of the FASTQ workload's leaves)
        if ent:
            traffic = int(ent["bytes_per_launch"] * V / ent["vblocks"])
            traffic_note = f"dram__bytes_read+write of {kname} captured at {ent['vblocks']} VBlocks ({ent.get('source', 'profiles/')}), scaled by {V}/{ent['vblocks']}"
    except Exception:
        pass
    # SURVEY §8d secondary bound: chains in flight x f_clk / cycles per symbol, from the longest arithmetic leaf (DIVRQUAL: 4 ranks per coder symbol)
    f_clk = (clk.get("sm_mhz") or 1965.0) * 1e6
    longest = max((m["len"].get("DIVRQUAL", 0) + 3) // 4 for m in meta) if dom.startswith("arith") else max(m["len"].get("NONREF_X", 0) // 4 for m in meta)
    chain_bound = None
    if longest and kern[dom] > 0:
        cyc = kern[dom] * 1e-3 * f_clk / longest
        chain_bound = {"longest_chain_symbols": int(longest), "cycles_per_symbol_in_batch": cyc, "chains_in_flight": V,
                       "symbols_per_s": V * f_clk / cyc, "note": "the launch lasts as long as its longest dependency chain: one arithmetic chain (4 rANS chains) per leaf, fixed by the bitstream"}
    roof = {"bound": "hbm", "kernel": kname, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_note": traffic_note,
            "peak_source": which_peak, "algorithmic_bytes_per_launch": dom_bytes, "launch_ms": kern[dom], "chain_bound": chain_bound,
            "note": "entropy chains are dependency-bound (4 chains per rANS leaf, 1 per arithmetic leaf, fixed by the bitstream): the chain phase lasts as long as "
                    "its longest leaf; algorithmic bytes = N + C of the sections of the dominant coder's chain phase (its kernels run side by side)",
            "kernel_ms_per_step": kern, "kernel_ms_detail": dict(detail)}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        try:
            os.sched_setaffinity(0, ALL_CPUS)                   # the CPU leg gets every host core again
        except Exception:
            pass
        cores = os.cpu_count() or 1
        n_vb = min(V, 4 * cores)                                # ~10-30 s of CPU work
        dnp = cpu_data
        tz, tp, kind = cpu_path_time(dnp, args.reads, args.read_len, codecs, cores, n_vb)
        nb = n_vb * txt_bytes_per_vb(args.reads, args.read_len)
        cpu = {"value": nb / (tz + tp) / 1e9, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": f"{n_vb} of the step's {V} VBlocks dealt to {min(cores, n_vb)} host processes (zip {tz:.2f} s, piz {tp:.2f} s)",
               "zip_GBps": nb / tz / 1e9, "piz_GBps": nb / tp / 1e9}

    if rank == 0:
        comp_total = sum(sum(m["comp_len"].values()) for m in meta)
        print(json.dumps({
            "metric": spec.metric, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": zip_ms + piz_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "zip_GBps": world * txt_bytes / (zip_ms * 1e-3) / 1e9, "piz_GBps": world * txt_bytes / (piz_ms * 1e-3) / 1e9,
            "config": {"workload": spec.workload, "vblocks_per_gpu_per_step": V,
                       "reads_per_vblock": args.reads, "read_len": args.read_len, "txt_bytes_per_step_per_gpu": txt_bytes,
                       "codecs": codecs, "codecs_rederived_equal": rederived_equal, "compressed_bytes_per_vblock": comp_total / V, "l2": "inputs (>= 0.9 GB per step) are larger than L2; no flush needed",
                       "sections_per_step": sum(1 for m in meta for n in m["len"].values() if n), "compressed_bytes_per_step": comp_total,
                       "txt_accounting": spec.accounting,
                       "excluded": spec.excluded, "sharding": "VBlocks round-robin by vblock_i, no data-path collective; NCCL all_gather of the section list only",
                       "engines_per_gpu": len(path.engs), "device_groups": len(path.groups), "cpu_binding": numa,
                       **({"e2e_skipped": e2e_err} if (not args.no_e2e and e2e is None) else {})},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clk, "roofline": roof, "cpu_baseline": cpu,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    a.seed_base = 0
    if a.workload not in ("fastq", "bam"):
        import bench_domain
        bench_domain.run_reference(a) if a.impl == "reference" else bench_domain.run_gpu(a, ClockSampler)
    elif a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
