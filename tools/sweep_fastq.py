#!/usr/bin/env python
"""Sweep of the chain phase's switches on the FASTQ workload inside ONE process (the data and the buffers are made once):

    python tools/sweep_fastq.py --vblocks 512 --cfg GZB_AR_CTAS=4 --cfg GZB_AR_CTAS=16,GZB_AR0_CTAS=16 ...

Each --cfg is a comma-separated list of VAR=value (libgzb200 reads these variables at every batch call, arith_chain.cu
chain_tune).  Per configuration: one warm-up step, then `--steps` timed zip + piz passes with the inputs resident in HBM
(CUDA events on the engine's stream), the chain kernels' own durations, and a round-trip check.  One JSON line each."""
import argparse, json, os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vblocks", type=int, default=512)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--reads", type=int, default=92000)
    ap.add_argument("--read-len", type=int, default=150)
    ap.add_argument("--cfg", action="append", default=[])
    ap.add_argument("--sub-batch", type=int, default=128, help="VBlocks per codec_domq_compress call (its worst-case scratch is 4n bytes per VBlock of a sub-batch)")
    ap.add_argument("--streams", default="", help="comma-separated stream names: time only the simple-codec sections of these streams (compress + uncompress on one engine), "
                                                  "e.g. DIVRQUAL = the longest arithmetic chain of every VBlock without the other leaves around it")
    a = ap.parse_args()
    import torch
    from genozip_b200 import Engine
    from genozip_b200.fastq_path import FastqCodecPath, synth_vblocks, txt_bytes_per_vb
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    path = FastqCodecPath(eng, a.vblocks, a.reads, a.read_len, sub_batch=a.sub_batch)
    data = synth_vblocks(a.vblocks, a.reads, a.read_len, 1000, dev)
    torch.cuda.empty_cache()
    path.codec = json.load(open(os.path.join(ROOT, "bench_codecs.json")))
    meta = path.zip_device(data); path.alloc_piz(meta); path.scrub_intermediates(); path.piz_device(meta)
    torch.cuda.synchronize()
    assert torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"]), "round trip failed"
    txt = a.vblocks * txt_bytes_per_vb(a.reads, a.read_len)
    keys = set()
    for cfg in a.cfg:
        keys |= {kv.split("=")[0] for kv in cfg.split(",") if kv}
    if a.streams:
        only_streams(a, path, data, meta, keys, txt)
        return
    for cfg in a.cfg or [""]:
        for k in keys:
            os.environ.pop(k, None)
        for kv in cfg.split(","):
            if kv:
                k, v = kv.split("="); os.environ[k] = v
        tz = tp = 0.0
        kz = {}; kp = {}
        for i in range(1 + a.steps):
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            with torch.cuda.stream(path.stream):
                e0.record(path.stream)
                meta = path.zip_device(data); dz = dict(path.kernel_ms_detail)
                e1.record(path.stream)
                path.piz_device(meta); dp = dict(path.kernel_ms_detail)
                e2.record(path.stream)
            torch.cuda.synchronize()
            if i:
                tz += e0.elapsed_time(e1); tp += e1.elapsed_time(e2)
                for k, v in dz.items(): kz[k] = kz.get(k, 0.0) + v
                for k, v in dp.items(): kp[k] = kp.get(k, 0.0) + v
        ok = bool(torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"]))
        tz /= a.steps; tp /= a.steps
        print(json.dumps({"cfg": cfg, "V": a.vblocks, "zip_ms": round(tz, 1), "piz_ms": round(tp, 1), "value_GBps": round(txt / ((tz + tp) * 1e-3) / 1e9, 2),
                          "zip_kern": {k: round(v / a.steps, 1) for k, v in kz.items()}, "piz_kern": {k: round(v / a.steps, 1) for k, v in kp.items()},
                          "round_trip": ok}), flush=True)


def only_streams(a, path, data, meta, keys, txt):
    import torch
    import numpy as np
    from genozip_b200.fastq_path import NAMES, S_IDX, GZB_DEVICE_PTRS
    names = tuple(a.streams.split(","))
    eng = path.eng
    name_rows = {s: path._rows(data[s]) for s in NAMES}
    inp = path._in_ptrs(meta, name_rows, path.dq_arena.data_ptr())
    outp = path._in_ptrs(meta, {s: path._rows(path.names_dec_d[s]) for s in NAMES}, path.dq_arena.data_ptr())
    nsym = int(sum(meta.len[:, S_IDX[s]].sum() for s in names))
    for cfg in a.cfg or [""]:
        for k in keys:
            os.environ.pop(k, None)
        for kv in cfg.split(","):
            if kv:
                k, v = kv.split("="); os.environ[k] = v
        res = []
        for i in range(1 + a.steps):
            torch.cuda.synchronize()
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            with torch.cuda.stream(path.stream):
                e0.record(path.stream)
                secs, arr, vv, ss = path._section_array(meta, inp, names)
                path._compress_packed(eng, secs, arr, vv.size, GZB_DEVICE_PTRS, "comp_probe", False)
                path._kernel_ms([eng]); dz = dict(path.kernel_ms_detail)
                e1.record(path.stream)
                comp_len = arr["out_len"][:vv.size].copy(); comp_ptr = arr["out"][:vv.size].copy()
                secs2, arr2, vv2, ss2 = path._section_array(meta, inp, names)
                arr2["in_"][:vv.size] = comp_ptr; arr2["in_len"][:vv.size] = comp_len; arr2["out"][:vv.size] = outp[vv2, ss2]; arr2["out_cap"][:vv.size] = meta.len[vv2, ss2]
                eng.uncompress_raw(secs2, vv.size, GZB_DEVICE_PTRS)
                path._kernel_ms([eng]); dp = dict(path.kernel_ms_detail)
                e2.record(path.stream)
            torch.cuda.synchronize()
            if i:
                res.append((e0.elapsed_time(e1), e1.elapsed_time(e2), dz, dp))
        tz = sum(r[0] for r in res) / len(res); tp = sum(r[1] for r in res) / len(res)
        print(json.dumps({"cfg": cfg, "streams": names, "V": a.vblocks, "uncompressed_bytes": nsym, "compress_ms": round(tz, 1), "uncompress_ms": round(tp, 1),
                          "zip_kern": {k: round(v, 1) for k, v in res[-1][2].items()}, "piz_kern": {k: round(v, 1) for k, v in res[-1][3].items()}}), flush=True)


if __name__ == "__main__":
    main()
