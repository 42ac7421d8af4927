#!/usr/bin/env python
"""Where the pipelined host leg's time goes: raw PCIe rates of this box (pinned H2D / D2H alone and together, through gzb_stage_*),
then the pieces of PipelinedHost's zip and piz steps (wait for staging, kernels, fetch) with host timestamps.
    python tools/e2e_probe.py [--vblocks 768] [--steps 3]"""
import argparse, json, os, sys, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vblocks", type=int, default=768)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    import torch
    from genozip_b200 import Engine
    from genozip_b200.fastq_path import FastqCodecPath, PipelinedHost, synth_vblocks, ALL, UPLOADS, FETCHES
    dev = torch.device("cuda", 0)
    eng = Engine(0)
    L = eng.L
    # ---- raw link rates
    nb = 4 << 30
    h1 = torch.empty(nb, dtype=torch.uint8).pin_memory(); h2 = torch.empty(nb, dtype=torch.uint8).pin_memory()
    d1 = torch.empty(nb, dtype=torch.uint8, device=dev); d2 = torch.empty(nb, dtype=torch.uint8, device=dev)
    def timed(f):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); L.gzb_stage_wait(eng.h, ALL); return time.perf_counter() - t0
    for _ in range(2):
        tu = timed(lambda: L.gzb_stage_upload(eng.h, d1.data_ptr(), h1.data_ptr(), nb))
        td = timed(lambda: L.gzb_stage_fetch(eng.h, h2.data_ptr(), d2.data_ptr(), nb))
        tb = timed(lambda: (L.gzb_stage_upload(eng.h, d1.data_ptr(), h1.data_ptr(), nb), L.gzb_stage_fetch(eng.h, h2.data_ptr(), d2.data_ptr(), nb)))
    print(json.dumps({"pinned_h2d_GBps": nb / tu / 1e9, "pinned_d2h_GBps": nb / td / 1e9, "both_directions_each_GBps": nb / tb / 1e9}), flush=True)
    # ---- does a small batch call (descriptors up, one kernel, result back, stream synchronize) wait for a staged upload in progress?
    from genozip_b200.lib import GZB_DEVICE_PTRS
    lat = []
    eng.adler32_ptrs([(d2.data_ptr(), 1 << 20)], GZB_DEVICE_PTRS)
    L.gzb_stage_upload(eng.h, d1.data_ptr(), h1.data_ptr(), nb)
    t_start = time.perf_counter()
    while time.perf_counter() - t_start < 0.06:
        t0 = time.perf_counter(); eng.adler32_ptrs([(d2.data_ptr(), 1 << 20)], GZB_DEVICE_PTRS); lat.append(time.perf_counter() - t0)
    L.gzb_stage_wait(eng.h, ALL)
    print(json.dumps({"small_call_beside_upload_ms": {"n": len(lat), "mean": 1e3 * sum(lat) / len(lat), "max": 1e3 * max(lat), "first": 1e3 * lat[0]}}), flush=True)
    del h1, h2, d1, d2
    torch.cuda.empty_cache()
    # ---- the pipelined steps
    V = a.vblocks
    path = FastqCodecPath(eng, V, 92000, 150)
    data = synth_vblocks(V, 92000, 150, 1000, dev)
    torch.cuda.empty_cache()
    path.codec = json.load(open(os.path.join(ROOT, "bench_codecs.json")))
    meta = path.zip_device(data); path.alloc_piz(meta); path.piz_device(meta); torch.cuda.synchronize()
    host = {k: v.cpu() for k, v in data.items()}
    del data
    path.seq_out_d = path.qual_out_d = path.names_dec_d = path.dec_d = None
    torch.cuda.empty_cache()
    ph = PipelinedHost(path, host)
    log = []
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    from timeline import Tracer
    calls, origin = [], [0.0]
    for e in path.engs:                                                  # every C-ABI call of the batch, with host timestamps
        e.L = Tracer(e.L, calls, origin)
    path.L = path.engs[0].L
    orig_wait, orig_zip, orig_piz = ph._wait, path.zip_device, path.piz_device
    def tr(name, f):
        def g(*x, **k):
            t0 = time.perf_counter(); r = f(*x, **k); log.append((name, t0, time.perf_counter())); return r
        return g
    ph._wait = tr("stage_wait", orig_wait); path.zip_device = tr("zip_device", orig_zip); path.piz_device = tr("piz_device", orig_piz)
    for name, fn in (("zip_steps", ph.zip_steps), ("piz_steps", ph.piz_steps)):
        fn(1)
        del log[:]
        del calls[:]
        torch.cuda.synchronize(); t0 = time.perf_counter(); origin[0] = t0
        fn(a.steps)
        t1 = time.perf_counter()
        print(f"--- {name}({a.steps}): {1e3 * (t1 - t0) / a.steps:.1f} ms per step")
        for n_, s, e in log:
            print(f"   {n_:12s} {1e3 * (s - t0):8.1f} -> {1e3 * (e - t0):8.1f}  ({1e3 * (e - s):7.1f} ms)")
        for th, name_, c0, c1, n_ in calls[:40]:
            print(f"      {th[-3:]} {name_:30s} {1e3 * c0:8.1f} -> {1e3 * c1:8.1f}  ({1e3 * (c1 - c0):7.1f} ms, n={n_})")


if __name__ == "__main__":
    main()
