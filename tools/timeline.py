#!/usr/bin/env python
"""Host-side timeline of one zip + piz step of the FASTQ codec path: every C-ABI call of every engine with its start and end
(the calls are synchronous: they return when their stream is idle), printed as a table per host thread.  Answers "which
pipeline is the long pole" and "do the engines really overlap".

  python tools/timeline.py [--vblocks 64] [--mode device|host] [--steps 2]        (GPU box)
  python tools/timeline.py --mock                                                 (CPU: the mock library of the test suite)
"""
import argparse, os, sys, threading, time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

TRACED = ("gzb_acgt_pack_batch", "gzb_acgt_unpack_batch", "gzb_domq_prepare", "gzb_domq_split", "gzb_domq_reconstruct",
          "gzb_compress_sections", "gzb_compress_sections_packed", "gzb_uncompress_sections", "gzb_copy_batch")


class Tracer:
    """wraps the library object of an engine: records (thread, call, t0, t1) for the traced entry points"""

    def __init__(self, lib, log, t_origin):
        self._lib, self._log, self._t0 = lib, log, t_origin

    def __getattr__(self, name):
        f = getattr(self._lib, name)
        if name not in TRACED:
            return f

        def wrapped(*a):
            t0 = time.perf_counter()
            r = f(*a)
            self._log.append((threading.current_thread().name, name, t0 - self._t0[0], time.perf_counter() - self._t0[0], a[2] if len(a) > 2 else 0))
            return r
        return wrapped


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--vblocks", type=int, default=64)
    ap.add_argument("--reads", type=int, default=92000)
    ap.add_argument("--mode", default="device", choices=["device", "host"])
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--mock", action="store_true")
    a = ap.parse_args()
    import torch
    from genozip_b200.fastq_path import FastqCodecPath, synth_vblocks
    if a.mock:
        from mock_gzb import MockEngine as Eng
        dev, a.reads, a.vblocks = torch.device("cpu"), min(a.reads, 300), min(a.vblocks, 3)
    else:
        from genozip_b200 import Engine as Eng
        dev = torch.device("cuda", 0)
    log, origin = [], [0.0]
    eng = Eng(0)
    path = FastqCodecPath(eng, a.vblocks, a.reads, 150)
    for e in path.engs:                                     # trace every engine's library object
        e.L = Tracer(e.L, log, origin)
    path.L = path.engs[0].L
    data = synth_vblocks(a.vblocks, a.reads, 150, 7, dev)
    path.codec = dict(path.assign_codecs(data))
    meta = path.zip_device(data); path.alloc_piz(meta); path.piz_device(meta)
    if a.mode == "host":
        path.alloc_host(data)
    for step in range(a.steps):
        del log[:]
        if dev.type == "cuda":
            torch.cuda.synchronize()
        origin[0] = time.perf_counter()
        if a.mode == "device":
            meta = path.zip_device(data); t_zip = time.perf_counter() - origin[0]
            path.piz_device(meta)
        else:
            meta = path.zip_host()[0]; t_zip = time.perf_counter() - origin[0]
            path.piz_host(meta)
        t_all = time.perf_counter() - origin[0]
        print(f"--- step {step}: zip {1e3 * t_zip:.1f} ms, piz {1e3 * (t_all - t_zip):.1f} ms ({a.vblocks} VBlocks, {a.mode} mode)")
        for th, name, t0, t1, n in sorted(log, key=lambda r: r[2]):
            print(f"  {th:28s} {name:26s} {1e3 * t0:9.1f} -> {1e3 * t1:9.1f} ms  ({1e3 * (t1 - t0):8.1f} ms, n={n})")
    path.close()


if __name__ == "__main__":
    main()
