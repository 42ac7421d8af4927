#!/usr/bin/env python
"""Per-stream chain latency probe: one synthetic FASTQ VBlock (bench.py's generator), each stream compressed and
decompressed ALONE through the C-ABI with the codec bench.py assigns, so that the chain kernel time is the latency
of that stream's longest leaf.  Prints ns per symbol (rANS: per step of 4 symbols).

  python tools/leaf_probe.py [--frac 1.0] [--only DIVRQUAL,NONREF_X] [--copies 1]
"""
import argparse, json, os, sys
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from genozip_b200 import Engine
from genozip_b200.fastq_path import _synth_chunk

CODECS = {"QUAL": "ARTb", "DOMQRUNS": "ARTW", "QUALMPLX": "RANB", "DIVRQUAL": "ARTb", "NONREF_X": "RANB",
          "Q_TILE": "RANB", "Q_X": "ARTW", "Q_Y": "ARTW", "Q_MISC": "ARTB"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frac", type=float, default=1.0)
    ap.add_argument("--only", default="")
    ap.add_argument("--copies", type=int, default=1)
    ap.add_argument("--codec", default="", help="override: STREAM=CODEC,...")
    a = ap.parse_args()
    codecs = dict(CODECS)
    for kv in filter(None, a.codec.split(",")):
        k, v = kv.split("="); codecs[k] = v
    eng = Engine(0)
    n_reads, read_len = 92000, 150
    d = _synth_chunk(1, n_reads, read_len, 1000 * 100003, torch.device("cuda", 0))
    seq, qual = d["seq"][0].cpu().numpy(), d["qual"][0].cpu().numpy()
    off = np.arange(n_reads, dtype=np.uint64) * np.uint64(read_len); ln = np.full(n_reads, read_len, np.uint32)
    _, x, allz = eng.acgt_pack(seq)
    e = eng.domq_encode([(qual, off, ln)])[0]
    streams = {"QUAL": e["qual"], "DOMQRUNS": e["runs"], "QUALMPLX": e["mplx"], "DIVRQUAL": e["divr"], "NONREF_X": x,
               "Q_TILE": d["Q_TILE"][0].cpu().numpy(), "Q_X": d["Q_X"][0].cpu().numpy(), "Q_Y": d["Q_Y"][0].cpu().numpy(),
               "Q_MISC": d["Q_MISC"][0].cpu().numpy()}
    only = [s for s in a.only.split(",") if s] or list(streams)
    res = {}
    for s in only:
        data = streams[s][: max(64, int(streams[s].size * a.frac))]
        c = codecs[s]
        which = 0 if c.startswith("RAN") else 1
        for rep in range(2):                                   # second pass = warm workspace
            comp = eng.compress([(c, data)] * a.copies)
            enc_ms = eng.L.gzb_last_kernel_ms(eng.h, which)
            out = eng.uncompress([(c, comp[0], data.size)] * a.copies)
            dec_ms = eng.L.gzb_last_kernel_ms(eng.h, which)
        assert np.array_equal(out[0], data)
        res[s] = dict(codec=c, n=int(data.size), nsym=int(np.unique(data).size), comp=int(comp[0].size), enc_ms=enc_ms, dec_ms=dec_ms,
                      enc_ns_per_byte=1e6 * enc_ms / data.size, dec_ns_per_byte=1e6 * dec_ms / data.size)
        print(s, json.dumps(res[s]), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
