#!/usr/bin/env python
"""bench_codecs_bam.json: codec_assign_best_codec's size criterion (src/codec.c:234-389) over the eight in-scope codecs on the first
<= 99,999 bytes of every stream of VB 1 of the BAM workload — BamCodecPath.assign_codecs itself, here answered by the CPU checkers
(tests/mock_gzb.py) so that the table can be committed without a GPU.  The GPU arm re-derives it on the device every run and
reports whether it agrees (`config.codecs_rederived_equal`); both arms run the committed table."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from mock_gzb import MockEngine
from genozip_b200.bam_path import BamCodecPath, synth_bam_vblocks

n_reads, read_len = int(sys.argv[1]) if len(sys.argv) > 1 else 92000, 150
data = synth_bam_vblocks(1, n_reads, read_len, 1000, torch.device("cpu"))
path = BamCodecPath(MockEngine(0), 1, n_reads, read_len, n_engines=1)
table = path.assign_codecs(data)
json.dump(table, open(os.path.join(ROOT, "bench_codecs_bam.json"), "w"), indent=1)
print(json.dumps(table))
