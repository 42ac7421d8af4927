"""One BASELINE-size FASTQ (or, with `bam` as the first argument, aligned BAM) VBlock (92 000 reads x 150) through the whole codec path — host driver, C-ABI, every kernel — on the SIMT
emulator (tests/host/simt), WITHOUT a GPU: ACGT words and all sections against the reference's own compiled objects, then piz
back to the input.  Takes ~3 minutes.  Test tooling (the CPU suite runs the same path at 400 reads: tests/test_fastq_path_cpu.py)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, orc
from simt_lib import simt_engine_class
from genozip_b200.fastq_path import FastqCodecPath, synth_vblocks
if sys.argv[1:2] == ["bam"]:
    from genozip_b200.bam_path import BamCodecPath as FastqCodecPath, synth_bam_vblocks as synth_vblocks
from datagen import line_table
t0=time.time()
V, n_reads, read_len = 1, 92000, 150
data = synth_vblocks(V, n_reads, read_len, 7, torch.device("cpu"))
path = FastqCodecPath(simt_engine_class()(0), V, n_reads, read_len, n_engines=1)
codec = path.assign_codecs(data); print("codecs", codec, f"{time.time()-t0:.0f}s", flush=True)
meta = path.zip_device(data); print("zip done", f"{time.time()-t0:.0f}s", flush=True)
path.alloc_piz(meta)
seq = data["seq"][0].numpy(); qual = data["qual"][0].numpy()
off, ln = line_table(n_reads, read_len)
enc = orc.ref_domq_encode(qual, off, ln) if orc.have_gz_ref() else orc.domq_encode(qual, off, ln)
pk, x, allz = orc.ref_acgt_pack(seq)
streams = {"QUAL": enc["qual"], "DOMQRUNS": enc["runs"], "QUALMPLX": enc["mplx"], "DIVRQUAL": enc["divr"], "NONREF_X": np.zeros(0, np.uint8) if allz else x}
for k in path.NAMES: streams[k] = data[k][0].numpy()
assert np.array_equal(path.packed_d[0][:pk.size].numpy(), pk)
for s in path.STREAMS:
    assert meta[0]["len"][s] == streams[s].size, s
    if streams[s].size:
        want = orc.compress("ref", "rans" if codec[s].startswith("RAN") else "arith", streams[s], orc.ORDER[codec[s]])
        got = path.section_bytes(meta, 0, s)
        assert got.size == want.size and np.array_equal(got, want), s
        print(s, codec[s], streams[s].size, "->", want.size, "identical to the reference", flush=True)
path.scrub_intermediates()
path.piz_device(meta)
assert torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"])
for s in path.NAMES:
    assert torch.equal(path.dec_d[s][:, :data[s].shape[1]], data[s]), s
print("full-size VBlock: zip bytes identical to the reference's compiled objects, piz bit-exact", f"{time.time()-t0:.0f}s")
