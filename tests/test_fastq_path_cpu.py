"""CPU: the host driver of the FASTQ codec path (genozip_b200/fastq_path.py — descriptor set-up, per-pipeline engines and
threads, buffer sizing, stream bookkeeping, byte accounting) run end to end against tests/mock_gzb.py, a stand-in for
libgzb200.so that computes every entry point with the CPU checkers — and against the product's own CUDA sources executed by
the SIMT emulator (tests/host/simt; "device" memory is host memory there), which takes the whole path, kernels included,
through the C-ABI on a machine without a GPU.  Same assertions as the GPU test of this path."""
import numpy as np, pytest, torch
import orc
from datagen import line_table
from mock_gzb import MockEngine


def _oracle_sections(data, v, n_reads, read_len, codec):
    seq = data["seq"][v].numpy(); qual = data["qual"][v].numpy()
    off, ln = line_table(n_reads, read_len)
    pk, x, allz = orc.acgt_pack(seq)
    enc = orc.domq_encode(qual, off, ln)
    streams = {"QUAL": enc["qual"], "DOMQRUNS": enc["runs"], "QUALMPLX": enc["mplx"], "DIVRQUAL": enc["divr"], "NONREF_X": np.zeros(0, np.uint8) if allz else x}
    for k in ("Q_TILE", "Q_X", "Q_Y", "Q_MISC"):
        streams[k] = data[k][v].numpy()
    comp = {s: orc.compress("port", "rans" if codec[s].startswith("RAN") else "arith", d, orc.ORDER[codec[s]]) for s, d in streams.items() if d.size}
    return pk, streams, comp


@pytest.mark.parametrize("backend,n_engines,sub_batch", [("mock", 1, 32), ("mock", 3, 2), ("simt", 1, 1), ("simt", 3, 32)])
def test_fastq_path_host_driver(backend, n_engines, sub_batch):
    from genozip_b200.fastq_path import FastqCodecPath, synth_vblocks, STREAMS
    if backend == "simt":
        from simt_lib import simt_engine_class
        Eng = simt_engine_class()
    else:
        Eng = MockEngine
    V, n_reads, read_len = 3, 400, 150
    data = synth_vblocks(V, n_reads, read_len, 7, torch.device("cpu"))
    data["seq"][1][data["seq"][1] == ord("N")] = ord("A")                  # VBlock 1: pure ACGT -> acgt_no_x, no NONREF_X section
    path = FastqCodecPath(Eng(0), V, n_reads, read_len, n_engines=n_engines, sub_batch=sub_batch)   # (small sub-batches: the compact buffer grows on the way)
    codec = path.assign_codecs(data)
    assert set(codec) == set(STREAMS)
    meta = path.zip_device(data)
    path.alloc_piz(meta)
    assert meta[1]["acgt_no_x"] and meta[1]["len"]["NONREF_X"] == 0 and not meta[0]["acgt_no_x"]
    for v in range(V):
        pk, streams, comp = _oracle_sections(data, v, n_reads, read_len, codec)
        assert np.array_equal(path.packed_d[v][:pk.size].numpy(), pk)
        for s in STREAMS:
            assert meta[v]["len"][s] == streams[s].size, (s, meta[v]["len"][s], streams[s].size)
            if streams[s].size:
                got = path.section_bytes(meta, v, s)
                assert got.size == comp[s].size and np.array_equal(got, comp[s]), f"section {s} of VB {v}"
    path.scrub_intermediates()                                             # piz decodes into the buffers zip's intermediates occupied
    path.piz_device(meta)
    assert torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"])
    for s in ("Q_TILE", "Q_X", "Q_Y", "Q_MISC"):
        assert torch.equal(path.dec_d[s][:, :data[s].shape[1]], data[s])
    # host-buffer mode: separate host buffers in and out, the DOMQ / exception streams stay in "device" memory
    path.alloc_host(data)
    meta_h, h2d, d2h = path.zip_host()
    for v in range(V):
        for s in STREAMS:
            assert meta_h[v]["len"][s] == meta[v]["len"][s] and meta_h[v]["comp_len"].get(s) == meta[v]["comp_len"].get(s)
            if meta[v]["len"][s]:
                a = path.section_bytes(meta_h, v, s, host=True); b = path.section_bytes(meta, v, s)
                assert np.array_equal(a, b), f"host path: section {s}"
    path.h["seq_out"].zero_(); path.h["qual_out"].zero_(); path.scrub_intermediates()
    h2d_p, d2h_p = path.piz_host(meta_h)
    assert torch.equal(path.h["seq_out"], path.h["seq"]) and torch.equal(path.h["qual_out"], path.h["qual"])
    n = n_reads * read_len
    assert h2d >= 2 * V * n and d2h_p >= 2 * V * n and d2h > 0 and h2d_p > 0
    path.close()



@pytest.mark.parametrize("backend", ["mock", "simt"])
def test_pipelined_host(backend):
    """PipelinedHost: several steps back to back, the next step's text staged while the current step's kernels run, results fetched
    behind — same sections as the synchronous path, bit-exact round trip into the host output buffers"""
    from genozip_b200.fastq_path import FastqCodecPath, PipelinedHost, synth_vblocks, STREAMS
    if backend == "simt":
        from simt_lib import simt_engine_class
        Eng = simt_engine_class()
    else:
        Eng = MockEngine
    V, n_reads, read_len = 3, 300, 150
    data = synth_vblocks(V, n_reads, read_len, 11, torch.device("cpu"))
    ref_path = FastqCodecPath(Eng(0), V, n_reads, read_len, n_engines=1)
    codec = ref_path.assign_codecs(data)
    meta = ref_path.zip_device(data)
    path = FastqCodecPath(Eng(0), V, n_reads, read_len, n_engines=2)
    path.codec = dict(codec)
    path.alloc_piz(path.zip_device(data))                                  # (sizes the working buffers, like bench.py's device-resident leg before it)
    ph = PipelinedHost(path, {k: v.clone() for k, v in data.items()})
    h2d, d2h = ph.zip_steps(K=3)
    for v in range(V):
        for s in STREAMS:
            assert ph.meta[v]["len"][s] == meta[v]["len"][s]
            if meta[v]["len"][s]:
                assert np.array_equal(path.section_bytes(ph.meta, v, s), ref_path.section_bytes(meta, v, s)), f"section {s} of VB {v}"
    n = n_reads * read_len
    assert h2d >= 2 * V * n and 0 < d2h < V * n
    path.scrub_intermediates(); path.packed_d.zero_(); path.comp_arena.zero_(); path.comp_arena2.zero_(); ph.scrub()
    h2d_p, d2h_p = ph.piz_steps(K=3)
    assert ph.check()
    assert d2h_p >= 2 * V * n and 0 < h2d_p < V * n
    path.close(); ref_path.close()
