"""CPU: the codec path over aligned BAM VBlocks (genozip_b200/bam_path.py — QUAL through DOMQ, only NONREF through codec_acgt, the
SQBITMAP / STRAND / GPOS / field / aux / QNAME streams through the simple codecs) end to end against the CPU checkers
(tests/mock_gzb.py) and with the product's CUDA sources on the SIMT emulator: every section byte-identical to the checker's,
bit-exact round trip, device-resident, host-buffer and pipelined legs.  Same assertions as tests/test_fastq_path_cpu.py."""
import numpy as np, pytest, torch
import orc
from datagen import line_table
from mock_gzb import MockEngine


def _engine(backend):
    if backend == "simt":
        from simt_lib import simt_engine_class
        return simt_engine_class()
    return MockEngine


def _oracle_sections(path, data, v, n_reads, read_len, codec):
    seq = data["seq"][v].numpy(); qual = data["qual"][v].numpy()
    off, ln = line_table(n_reads, read_len)
    pk, x, allz = orc.acgt_pack(seq)
    enc = orc.domq_encode(qual, off, ln)
    streams = {"QUAL": enc["qual"], "DOMQRUNS": enc["runs"], "QUALMPLX": enc["mplx"], "DIVRQUAL": enc["divr"], "NONREF_X": np.zeros(0, np.uint8) if allz else x}
    for k in path.NAMES:
        streams[k] = data[k][v].numpy()
    comp = {s: orc.compress("port", "rans" if codec[s].startswith("RAN") else "arith", d, orc.ORDER[codec[s]]) for s, d in streams.items() if d.size}
    return pk, streams, comp


@pytest.mark.parametrize("backend,n_engines,sub_batch", [("mock", 3, 2), ("simt", 3, 32)])
def test_bam_path_host_driver(backend, n_engines, sub_batch):
    from genozip_b200.bam_path import BamCodecPath, synth_bam_vblocks, bam_fields, nonref_len
    V, n_reads, read_len = 3, 400, 150
    data = synth_bam_vblocks(V, n_reads, read_len, 7, torch.device("cpu"))
    assert {k: t.shape[1] for k, t in data.items() if k not in ("seq", "qual")} == bam_fields(n_reads, read_len)
    assert data["seq"].shape[1] == nonref_len(n_reads, read_len) and data["seq"].shape[1] % 4 == 0 and data["qual"].shape[1] == n_reads * read_len
    data["seq"][1][data["seq"][1] == ord("N")] = ord("A")                  # VBlock 1: pure ACGT -> no NONREF_X section
    data["seq"][0][5] = ord("N")
    path = BamCodecPath(_engine(backend)(0), V, n_reads, read_len, n_engines=n_engines, sub_batch=sub_batch)
    S = path.STREAMS
    assert len(S) == 5 + 16 and S[:5] == ["QUAL", "DOMQRUNS", "QUALMPLX", "DIVRQUAL", "NONREF_X"]
    codec = path.assign_codecs(data)
    assert set(codec) == set(S)
    meta = path.zip_device(data)
    path.alloc_piz(meta)
    assert meta[1]["acgt_no_x"] and meta[1]["len"]["NONREF_X"] == 0 and not meta[0]["acgt_no_x"] and meta[0]["len"]["NONREF_X"] == path.n_seq
    for v in range(V):
        pk, streams, comp = _oracle_sections(path, data, v, n_reads, read_len, codec)
        assert np.array_equal(path.packed_d[v][:pk.size].numpy(), pk)
        for s in S:
            assert meta[v]["len"][s] == streams[s].size, (s, meta[v]["len"][s], streams[s].size)
            if streams[s].size:
                got = path.section_bytes(meta, v, s)
                assert got.size == comp[s].size and np.array_equal(got, comp[s]), f"section {s} of VB {v}"
    path.scrub_intermediates()
    path.piz_device(meta)
    assert torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"])
    for s in path.NAMES:
        assert torch.equal(path.dec_d[s][:, :data[s].shape[1]], data[s]), s
    # host-buffer mode
    path.alloc_host(data)
    meta_h, h2d, d2h = path.zip_host()
    for v in range(V):
        for s in S:
            assert meta_h[v]["len"][s] == meta[v]["len"][s] and meta_h[v]["comp_len"].get(s) == meta[v]["comp_len"].get(s)
            if meta[v]["len"][s]:
                assert np.array_equal(path.section_bytes(meta_h, v, s, host=True), path.section_bytes(meta, v, s)), f"host path: section {s}"
    path.h["seq_out"].zero_(); path.h["qual_out"].zero_(); path.scrub_intermediates()
    for t in path.h["dec"].values(): t.zero_()
    h2d_p, d2h_p = path.piz_host(meta_h)
    assert torch.equal(path.h["seq_out"], path.h["seq"]) and torch.equal(path.h["qual_out"], path.h["qual"])
    for s in path.NAMES:
        assert torch.equal(path.h["dec"][s][:, :data[s].shape[1]], data[s]), s
    n = n_reads * read_len
    fields = sum(bam_fields(n_reads, read_len).values())
    assert h2d >= V * (n + path.n_seq + fields) and d2h_p >= V * (n + path.n_seq + fields) and d2h > 0 and h2d_p > 0
    path.close()


@pytest.mark.parametrize("backend", ["mock", "simt"])
def test_bam_pipelined_host(backend):
    from genozip_b200.fastq_path import PipelinedHost
    from genozip_b200.bam_path import BamCodecPath, synth_bam_vblocks
    Eng = _engine(backend)
    V, n_reads, read_len = 3, 300, 150
    data = synth_bam_vblocks(V, n_reads, read_len, 11, torch.device("cpu"))
    ref_path = BamCodecPath(Eng(0), V, n_reads, read_len, n_engines=1)
    codec = ref_path.assign_codecs(data)
    meta = ref_path.zip_device(data)
    path = BamCodecPath(Eng(0), V, n_reads, read_len, n_engines=2)
    path.codec = dict(codec)
    path.alloc_piz(path.zip_device(data))
    ph = PipelinedHost(path, {k: v.clone() for k, v in data.items()})
    h2d, d2h = ph.zip_steps(K=3)
    for v in range(V):
        for s in path.STREAMS:
            assert ph.meta[v]["len"][s] == meta[v]["len"][s]
            if meta[v]["len"][s]:
                assert np.array_equal(path.section_bytes(ph.meta, v, s), ref_path.section_bytes(meta, v, s)), f"section {s} of VB {v}"
    n = n_reads * read_len
    assert h2d >= V * (n + path.n_seq) and 0 < d2h < V * n
    path.scrub_intermediates(); path.packed_d.zero_(); path.comp_arena.zero_(); path.comp_arena2.zero_(); ph.scrub()
    h2d_p, d2h_p = ph.piz_steps(K=3)
    assert ph.check()
    assert d2h_p >= V * (n + path.n_seq) and 0 < h2d_p < V * n
    path.close(); ref_path.close()
