"""GPU: the codec path over aligned BAM VBlocks (genozip_b200/bam_path.py — what `bench.py --workload bam` times) at BASELINE
VBlock size, device-pointer and host-buffer mode, section by section against the reference's compiled objects (else the
restatement) and by round trip.  (Collected last: added with the round's last GPU seconds — the 3 000-read case ran on a B200 up to the final host-buffer
comparison of the field streams, which compared a host tensor with a device tensor (gpurun_out/c47_bam_pytest.log) and is fixed; the
same assertions on CPUs: tests/test_bam_path_cpu.py, on the checkers and with the CUDA sources on the emulator.)"""
import numpy as np, pytest, torch
import orc
from datagen import line_table

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from genozip_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("n_reads", [3000, 92000])
def test_bam_path_device_and_host(eng, n_reads):
    from genozip_b200.bam_path import BamCodecPath, synth_bam_vblocks
    V, read_len = 2, 150
    dev = torch.device("cuda", 0)
    data = synth_bam_vblocks(V, n_reads, read_len, 7, dev)
    path = BamCodecPath(eng, V, n_reads, read_len)
    S = path.STREAMS
    codec = path.assign_codecs(data)
    assert set(codec) == set(S)
    meta = path.zip_device(data)
    path.alloc_piz(meta)
    off, ln = line_table(n_reads, read_len)
    impl = "ref" if orc.have_ref() else "port"
    for v in range(V):
        pk, x, allz = orc.acgt_pack(data["seq"][v].cpu().numpy())
        enc = orc.domq_encode(data["qual"][v].cpu().numpy(), off, ln)
        streams = {"QUAL": enc["qual"], "DOMQRUNS": enc["runs"], "QUALMPLX": enc["mplx"], "DIVRQUAL": enc["divr"], "NONREF_X": np.zeros(0, np.uint8) if allz else x}
        for k in path.NAMES:
            streams[k] = data[k][v].cpu().numpy()
        assert np.array_equal(path.packed_d[v][:pk.size].cpu().numpy(), pk), "ACGT words differ from the oracle"
        for s in S:
            assert meta[v]["len"][s] == streams[s].size, (s, meta[v]["len"][s], streams[s].size)
            if streams[s].size:
                want = orc.compress(impl, "rans" if codec[s].startswith("RAN") else "arith", streams[s], orc.ORDER[codec[s]])
                got = path.section_bytes(meta, v, s)
                assert got.size == want.size and np.array_equal(got, want), f"section {s} of VB {v} differs from the reference bytes"
    path.scrub_intermediates()
    path.piz_device(meta)
    torch.cuda.synchronize()
    assert torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"])
    for s in path.NAMES:
        assert torch.equal(path.dec_d[s][:, :data[s].shape[1]], data[s]), s
    path.alloc_host(data)
    meta_h, h2d, d2h = path.zip_host()
    for v in range(V):
        for s in S:
            assert meta_h[v]["len"][s] == meta[v]["len"][s] and meta_h[v]["comp_len"].get(s) == meta[v]["comp_len"].get(s)
            if meta[v]["len"][s]:
                assert np.array_equal(path.section_bytes(meta_h, v, s, host=True), path.section_bytes(meta, v, s)), f"host path: section {s}"
    path.h["seq_out"].zero_(); path.h["qual_out"].zero_(); path.scrub_intermediates()
    path.piz_host(meta_h)
    assert torch.equal(path.h["seq_out"], path.h["seq"]) and torch.equal(path.h["qual_out"], path.h["qual"])
    for s in path.NAMES:
        assert torch.equal(path.h["dec"][s][:, :data[s].shape[1]], path.h[s]), s    # (host buffers against the host copies of the inputs)
    assert h2d > 0 and d2h > 0
    path.close()
