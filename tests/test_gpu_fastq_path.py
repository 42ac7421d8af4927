"""GPU: the whole FASTQ VBlock codec path (genozip_b200/fastq_path.py — what bench.py times) at BASELINE VBlock size,
device-pointer mode and host-buffer mode, checked section by section against the oracle and by round trip."""
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


def _oracle_sections(data, v, n_reads, read_len, codec):
    seq = data["seq"][v].cpu().numpy(); qual = data["qual"][v].cpu().numpy()
    off, ln = line_table(n_reads, read_len)
    pk, x, allz = orc.acgt_pack(seq)
    enc = orc.domq_encode(qual, off, ln)
    streams = {"QUAL": enc["qual"], "DOMQRUNS": enc["runs"], "QUALMPLX": enc["mplx"], "DIVRQUAL": enc["divr"], "NONREF_X": np.zeros(0, np.uint8) if allz else x}
    for k in ("Q_TILE", "Q_X", "Q_Y", "Q_MISC"):
        streams[k] = data[k][v].cpu().numpy()
    impl = "ref" if orc.have_ref() else "port"
    comp = {s: orc.compress(impl, "rans" if codec[s].startswith("RAN") else "arith", d, orc.ORDER[codec[s]]) for s, d in streams.items() if d.size}
    return pk, streams, comp


@pytest.mark.parametrize("n_reads", [3000, 92000])
def test_fastq_path_device_and_host(eng, n_reads):
    from genozip_b200.fastq_path import FastqCodecPath, synth_vblocks, STREAMS
    V, read_len = 2, 150
    dev = torch.device("cuda", 0)
    data = synth_vblocks(V, n_reads, read_len, 7, dev)
    path = FastqCodecPath(eng, V, n_reads, read_len)
    codec = path.assign_codecs(data)
    assert set(codec) == set(STREAMS)
    meta = path.zip_device(data)
    path.alloc_piz(meta)
    for v in range(V):
        pk, streams, comp = _oracle_sections(data, v, n_reads, read_len, codec)
        assert np.array_equal(path.packed_d[v][:pk.size].cpu().numpy(), pk), "ACGT words differ from the oracle"
        for s in STREAMS:
            assert meta[v]["len"][s] == streams[s].size, (s, meta[v]["len"][s], streams[s].size)
            if streams[s].size:
                got = path.section_bytes(meta, v, s)
                assert got.size == comp[s].size and np.array_equal(got, comp[s]), f"section {s} of VB {v} differs from the reference bytes"
    path.scrub_intermediates()                                             # piz decodes into the buffers zip's intermediates occupied
    path.piz_device(meta)
    torch.cuda.synchronize()
    assert torch.equal(path.seq_out_d, data["seq"]) and torch.equal(path.qual_out_d, data["qual"])
    # host-buffer path (pinned host memory in, host memory out; DOMQ streams stay on the device between the two codecs)
    path.alloc_host(data)
    meta_h, h2d, d2h = path.zip_host()
    for v in range(V):
        for s in STREAMS:
            assert meta_h[v]["len"][s] == meta[v]["len"][s] and meta_h[v]["comp_len"].get(s) == meta[v]["comp_len"].get(s)
            if meta[v]["len"][s]:
                a = path.section_bytes(meta_h, v, s, host=True); b = path.section_bytes(meta, v, s)
                assert np.array_equal(a, b), f"host path: section {s} differs from the device path"
    path.h["seq_out"].zero_(); path.h["qual_out"].zero_(); path.scrub_intermediates()
    path.piz_host(meta_h)
    assert torch.equal(path.h["seq_out"], path.h["seq"]) and torch.equal(path.h["qual_out"], path.h["qual"])
    assert h2d > 0 and d2h > 0
