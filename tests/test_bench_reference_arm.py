"""CPU: `bench.py --impl reference` (the reference's CPU implementation of the path on the host cores — the arm the driver runs
beside the GPU arm) prints the contract's JSON line for the read workloads, on the committed codec tables, and its own round trip
holds (the worker asserts it)."""
import json, os, subprocess, sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("workload,metric,table", [("fastq", "fastq_input_GBps_zip_plus_piz", "bench_codecs.json"),
                                                   ("bam", "bam_input_GBps_zip_plus_piz", "bench_codecs_bam.json")])
def test_reference_arm_line(workload, metric, table):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", workload, "--steps", "1", "--warmup", "0",
                        "--reads", "2000"], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.stdout + r.stderr)[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == metric and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["value"] > 0 and line["steps"] == 1 and line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    assert line["cpu_baseline"]["kind"] in ("reference", "port") and line["cpu_baseline"]["cores"] >= 1
    assert line["config"]["codecs"] == json.load(open(os.path.join(ROOT, table)))
    assert line["config"]["compressed_bytes_per_vblock"] > 0


def test_committed_codec_tables_cover_the_paths_streams():
    """both arms of bench.py run the committed tables: each names a simple codec for exactly the streams its path class compresses"""
    sys.path.insert(0, ROOT)
    from genozip_b200.fastq_path import STREAMS, SIMPLE
    from genozip_b200.bam_path import bam_fields
    fq = json.load(open(os.path.join(ROOT, "bench_codecs.json"))); bam = json.load(open(os.path.join(ROOT, "bench_codecs_bam.json")))
    assert list(fq) == STREAMS and set(fq.values()) <= set(SIMPLE)
    assert list(bam) == STREAMS[:5] + list(bam_fields(92000, 150)) and set(bam.values()) <= set(SIMPLE)
    assert all(bam[s] == fq[s] for s in fq)          # the streams the two workloads share got the same codecs
