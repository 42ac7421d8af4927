#!/bin/bash
# round 2, GPU call 20: final sources — parity, link rates through the blocking-sync feeders, the default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c20_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c20_pytest.log)"
timeout 600 python tools/e2e_probe.py --vblocks 256 --steps 2 2>&1 | head -3 | cut -c1-300
timeout 1200 python bench.py > gpurun_out/c20_bench_fastq.json 2> gpurun_out/c20_bench_fastq.log; echo "fastq rc=$?"; cat gpurun_out/c20_bench_fastq.json | cut -c1-4000; tail -3 gpurun_out/c20_bench_fastq.log
python -c "
import __graft_entry__ as g
g.smoke()"
