#!/bin/bash
# round 2, GPU call 12: stage B from shared memory, pipelined host leg (staging API), Adler-32; parity, sweep, the default bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c12_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c12_pytest.log)"
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 2 --cfg "" --cfg GZB_AR_CTAS=2,GZB_AR0_CTAS=2 > gpurun_out/c12_sweep768.jsonl 2> gpurun_out/c12_sweep768.log; echo "sweep rc=$?"; cut -c1-600 gpurun_out/c12_sweep768.jsonl; tail -3 gpurun_out/c12_sweep768.log
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 1 --streams DIVRQUAL --cfg "" 2>&1 | tail -1 | cut -c1-500
timeout 1200 python bench.py > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.log; echo "bench rc=$?"; cat gpurun_out/c12_bench.json; tail -5 gpurun_out/c12_bench.log
