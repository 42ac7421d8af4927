#!/bin/bash
# round 2, GPU call 7: ranked split stage A, long-leaf decoder (shared-memory mirror), two device pipelines: parity, sweeps, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c7_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c7_pytest.log)"
GZB_AR_LONG_MIN=64 GZB_AR_SPLIT_MIN=64 timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "hts or golden or fastq" > gpurun_out/c7_pytest_low.log 2>&1; echo "pytest (thresholds lowered) rc=$? $(tail -1 gpurun_out/c7_pytest_low.log)"
for V in 512 768; do
timeout 900 python tools/sweep_fastq.py --vblocks $V --steps 2 \
  --cfg "" --cfg GZB_AR_LONG_MIN=off --cfg GZB_AR_LONG_ENT=32 --cfg GZB_AR_LONG_MIN=16384 --cfg GZB_AR_RUN4=0 --cfg GZB_AR_CTAS=16,GZB_AR0_CTAS=16 --cfg GZB_AR_CTAS=2,GZB_AR0_CTAS=2 \
  > gpurun_out/c7_sweep$V.jsonl 2> gpurun_out/c7_sweep$V.log; echo "sweep$V rc=$?"; cat gpurun_out/c7_sweep$V.jsonl; tail -3 gpurun_out/c7_sweep$V.log
done
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/c7_bench.json 2> gpurun_out/c7_bench.log; echo "bench rc=$?"; cat gpurun_out/c7_bench.json; tail -5 gpurun_out/c7_bench.log
