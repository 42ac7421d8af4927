#!/bin/bash
# round 2, GPU call 6: persistent chain kernels (CTAs per SM), split encoder on its own stream, run4 decoder: parity, then the sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c6_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c6_pytest.log)"
timeout 900 python tools/sweep_fastq.py --vblocks 512 --steps 2 \
  --cfg "" --cfg GZB_AR_CTAS=16,GZB_AR0_CTAS=16 --cfg GZB_AR_CTAS=16,GZB_AR0_CTAS=16,GZB_AR_SPLIT_STREAM=0,GZB_AR_RUN4=0 \
  --cfg GZB_AR_CTAS=2 --cfg GZB_AR_CTAS=3 --cfg GZB_AR_CTAS=6 --cfg GZB_AR_CTAS=8 \
  --cfg GZB_AR_CTAS=4,GZB_AR0_CTAS=2 --cfg GZB_AR_CTAS=4,GZB_AR0_CTAS=8 --cfg GZB_AR_CTAS=2,GZB_AR0_CTAS=2 \
  --cfg GZB_AR_RUN4=0 --cfg GZB_AR_SPLIT_STREAM=0 > gpurun_out/c6_sweep512.jsonl 2> gpurun_out/c6_sweep512.log; echo "sweep512 rc=$?"; cat gpurun_out/c6_sweep512.jsonl; tail -3 gpurun_out/c6_sweep512.log
timeout 900 python tools/sweep_fastq.py --vblocks 768 --steps 2 \
  --cfg "" --cfg GZB_AR_CTAS=16,GZB_AR0_CTAS=16 --cfg GZB_AR_CTAS=3 --cfg GZB_AR_CTAS=6 --cfg GZB_AR_CTAS=8 --cfg GZB_AR_CTAS=6,GZB_AR0_CTAS=2 \
  > gpurun_out/c6_sweep768.jsonl 2> gpurun_out/c6_sweep768.log; echo "sweep768 rc=$?"; cat gpurun_out/c6_sweep768.jsonl; tail -3 gpurun_out/c6_sweep768.log
timeout 600 python bench.py --workload longread --vblocks 296 --steps 1 --warmup 1 --no-e2e > gpurun_out/c6_lr_cpu.json 2> gpurun_out/c6_lr_cpu.log; echo "lr rc=$?"; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/c6_lr_cpu.json').read().strip().splitlines()[-1]); print('lr value', d['value'], 'zip', d['zip_GBps'], 'piz', d['piz_GBps'], 'cpu', d['cpu_baseline'])
except Exception as ex:
    print('lr failed', ex); print(open('gpurun_out/c6_lr_cpu.log').read()[-1500:])
P
