#!/bin/bash
# round 2, GPU call 9: bandwidth passes before chains (device + host mode), DOMQ line kernels + single transfers
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c9_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c9_pytest.log)"
for V in 768 1024; do
timeout 900 python tools/sweep_fastq.py --vblocks $V --steps 2 --cfg "" --cfg GZB_AR_LONG_MIN=off --cfg GZB_AR_LONG_MIN=off,GZB_AR_CTAS=2,GZB_AR0_CTAS=2 --cfg GZB_AR_LONG_MIN=off,GZB_AR_CTAS=8,GZB_AR0_CTAS=8 \
  > gpurun_out/c9_sweep$V.jsonl 2> gpurun_out/c9_sweep$V.log; echo "sweep$V rc=$?"; cat gpurun_out/c9_sweep$V.jsonl; tail -3 gpurun_out/c9_sweep$V.log
done
timeout 600 python tools/timeline.py --vblocks 768 --mode device --steps 1 > gpurun_out/c9_timeline_dev.txt 2>&1; tail -32 gpurun_out/c9_timeline_dev.txt
timeout 600 python tools/timeline.py --vblocks 512 --mode host --steps 1 > gpurun_out/c9_timeline_host.txt 2>&1; tail -45 gpurun_out/c9_timeline_host.txt
for G in 1 2; do
GZB_AR_LONG_MIN=off timeout 900 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-groups $G > gpurun_out/c9_bench_g$G.json 2> gpurun_out/c9_bench_g$G.log; echo "bench G=$G rc=$?"
python - <<P
import json
try:
    d=json.loads(open('gpurun_out/c9_bench_g$G.json').read().strip().splitlines()[-1]); print('G=$G V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],1), 'piz', round(d['piz_GBps'],1), 'e2e', d['e2e'])
except Exception as ex:
    print('failed', ex); print(open('gpurun_out/c9_bench_g$G.log').read()[-1500:])
P
done
