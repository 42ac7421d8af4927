#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/c14_bench.json 2> gpurun_out/c14_bench.log; echo "bench rc=$?"; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/c14_bench.json').read().strip().splitlines()[-1]); print('V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],1), 'piz', round(d['piz_GBps'],1), 'e2e', d['e2e'])
except Exception as ex:
    print('failed', ex); print(open('gpurun_out/c14_bench.log').read()[-1500:])
P
