#!/bin/bash
# round 2, GPU call 21: prefix-accepting 4-step speculation (GZB_AR_RUN4 0 / 1 / 2), spin-wait staging again; parity first
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c21_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c21_pytest.log)"
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 1 --streams DIVRQUAL --cfg GZB_AR_RUN4=2 --cfg GZB_AR_RUN4=1 --cfg GZB_AR_RUN4=0 2>&1 | tail -3 | cut -c1-330
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 2 --cfg GZB_AR_RUN4=2 --cfg GZB_AR_RUN4=1 2>&1 | tail -2 | cut -c1-500
timeout 1200 python bench.py > gpurun_out/c21_bench_fastq.json 2> gpurun_out/c21_bench_fastq.log; echo "fastq rc=$?"; python - <<P
import json
d=json.loads(open('gpurun_out/c21_bench_fastq.json').read().strip().splitlines()[-1])
print('V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],1), 'piz', round(d['piz_GBps'],1), 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='how'}, 'cpu', round(d['cpu_baseline']['value'],3), d['config'].get('cpu_binding'))
P
