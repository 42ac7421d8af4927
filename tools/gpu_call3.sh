#!/bin/bash
# round 2, GPU call 3: the lean FASTQ path — parity, then the batch size and the split encoder A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c3_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c3_pytest.log)"
run() { name=$1; shift; timeout 900 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.log; echo "$name rc=$?"; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('   V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],2), 'piz', round(d['piz_GBps'],2), 'e2e', d['e2e'] and (round(d['e2e']['value'],2), round(d['e2e']['zip_ms']), round(d['e2e']['piz_ms'])), 'kern', {k: round(v) for k, v in d['roofline']['kernel_ms_per_step'].items()}, 'ms/step', round(d['ms_per_step'],1), 'cpu', d['cpu_baseline'] and round(d['cpu_baseline']['value'],3))
except Exception as ex:
    print('   failed', ex); print(open('gpurun_out/$name.log').read()[-1200:])
P
}
run c3_v512_split    GZB_X=1 python bench.py --vblocks 512 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline
run c3_v512_nosplit  GZB_AR_SPLIT_MIN=off python bench.py --vblocks 512 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline
run c3_v1024_split   GZB_X=1 python bench.py --vblocks 1024 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline
run c3_auto          GZB_X=1 python bench.py --steps 3 --warmup 3
nvidia-smi --query-gpu=memory.used,memory.total --format=csv; free -g | head -2
