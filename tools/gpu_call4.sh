#!/bin/bash
# round 2, GPU call 4: batch size of the lean FASTQ path, per-kernel launch list, full VCF / long-read lines
mkdir -p gpurun_out
run() { name=$1; shift; timeout 1500 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.log; echo "$name rc=$?"; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('   V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],2), 'piz', round(d['piz_GBps'],2), 'e2e', d['e2e'] and (round(d['e2e']['value'],2), round(d['e2e']['zip_ms']), round(d['e2e']['piz_ms'])), 'kern', {k: round(v) for k, v in d['roofline']['kernel_ms_per_step'].items()}, 'ms/step', round(d['ms_per_step'],1), 'cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'],3), d['cpu_baseline']['cores']))
except Exception as ex:
    print('   failed', ex); print(open('gpurun_out/$name.log').read()[-1500:])
P
}
run c4_v768   GZB_X=1 python bench.py --vblocks 768 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline
run c4_v1024  GZB_X=1 python bench.py --vblocks 1024 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --kernel-name regex:'k_' -c 600 --csv --log-file gpurun_out/c4_launches.csv python bench.py --vblocks 256 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/c4_ncu.log 2>&1; echo "ncu rc=$?"
run c4_auto   GZB_X=1 python bench.py --steps 3 --warmup 3
run c4_vcf    GZB_X=1 python bench.py --workload vcf --steps 3 --warmup 3
run c4_lr     GZB_X=1 python bench.py --workload longread --steps 2 --warmup 3
