#!/bin/bash
# round 2, GPU call 5: plug-in harness on the GPU, the vectorised PBWT rows, LONGR decode with L2 hints, the FASTQ line, ncu --set full captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c5_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c5_pytest.log)"
run() { name=$1; shift; timeout 1500 env "$@" > gpurun_out/$name.json 2> gpurun_out/$name.log; echo "$name rc=$?"; python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$name.json').read().strip().splitlines()[-1])
    print('   V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],2), 'piz', round(d['piz_GBps'],2), 'e2e', d['e2e'] and (round(d['e2e']['value'],2), round(d['e2e']['zip_ms']), round(d['e2e']['piz_ms'])), 'kern', {k: round(v) for k, v in d['roofline']['kernel_ms_per_step'].items()}, 'ms/step', round(d['ms_per_step'],1), 'cpu', d['cpu_baseline'] and (round(d['cpu_baseline']['value'],3), d['cpu_baseline']['cores']))
except Exception as ex:
    print('   failed', ex); print(open('gpurun_out/$name.log').read()[-1500:])
P
}
run c5_vcf148  GZB_X=1 python bench.py --workload vcf --vblocks 148 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline
run c5_vcf     GZB_X=1 python bench.py --workload vcf --steps 3 --warmup 3
run c5_lr296   GZB_X=1 python bench.py --workload longread --vblocks 296 --lr-bases 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
run c5_lr1184  GZB_X=1 python bench.py --workload longread --vblocks 1184 --lr-bases 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e
run c5_fastq   GZB_X=1 python bench.py --steps 3 --warmup 3
NCU="ncu --set full --clock-control none --import-source on"
timeout 600 $NCU -k regex:'k_arith_decode_t|k_arith_encode_t|k_ar_split|k_rans' -c 12 -o gpurun_out/r02_fastq64 python bench.py --vblocks 64 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/c5_ncu_fastq.log 2>&1; echo "ncu fastq rc=$?"
timeout 600 $NCU -k regex:'k_pbwt_rows' -c 2 -o gpurun_out/r02_vcf148 python bench.py --workload vcf --vblocks 148 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/c5_ncu_vcf.log 2>&1; echo "ncu vcf rc=$?"
timeout 600 $NCU -k regex:'k_longr_channels|k_longr_decode|k_longr_place' -c 3 -o gpurun_out/r02_lr296 python bench.py --workload longread --vblocks 296 --lr-bases 500000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/c5_ncu_lr.log 2>&1; echo "ncu lr rc=$?"
ls -la gpurun_out/*.ncu-rep
