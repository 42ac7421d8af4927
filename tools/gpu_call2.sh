#!/bin/bash
# round 2, GPU call 2: parity of everything, then first timings of the new PBWT / LONGR kernels (probe sizes)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c2_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c2_pytest.log)"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_vcf_longr.py -m gpu -x -q -k "not config_size" -p no:cacheprovider > gpurun_out/c2_memcheck.log 2>&1; echo "memcheck rc=$? $(tail -2 gpurun_out/c2_memcheck.log | tr '\n' ' ')"
timeout 600 python bench.py --workload vcf --vblocks 148 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/c2_vcf148.json 2> gpurun_out/c2_vcf148.log; echo "vcf148 rc=$?"
timeout 600 python bench.py --workload vcf --vblocks 296 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/c2_vcf296.json 2> gpurun_out/c2_vcf296.log; echo "vcf296 rc=$?"
timeout 900 python bench.py --workload longread --vblocks 296 --lr-bases 2000000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/c2_lr296.json 2> gpurun_out/c2_lr296.log; echo "lr296 rc=$?"
timeout 900 python bench.py --workload longread --vblocks 1184 --lr-bases 2000000 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/c2_lr1184.json 2> gpurun_out/c2_lr1184.log; echo "lr1184 rc=$?"
for f in c2_vcf148 c2_vcf296 c2_lr296 c2_lr1184; do python - <<P
import json
try:
    d=json.loads(open('gpurun_out/$f.json').read().strip().splitlines()[-1])
    print('$f', 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],2), 'piz', round(d['piz_GBps'],2), 'e2e', d['e2e'] and round(d['e2e']['value'],2), 'kernel ms', d['roofline']['kernel_ms_per_step'], 'ms/step', round(d['ms_per_step'],1))
except Exception as ex:
    print('$f failed', ex); print(open('gpurun_out/$f.log').read()[-1500:])
P
done
