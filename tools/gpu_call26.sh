#!/bin/bash
# round 2, GPU call 26: HEAD as the driver will run it — build check, smoke, parity, the default bench, the reference arm
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c26_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c26_pytest.log)"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c26_ref.json 2> gpurun_out/c26_ref.log; echo "ref rc=$?"; cut -c1-260 gpurun_out/c26_ref.json
timeout 1200 python bench.py > gpurun_out/c26_bench.json 2> gpurun_out/c26_bench.log; echo "bench rc=$?"; python - <<P
import json
d=json.loads(open('gpurun_out/c26_bench.json').read().strip().splitlines()[-1])
print('V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],1), 'piz', round(d['piz_GBps'],1), 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='how'}, 'cpu', round(d['cpu_baseline']['value'],3), 'launches', d['gpu_launches'], 'traffic', d['roofline']['traffic'])
P
