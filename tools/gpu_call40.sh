#!/bin/bash
# round 2, GPU call 40: the launch list of a FASTQ step on the final build (the DOMQ kernels changed), then the two bench arms as the driver runs them
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_fastq_v64.csv python tools/sweep_fastq.py --vblocks 64 --steps 1 --cfg "" > gpurun_out/c40_insts.log 2>&1; echo "ncu fastq launch list rc=$?"
python tools/ncu_table.py gpurun_out/r02_launches_fastq_v64.csv > gpurun_out/r02_launches_fastq_v64.md; grep "k_" gpurun_out/r02_launches_fastq_v64.md | head -12 | cut -c1-170
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c40_ref.json 2> gpurun_out/c40_ref.log; echo "ref rc=$?"; cut -c1-260 gpurun_out/c40_ref.json
timeout 1200 python bench.py > gpurun_out/c40_bench.json 2> gpurun_out/c40_bench.log; echo "bench rc=$?"; python - <<P
import json
d=json.loads(open('gpurun_out/c40_bench.json').read().strip().splitlines()[-1])
print('V', d['config']['vblocks_per_gpu_per_step'], 'value', round(d['value'],2), 'zip', round(d['zip_GBps'],1), 'piz', round(d['piz_GBps'],1), 'e2e', {k:(round(v,1) if isinstance(v,float) else v) for k,v in d['e2e'].items() if k!='how'}, 'cpu', round(d['cpu_baseline']['value'],3), 'launches', d['gpu_launches'], 'roofline', d['roofline'])
P
