#!/bin/bash
# round 2, GPU call 18: the evidence pass on the final build — launch list + instruction counts of a FASTQ step, ncu --set full of the chain kernels, PBWT and LONGR kernels
mkdir -p gpurun_out
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active
timeout 900 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r02_launches_fastq_v64.csv python tools/sweep_fastq.py --vblocks 64 --steps 1 --cfg "" > gpurun_out/c18_insts.log 2>&1; echo "ncu fastq launch list rc=$?"
python tools/ncu_table.py gpurun_out/r02_launches_fastq_v64.csv > gpurun_out/r02_launches_fastq_v64.md; head -16 gpurun_out/r02_launches_fastq_v64.md | cut -c1-200
timeout 1500 ncu --set full --clock-control none -k regex:'k_arith_decode_t|k_arith_encode_t|k_ar_split|k_rans_encode|k_rans_decode' --launch-skip 20 -c 9 -o gpurun_out/r02_fastq64 -f python tools/sweep_fastq.py --vblocks 64 --steps 1 --cfg "" > gpurun_out/c18_full.log 2>&1; echo "ncu fastq full rc=$?"
ncu -i gpurun_out/r02_fastq64.ncu-rep --page raw --csv > gpurun_out/r02_fastq64_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'k_pbwt_rows|k_pbwt_emit' -c 4 -o gpurun_out/r02_vcf148 -f python bench.py --workload vcf --vblocks 148 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/c18_vcf.log 2>&1; echo "ncu vcf rc=$?"
ncu -i gpurun_out/r02_vcf148.ncu-rep --page raw --csv > gpurun_out/r02_vcf148_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none -k regex:'k_longr_channels|k_longr_decode|k_longr_place' -c 3 -o gpurun_out/r02_lr296 -f python bench.py --workload longread --vblocks 296 --lr-bases 500000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/c18_lr.log 2>&1; echo "ncu lr rc=$?"
ncu -i gpurun_out/r02_lr296.ncu-rep --page raw --csv > gpurun_out/r02_lr296_raw.csv 2>/dev/null
rm -f gpurun_out/r02_fastq64.ncu-rep gpurun_out/r02_vcf148.ncu-rep gpurun_out/r02_lr296.ncu-rep     # (the raw pages are what is kept; the reports exceed what comes back)
ls -la gpurun_out/r02_*raw.csv
