#!/bin/bash
# round 2, GPU call 10: instruction counts of every kernel of a zip + piz step (one cheap ncu pass), then --set full of the chain kernels
mkdir -p gpurun_out
export GZB_AR_LONG_MIN=off
timeout 900 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,launch__block_size,launch__registers_per_thread,sm__warps_active.avg.pct_of_peak_sustained_active \
  --clock-control none --csv --log-file gpurun_out/r02_insts_v64.csv python tools/sweep_fastq.py --vblocks 64 --steps 1 --cfg "" > gpurun_out/c10_insts.log 2>&1; echo "ncu insts rc=$?"
python tools/ncu_table.py gpurun_out/r02_insts_v64.csv | tee gpurun_out/r02_insts_v64.md | head -70
timeout 1500 ncu --set full --clock-control none -k regex:'k_arith_decode_t|k_arith_encode_t|k_ar_split|k_rans_encode|k_rans_decode' --launch-skip 20 -c 9 -o gpurun_out/r02_fastq64 python tools/sweep_fastq.py --vblocks 64 --steps 1 --cfg "" > gpurun_out/c10_full.log 2>&1; echo "ncu full rc=$?"
ls -la gpurun_out/*.ncu-rep
ncu -i gpurun_out/r02_fastq64.ncu-rep --page raw --csv > gpurun_out/r02_fastq64_raw.csv 2>/dev/null; wc -c gpurun_out/r02_fastq64_raw.csv
