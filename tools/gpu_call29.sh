#!/bin/bash
# DOMQ / ACGT bandwidth kernels alone under ncu (times are cold-cache, serialised), then the FASTQ sweep at bench size
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
  --clock-control none -k regex:"domq|dqs|dqp|acgt" -c 60 --csv --log-file gpurun_out/r02_domq_kernels.csv \
  python tools/sweep_fastq.py --vblocks 128 --steps 1 --cfg "" > gpurun_out/c29_ncu.log 2>&1
python tools/ncu_table.py gpurun_out/r02_domq_kernels.csv 2>/dev/null | head -30
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 2 --cfg "" 2>&1 | tail -1 | cut -c1-200
