#!/bin/bash
# the entry points either side of the codec path on 32 VBlocks of 92 K reads x 150: ncu launch list of their kernels
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:"k_oq|k_hp|k_b250|k_local|k_normq|k_adler" --csv --log-file gpurun_out/r02_aux_kernels.csv \
  python tools/aux_kernels.py --vblocks 32 > gpurun_out/r02_aux_kernels.json 2> gpurun_out/c43.log; echo "rc=$?"
tail -1 gpurun_out/r02_aux_kernels.json | cut -c1-1500
python tools/ncu_table.py gpurun_out/r02_aux_kernels.csv | cut -c1-200
