#!/bin/bash
# one full-set capture each of the DOMQ histogram and normalise passes (128 VBlocks), raw page exported on the box
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"linehist|k_domq_normalize" --launch-skip 2 -c 2 -o /tmp/dq_full \
  python tools/sweep_fastq.py --vblocks 128 --steps 1 --cfg "" > gpurun_out/c33_ncu.log 2>&1
ncu -i /tmp/dq_full.ncu-rep --page raw --csv > gpurun_out/r02_domq_full_raw.csv 2>/dev/null
ncu -i /tmp/dq_full.ncu-rep --page source --csv -k regex:linehist > gpurun_out/r02_linehist_source.csv 2>/dev/null
ls -la gpurun_out/r02_domq_full_raw.csv gpurun_out/r02_linehist_source.csv
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 2 --cfg "" 2>&1 | tail -1 | cut -c1-200
