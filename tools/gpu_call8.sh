#!/bin/bash
mkdir -p gpurun_out
./tools/probes/smsp_probe > gpurun_out/c8_smsp.txt 2>&1; cat gpurun_out/c8_smsp.txt
timeout 600 python tools/timeline.py --vblocks 768 --mode device --steps 2 > gpurun_out/c8_timeline.txt 2>&1; tail -60 gpurun_out/c8_timeline.txt
