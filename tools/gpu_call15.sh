#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/e2e_probe.py --vblocks 768 --steps 3 > gpurun_out/c15_e2e_probe.txt 2>&1; cat gpurun_out/c15_e2e_probe.txt | tail -40
