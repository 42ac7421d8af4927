#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python bench.py > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.log; echo "bench rc=$?"; cat gpurun_out/c13_bench.json; tail -5 gpurun_out/c13_bench.log
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/c13_ref.json 2> gpurun_out/c13_ref.log; echo "ref rc=$?"; cut -c1-600 gpurun_out/c13_ref.json; tail -3 gpurun_out/c13_ref.log
