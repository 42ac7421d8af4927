#!/bin/bash
# after the lane-0 stores of the order-0 arithmetic model: parity, racecheck again, and what it costs
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_hts.py tests/test_gpu_fastq.py tests/test_gpu_assign.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for t in hts fastq; do
  if [ $t = hts ]; then SEL="edge_sizes or soft_fail or packed_output or kinds"; F=tests/test_gpu_hts.py; else SEL="(domq or acgt) and not full_vb"; F="tests/test_gpu_fastq.py"; fi
  timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest $F -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/r02_sanitizer_${t}_racecheck.log 2>&1
  echo "racecheck $t rc=$? $(grep -E 'RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_${t}_racecheck.log | tr '\n' ' ')"
done
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 2 --cfg "" 2>&1 | tail -1 | cut -c1-700
