#!/bin/bash
# b250 / transposes / OQ on the GPU: parity, then racecheck + memcheck of the new kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_b250.py tests/test_local_transpose.py tests/test_oq.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for tool in racecheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_b250.py tests/test_local_transpose.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_sanitizer_b250_$tool.log 2>&1
  echo "$tool b250+transpose rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_b250_$tool.log | tr '\n' ' ')"
done
