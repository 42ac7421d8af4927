#!/bin/bash
# OQ on the GPU (parity, racecheck, memcheck), then the e2e timeline with the new DOMQ passes
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_oq.py tests/test_gpu_assign.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for tool in racecheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_oq.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_sanitizer_oq_$tool.log 2>&1
  echo "$tool oq rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_oq_$tool.log | tr '\n' ' ')"
done
timeout 300 python tools/e2e_probe.py > gpurun_out/c37_e2e_probe.txt 2>&1
grep -m16 "zip_steps\|piz_steps\|zip_device\|piz_device\|domq_prepare\|domq_split\|acgt_pack\|domq_reconstruct" gpurun_out/c37_e2e_probe.txt
