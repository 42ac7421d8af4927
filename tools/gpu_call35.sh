#!/bin/bash
# compute-sanitizer over the DOMQ / ACGT kernels and the simple codecs' edge cases: memcheck, racecheck, synccheck
mkdir -p gpurun_out
SEL_FQ='(domq or acgt) and not full_vb'
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_fastq.py tests/test_gpu_assign.py -m gpu -x -q -k "$SEL_FQ or assign" -p no:cacheprovider > gpurun_out/r02_sanitizer_fastq_$tool.log 2>&1
  echo "$tool fastq rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_fastq_$tool.log | tr '\n' ' ')"
done
for tool in memcheck racecheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_hts.py -m gpu -x -q -k "edge_sizes or soft_fail or packed_output" -p no:cacheprovider > gpurun_out/r02_sanitizer_hts_$tool.log 2>&1
  echo "$tool hts rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_hts_$tool.log | tr '\n' ' ')"
done
