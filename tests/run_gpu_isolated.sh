#!/bin/bash
# runs every GPU test id in its own process (a CUDA fault is sticky for the whole process) and logs to gpurun_out/
mkdir -p gpurun_out
ids=$(python -m pytest tests -m gpu --collect-only -q 2>/dev/null | grep "::")
: > gpurun_out/isolated.log
for id in $ids; do
  timeout 600 python -m pytest "$id" -x -q > gpurun_out/one.log 2>&1
  rc=$?
  echo "$rc $id" >> gpurun_out/isolated.log
  if [ $rc -ne 0 ]; then echo "=== $id" >> gpurun_out/fail.log; grep -E "^E |Error|error" gpurun_out/one.log | head -12 >> gpurun_out/fail.log; fi
done
cat gpurun_out/isolated.log
