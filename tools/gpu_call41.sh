#!/bin/bash
# SMUX / TMPL / PACB on the GPU: parity, sanitizer; then smoke() and the whole GPU suite on the final build
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_smux.py tests/test_tmpl.py tests/test_pacb.py tests/test_oq.py -m gpu -x -q -p no:cacheprovider 2>&1 | tail -2
for tool in racecheck memcheck; do
  timeout 300 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_smux.py tests/test_tmpl.py tests/test_pacb.py tests/test_oq.py -m gpu -x -q -p no:cacheprovider > gpurun_out/r02_sanitizer_mux_$tool.log 2>&1
  echo "$tool mux rc=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' gpurun_out/r02_sanitizer_mux_$tool.log | tr '\n' ' ')"
done
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c41_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c41_pytest.log)"
