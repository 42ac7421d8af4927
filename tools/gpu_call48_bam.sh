#!/bin/bash
# NOT RUN (the round's GPU minutes were spent): what the next GPU call for the BAM workload (BASELINE configs[2]) would be —
# parity at both sizes, one bench line of each arm, a launch list of the same command.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_zz_gpu_bam_path.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/c48_bam_pytest.log
timeout 600 python bench.py --impl reference --workload bam --steps 2 --warmup 1 > gpurun_out/c48_bam_ref.json 2> gpurun_out/c48_bam_ref.log
timeout 900 python bench.py --workload bam --steps 5 --warmup 3 > gpurun_out/c48_bam_bench.json 2> gpurun_out/c48_bam_bench.log; echo "rc=$?"
tail -1 gpurun_out/c48_bam_bench.json | cut -c1-1500
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bam_v64.csv \
  python bench.py --workload bam --vblocks 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2> gpurun_out/c48_bam_ncu.log
python tools/ncu_table.py gpurun_out/r02_launches_bam_v64.csv | cut -c1-200
