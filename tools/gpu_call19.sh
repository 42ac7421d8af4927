#!/bin/bash
# round 2, GPU call 19: parity on the final sources (NORMQ, digest, staging), DOMQ sub-batch size, LONGR without the L2 hints
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c19_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c19_pytest.log)"
for SB in 128 256 384; do timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 2 --sub-batch $SB --cfg "" 2>&1 | tail -1 | cut -c1-400; done
timeout 900 python bench.py --workload longread --steps 1 --warmup 3 --no-e2e > gpurun_out/c19_bench_lr.json 2> gpurun_out/c19_bench_lr.log; echo "longread rc=$?"; cut -c1-330 gpurun_out/c19_bench_lr.json; tail -2 gpurun_out/c19_bench_lr.log
timeout 900 ncu --set full --clock-control none -k regex:'k_longr_channels|k_longr_decode|k_longr_place' -c 3 -o gpurun_out/r02_lr296 -f python bench.py --workload longread --vblocks 296 --lr-bases 500000 --steps 1 --warmup 0 --no-e2e --no-cpu-baseline > gpurun_out/c19_lr.log 2>&1; echo "ncu lr rc=$?"
ncu -i gpurun_out/r02_lr296.ncu-rep --page raw --csv > gpurun_out/r02_lr296_raw.csv 2>/dev/null; rm -f gpurun_out/r02_lr296.ncu-rep
