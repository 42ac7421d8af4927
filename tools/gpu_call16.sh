#!/bin/bash
# round 2, GPU call 16: parity, then the three workloads' bench lines as the driver runs them
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/c16_pytest.log 2>&1; echo "pytest rc=$? $(tail -1 gpurun_out/c16_pytest.log)"
timeout 1200 python bench.py > gpurun_out/c16_bench_fastq.json 2> gpurun_out/c16_bench_fastq.log; echo "fastq rc=$?"; cat gpurun_out/c16_bench_fastq.json; tail -3 gpurun_out/c16_bench_fastq.log
timeout 900 python bench.py --workload vcf --steps 3 --warmup 3 > gpurun_out/c16_bench_vcf.json 2> gpurun_out/c16_bench_vcf.log; echo "vcf rc=$?"; cat gpurun_out/c16_bench_vcf.json; tail -3 gpurun_out/c16_bench_vcf.log
timeout 900 python bench.py --workload longread --steps 2 --warmup 3 > gpurun_out/c16_bench_lr.json 2> gpurun_out/c16_bench_lr.log; echo "longread rc=$?"; cat gpurun_out/c16_bench_lr.json; tail -3 gpurun_out/c16_bench_lr.log
