#!/bin/bash
mkdir -p gpurun_out
GZB_AR_LONG_MIN=64 timeout 600 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "hts or golden" > gpurun_out/c25_pytest.log 2>&1; echo "pytest (long decoder, threshold lowered) rc=$? $(tail -1 gpurun_out/c25_pytest.log)"
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 2 --cfg GZB_AR_LONG_MIN=off --cfg GZB_AR_LONG_MIN=262144,GZB_AR_LONG_ENT=16 --cfg GZB_AR_LONG_MIN=262144,GZB_AR_LONG_ENT=8 --cfg GZB_AR_LONG_MIN=262144,GZB_AR_LONG_ENT=32 2>&1 | tail -4 | cut -c1-520
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 1 --streams DIVRQUAL --cfg GZB_AR_LONG_MIN=262144,GZB_AR_LONG_ENT=16 --cfg GZB_AR_LONG_MIN=262144,GZB_AR_LONG_ENT=8 2>&1 | tail -2 | cut -c1-100,330-520
