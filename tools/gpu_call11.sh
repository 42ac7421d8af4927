#!/bin/bash
mkdir -p gpurun_out
for S in DIVRQUAL NONREF_X Q_X,Q_Y; do
timeout 600 python tools/sweep_fastq.py --vblocks 768 --steps 1 --streams $S --cfg GZB_AR_LONG_MIN=off --cfg GZB_AR_LONG_MIN=65536 --cfg GZB_AR_LONG_MIN=65536,GZB_AR_LONG_ENT=32 --cfg GZB_AR_LONG_MIN=off,GZB_AR_RUN4=0 --cfg GZB_AR_LONG_MIN=off,GZB_AR_SPLIT_MIN=off 2>&1 | tail -6 | cut -c1-600
done
timeout 600 python tools/sweep_fastq.py --vblocks 256 --steps 1 --streams DIVRQUAL --cfg GZB_AR_LONG_MIN=off --cfg GZB_AR_LONG_MIN=65536 2>&1 | tail -3 | cut -c1-600
