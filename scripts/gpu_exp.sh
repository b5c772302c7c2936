#!/bin/bash
mkdir -p gpurun_out
export IRIS_VERBOSE=1
bash scripts/ab_mel.sh base= fr6s3=gpurun_scratch/fr6s3/libiris.so 2>&1 | tail -6
IRIS_LIB=$PWD/gpurun_scratch/fr6s3/libiris.so timeout 200 python scripts/kb_mel.py 2>&1 | grep "CTAs/SM" | sort | uniq | head -3
