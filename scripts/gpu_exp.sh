#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/tests.log
export IRIS_TAIL1=1 IRIS_TAIL2=3
bash scripts/ab_mel.sh base= nored=gpurun_scratch/nored/libiris.so nomm=gpurun_scratch/nomm/libiris.so nohint=gpurun_scratch/nohint/libiris.so
