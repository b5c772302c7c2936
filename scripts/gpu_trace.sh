#!/bin/bash
mkdir -p gpurun_out
export TR=$PWD/gpurun_scratch/trace/libiris.so
{
for spec in "TRACE_MODE=lmm" "TRACE_MODE=mel"; do
  echo "=== $spec"
  env $spec IRIS_LIB=$TR timeout 300 python scripts/trace_fused.py 256 2>&1 | tail -40
done
} > gpurun_out/trace.txt 2>&1
grep -A6 "consumer warp" gpurun_out/trace.txt
