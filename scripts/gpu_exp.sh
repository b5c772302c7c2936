#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/tests.log
timeout 600 python scripts/kbench.py 256 10 4 2>&1 | grep "C=" | tee gpurun_out/kbench_c4.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python scripts/sanitize_small.py > gpurun_out/sanitizer_$tool.log 2>&1
  tail -4 gpurun_out/sanitizer_$tool.log
done
