#!/bin/bash
# one gpurun call: GPU tests, time-stamp traces of k_fused, A/B of the claim schedule
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/tests.log
export TR=$PWD/gpurun_scratch/trace/libiris.so
{
for spec in "IRIS_TAIL1=0 IRIS_TAIL2=0" "IRIS_TAIL1=2 IRIS_TAIL2=4" "IRIS_TAIL1=2 IRIS_TAIL2=8" "IRIS_TAIL1=4 IRIS_TAIL2=8 TRACE_MODE=mel"; do
  echo "=== $spec"
  env $spec IRIS_LIB=$TR timeout 300 python scripts/trace_fused.py 256 2>&1 | tail -32
done
} > gpurun_out/trace.txt 2>&1
bash scripts/ab_mel.sh t00=,IRIS_TAIL1=0,IRIS_TAIL2=0 t13=,IRIS_TAIL1=1,IRIS_TAIL2=3 t24=,IRIS_TAIL1=2,IRIS_TAIL2=4 t28=,IRIS_TAIL1=2,IRIS_TAIL2=8 t48=,IRIS_TAIL1=4,IRIS_TAIL2=8 t216=,IRIS_TAIL1=2,IRIS_TAIL2=16
for late in 0 1; do
  IRIS_METRIC_LATE=$late timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_late$late.json 2> gpurun_out/bench_late$late.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_late$late.json'))
print('late=$late value %.0f ms/step %.4f frac %.3f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms']))
PY
done
