#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -2
bash scripts/ab_mel.sh new= base=gpurun_scratch/base/libiris.so 2>&1 | tail -6
for spec in "IRIS_X=new" "IRIS_LIB=$PWD/gpurun_scratch/base/libiris.so" "IRIS_X=new2" "IRIS_LIB=$PWD/gpurun_scratch/base/libiris.so"; do
  env $spec timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-e2e --no-cfg3 --no-consumer-check > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_x.json'))
print('$spec value %.0f ms/step %.4f frac %.4f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms']))
PY
done
