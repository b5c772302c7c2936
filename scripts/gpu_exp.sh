#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_step.py -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/tests.log
for spec in "IRIS_WARM=0" "IRIS_WARM=1" "IRIS_WARM=2" "IRIS_WARM=0" "IRIS_WARM=1" "IRIS_WARM=4"; do
  env $spec timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_x.json'))
print('$spec value %.0f ms/step %.4f frac %.3f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms']))
PY
done
