#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/tests.log
bash scripts/ab_mel.sh base=
for spec in "IRIS_X=0" "IRIS_X=1"; do
  env $spec timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e > gpurun_out/bench_x.json 2> gpurun_out/bench_x.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_x.json'))
print('$spec value %.0f ms/step %.4f frac %.3f kernel_ms %.4f' % (d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms']))
PY
done
