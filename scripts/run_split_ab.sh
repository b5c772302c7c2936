#!/bin/bash
# A/B of the split min-max log-mel launch (IRIS_SPLIT = clips per part, 0 = one launch) at batch 1024
# (= one of 8 shards of BASELINE configs[3]) and 8192 (all of it on one GPU).
show='
import json,sys
d=json.loads(sys.stdin.read())
c3=d.get("configs3_one_gpu") or {"value":0,"ms_per_step":0}
print("value %.0f ms/step %.4f frac %.3f kernel_ms %.4f clips/launch %d step_frac %.3f cfg3 %.0f (%.3f ms)" % (d["value"], d["ms_per_step"], d["roofline"]["frac"], d["roofline"]["kernel_ms"], d["roofline"]["clips_per_launch"], d["roofline"]["step_frac"], c3["value"], c3["ms_per_step"]))'
for sp in 0 256 512 128; do
  echo "B=1024 IRIS_SPLIT=$sp"; IRIS_SPLIT=$sp python bench.py --batch 1024 --steps 10 --warmup 5 --no-cpu-baseline --no-e2e --no-cfg3 | python -c "$show"
done
for sp in 0 256 512; do
  echo "B=256 + cfg3 IRIS_SPLIT=$sp"; IRIS_SPLIT=$sp python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-e2e | python -c "$show"
done
for ch in 1 2 4 8; do
  echo "B=256 IRIS_CHUNK=$ch"; IRIS_CHUNK=$ch python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-cfg3 | python -c "$show"
done
