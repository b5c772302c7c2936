#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench, ncu launch list, ncu full capture.
# usage: scripts/gpu_round.sh [tests] [smoke] [bench] [launches] [ncu]
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/gpu.txt 2>&1
for what in "$@"; do
case $what in
tests)
  timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/tests.log ;;
smoke)
  timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log ;;
bench)
  timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err
  tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json ;;
refbench)
  timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  cat gpurun_out/bench_ref.json ;;
launches)
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
     --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 --no-consumer-check \
     > gpurun_out/launches_bench.log 2>&1
  tail -2 gpurun_out/launches_bench.log ;;
ncu)
  timeout 1200 ncu --set full --clock-control none --import-source on -k regex:k_fused -s 4 -c 2 \
     -f -o gpurun_out/prof_fused python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-cfg3 --no-consumer-check \
     > gpurun_out/ncu_bench.log 2>&1
  tail -2 gpurun_out/ncu_bench.log ;;
kbench)
  timeout 600 python scripts/kbench.py 256 10 2>&1 | grep "C=" | tee gpurun_out/kbench.log ;;
esac
done
