#!/bin/bash
# usage: scripts/scale_run.sh N...  -- bench.py at each N on one box (torchrun for N > 1), lines -> gpurun_out/scale_nN.json
mkdir -p gpurun_out
for n in "$@"; do
  if [ "$n" = 1 ]; then
    timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n rc=$?"; grep -o '"value": [0-9.]*' gpurun_out/scale_n$n.json | head -2; tail -2 gpurun_out/scale_n$n.err
done
