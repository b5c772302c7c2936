#!/bin/bash
# bench.py at N GPUs of this box (N = number of visible GPUs, or $1), plus the NCCL parity test.
N=${1:-$(nvidia-smi -L | wc -l)}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
if [ "$N" -ge 2 ]; then
  timeout 600 python -m pytest tests/test_gpu_step.py -q -m gpu -k nccl 2>&1 | tail -3 | tee gpurun_out/nccl_test_n$N.log
fi
for n in ${SCALE_NS:-$N}; do
  if [ "$n" -eq 1 ]; then
    timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
      bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  tail -3 gpurun_out/scale_n$n.err
  python - <<PY
import json
try:
    d = json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1])
    e = d.get('e2e', {})
    print('N=%d value %.0f ms/step %.4f frac %.3f kernel_ms %.4f share %.3f e2e %.0f dev %.0f d2h/rank %.1f GB/s cpu %s' % (
        d['n_gpus'], d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel_ms'],
        d['roofline']['kernel_share_of_step'], e.get('value', 0), e.get('features_on_device_value', 0),
        e.get('d2h_gb_per_s_per_rank', 0), d.get('cpu_baseline', {}).get('value')))
except Exception as ex:
    print('no line', ex)
PY
done
