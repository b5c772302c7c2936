#!/bin/bash
# usage: scripts/kb_quick.sh tag [IRIS_LIB] [nokb] -- C=2 kernel timings + shared-memory wavefronts of k_fused<MEL>
tag=$1; lib=$2
export IRIS_LIB=$lib
if [ "$3" != "nokb" ]; then python scripts/kbench.py 256 10 2>&1 | grep "C=2" > gpurun_out/kb_$tag.log; fi
ncu --metrics l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_st.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed_op_shfl.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum \
  --clock-control none -k regex:k_fused -s 4 -c 1 --csv --log-file gpurun_out/ncu_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
echo "== $tag"; cat gpurun_out/kb_$tag.log; python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/ncu_$tag.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print('%-60s %s %s'%(d['Metric Name'],d['Metric Value'],d['Metric Unit']))
PY
