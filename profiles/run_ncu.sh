#!/bin/bash
# ncu evidence for the bench command (run under gpurun; outputs land in gpurun_out/).  See B200_PROFILING.md.
# Numbers printed by bench.py under ncu are never bench values.
set -x
TAG=${1:-r01c}
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv $B --no-graph > gpurun_out/${TAG}_ncu_launch_run.log 2>&1
for k in score_loss_v3_kernel rowlist_apply_kernel gemm_tc_kernel attn_bwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 6 -c 3 -o gpurun_out/${TAG}_prof_$k -f $B > gpurun_out/${TAG}_ncu_$k.log 2>&1
done
