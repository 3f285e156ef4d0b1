#!/bin/bash
# ncu evidence for the bench command (run under gpurun; outputs land in gpurun_out/).  See B200_PROFILING.md.
set -x
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/ncu_launch_run.log 2>&1
for k in score_loss_kernel rowlist_apply_kernel gemm_simt_kernel attn_bwd_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 4 -c 2 -o gpurun_out/prof_$k -f $B > gpurun_out/ncu_$k.log 2>&1
done
