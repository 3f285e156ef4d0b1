#!/bin/bash
# ncu evidence for the bench command (run under gpurun; outputs land in gpurun_out/).  See B200_PROFILING.md.
# Numbers printed by bench.py under ncu are never bench values.  --no-graph keeps the step eager so that every launch is listed
# (the whole-step CUDA graph replays the same kernels).
set -x
TAG=${1:-r01d}
B="python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph"
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/${TAG}_launches.csv $B > gpurun_out/${TAG}_ncu_launch_run.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:score_loss_v3_kernel -s 6 -c 2 -o gpurun_out/${TAG}_prof_score_loss_v3_kernel -f $B > gpurun_out/${TAG}_ncu_score.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rowlist_apply_kernel -s 6 -c 2 -o gpurun_out/${TAG}_prof_rowlist_apply_kernel -f $B > gpurun_out/${TAG}_ncu_apply.log 2>&1
# steady-state step: skip the first 8 steps' GEMMs (30 per step), then capture one full step's worth
ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 240 -c 30 -o gpurun_out/${TAG}_prof_gemm_tc_kernel -f $B > gpurun_out/${TAG}_ncu_gemm.log 2>&1
