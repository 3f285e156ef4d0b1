# round-2 ncu evidence for the default workload on one GPU (run under gpurun; CSV summaries land in gpurun_out/, the .ncu-rep stays on the box)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none -k regex:"score_loss_v3|rowlist_apply|gemm_tc|attn_|add_ln|seq_prep" --launch-skip 150 --launch-count 40 \
    -o /tmp/r02_full python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_full_bench.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv
ls -la /tmp/r02_full.ncu-rep gpurun_out/
