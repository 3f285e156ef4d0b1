# round-2 ncu evidence for the default workload on one GPU (run under gpurun; results land in gpurun_out/)
set -x
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"score_loss_v3|rowlist_apply|gemm_tc|attn_|add_ln|seq_prep" --launch-skip 150 --launch-count 50 \
    -o gpurun_out/r02_full python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_full_bench.log 2>&1
ls -la gpurun_out/r02_full.ncu-rep gpurun_out/r02_launches.csv
