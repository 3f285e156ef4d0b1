# round-2 ncu evidence for the default workload on one GPU (run under gpurun; CSV summaries land in gpurun_out/, the .ncu-rep stays on the box)
set -x
python -m pytest tests/test_gpu_ops.py tests/test_gpu_models.py -q -x 2>&1 | tail -2
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2y_default_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 1600 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_launches_bench.log 2>&1
ncu --set full --clock-control none -k regex:"score_loss_v3|rowlist_apply|gemm_tc|gemm_simt_small|attn_|add_ln|seq_prep" --launch-skip 170 --launch-count 50 \
    -o /tmp/r02_full python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-eager-baseline > gpurun_out/r02_full_bench.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > gpurun_out/r02_full_raw.csv
ls -la /tmp/r02_full.ncu-rep gpurun_out/ | tail -8
