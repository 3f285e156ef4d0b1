# final single-GPU evidence of round 2: full GPU test suite, smoke, bench lines of every single-GPU configuration
set -x
python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2z_pytest_gpu.txt; cat gpurun_out/r2z_pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.txt 2>&1; tail -1 gpurun_out/r2z_smoke.txt
python bench.py --steps 50 --warmup 5 > gpurun_out/r2z_default_n1.json 2> gpurun_out/r2z_default_n1.err
python bench.py --impl reference > gpurun_out/r2z_default_reference.json 2> gpurun_out/r2z_default_reference.err
python bench.py --steps 30 --warmup 5 --dropout 0.5 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2z_default_n1_dropout05.json 2>/dev/null
python bench.py --steps 30 --warmup 5 --gemm-precision tf32 --no-cpu-baseline --no-eager-baseline > gpurun_out/r2z_default_n1_tf32.json 2>/dev/null
for wl in c1_mf_bpr_u10k_i5k_d64_b256 c2_sasrec_d128_seq50_items1M_k256_b1024 c3_gru_d256_seq100_items5M_bpr5_b2048; do
  python bench.py --workload $wl --steps 30 --warmup 5 > gpurun_out/r2z_${wl}_n1.json 2> gpurun_out/r2z_${wl}_n1.err
done
python bench.py --workload c4_sasrec_d256_L4_seq200_items10M_k4096_b4096 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2z_c4_n1.json 2> gpurun_out/r2z_c4_n1.err
python profiles/gemm_shapes.py > gpurun_out/r2z_gemm_shapes.txt 2>&1
python profiles/small_gemm.py > gpurun_out/r2z_small_gemm.txt 2>&1
python profiles/gemm_roles.py > gpurun_out/r2z_gemm_roles.txt 2>&1
ls -la gpurun_out | tail -20
