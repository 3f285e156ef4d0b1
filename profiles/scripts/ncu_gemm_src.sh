# source-level ncu capture of the 3xTF32 instantiation of the tcgen05 GEMM (K-major operands) on one per-layer shape
ncu --set full --import-source on --clock-control none --kernel-name-base mangled -k regex:gemm_tc_kernelILb0ELb0ELb1E --launch-skip 3 --launch-count 1 -o /tmp/gemm_src python profiles/gemm_shapes.py "${1:-fwd qkv}" > /dev/null 2>&1
ncu -i /tmp/gemm_src.ncu-rep --page source --csv > gpurun_out/gemm_src_qkv.csv 2>/dev/null
ncu -i /tmp/gemm_src.ncu-rep --page raw --csv > gpurun_out/gemm_raw_qkv.csv 2>/dev/null
ls -la gpurun_out/gemm_src_qkv.csv gpurun_out/gemm_raw_qkv.csv
