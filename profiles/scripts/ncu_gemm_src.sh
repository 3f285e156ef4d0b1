ncu --set full --import-source on --clock-control none -k regex:gemm_tc --launch-skip 8 --launch-count 1 -o /tmp/gemm_src python profiles/gemm_shapes.py "fwd qkv" > /dev/null 2>&1
ncu -i /tmp/gemm_src.ncu-rep --page source --csv > gpurun_out/gemm_src_qkv.csv 2>/dev/null
ncu -i /tmp/gemm_src.ncu-rep --page raw --csv > gpurun_out/gemm_raw_qkv.csv 2>/dev/null
ls -la gpurun_out/gemm_src_qkv.csv gpurun_out/gemm_raw_qkv.csv
