run() { name=$1; shift; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29500 + RANDOM % 400)) bench.py --gpus 8 "$@" > gpurun_out/r2f_$name.json 2> gpurun_out/r2f_$name.err; echo "$name rc=$?"; grep -v "Warning\|warn\|OMP\|\*\*\*\|return func" gpurun_out/r2f_$name.err | tail -4; }
run def_n8 --steps 30 --warmup 5
