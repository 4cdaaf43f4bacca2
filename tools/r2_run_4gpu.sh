#!/bin/bash
# 4 GPUs: N = 4 bench line through the peer-store gather (smoke test of the N > 2 path; --steps 1 is not a bench value)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n4.json 2> gpurun_out/bench_n4.err; echo "bench n4: $?"
tail -c 900 gpurun_out/bench_n4.json; grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n4.err | tail -8
