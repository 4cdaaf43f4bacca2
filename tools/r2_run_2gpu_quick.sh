#!/bin/bash
# 2 GPUs, quick: N = 2 bench through the peer gather incl. the orderly shutdown (not a bench value: --steps 1)
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n2q.json 2> gpurun_out/bench_n2q.err; echo "bench n2: $?"
tail -c 400 gpurun_out/bench_n2q.json; grep -v "^W0\|OMP_NUM\|^\*\*\*\|^$" gpurun_out/bench_n2q.err | tail -6
