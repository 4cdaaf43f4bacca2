#!/bin/bash
# 2 GPUs: the NVLink peer-store frame gather against an NCCL gather, then the N = 2 bench line through it
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/peer_gather_check.py > gpurun_out/peer_gather_check.txt 2>&1; echo "peer check: $?"
tail -15 gpurun_out/peer_gather_check.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench n2: $?"
tail -c 1500 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
