#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_h2 tools/microbench/mufu_h2_rate.cu && timeout 60 /tmp/mufu_h2 | tee gpurun_out/mufu_h2_rate.txt
