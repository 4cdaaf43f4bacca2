#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_cvt tools/microbench/mufu_cvt_rate.cu && timeout 60 /tmp/mufu_cvt | tee gpurun_out/mufu_cvt_rate.txt
