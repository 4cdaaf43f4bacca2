#!/bin/bash
# MUFU / conversion throughput microbenchmarks behind profiles/r1_attention_experiments.md (run on the GPU box)
mkdir -p gpurun_out
for f in mufu_rate mufu_cvt_rate mufu_h2_rate; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/$f tools/microbench/$f.cu && timeout 60 /tmp/$f | tee gpurun_out/$f.txt
done
