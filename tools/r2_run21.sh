#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests"
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -4
for c in conv320 conv320_stats temporal320 temporal320_stats ff2 qkv320 geglu320 proj320res; do
  for m in 0 1; do
    echo -n "2cta=$m "
    MOFA_GEMM_2CTA=$m timeout 60 python tools/prof_gemm_case.py $c 20 2>&1 | tail -1
  done
done | tee gpurun_out/r2_2cta_ab2.txt
