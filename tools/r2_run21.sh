#!/bin/bash
mkdir -p gpurun_out
echo "== kernel tests (pairs incl. odd tile counts)"
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
echo "== kernel tests, MOFA_GEMM_2CTA=2"
MOFA_GEMM_2CTA=2 timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu 2>&1 | tail -3
for c in geglu1280 lin5120; do
  for m in 0 1; do
    echo -n "2cta=$m "
    MOFA_GEMM_2CTA=$m timeout 60 python tools/prof_gemm_case.py $c 20 2>&1 | tail -1
  done
done | tee gpurun_out/r2_2cta_ab3.txt
for c in geglu320 proj320res proj320rb; do
  for m in 1 2; do
    echo -n "2cta=$m "
    MOFA_GEMM_2CTA=$m timeout 60 python tools/prof_gemm_case.py $c 20 2>&1 | tail -1
  done
done | tee -a gpurun_out/r2_2cta_ab3.txt
