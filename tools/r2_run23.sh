#!/bin/bash
mkdir -p gpurun_out
for m in 0 1; do
  MOFA_GEMM_2CTA=$m timeout 60 python tools/prof_gemm_case.py conv320 20 2>&1 | tail -1
  MOFA_GEMM_2CTA=$m timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_ncu_conv320_2cta$m -f python tools/prof_gemm_case.py conv320 3 > gpurun_out/ncu_conv$m.log 2>&1
done
ls -la gpurun_out/*.ncu-rep | tail -3
