#!/bin/bash
mkdir -p gpurun_out
for d in 0 32 96 4096 4128 8192 8224; do
  echo "== MOFA_FF_DEBUG=$d"
  MOFA_FF_DEBUG=$d timeout 60 python tools/prof_gemm_case.py ff_fused 20 2>&1 | tail -1
done | tee gpurun_out/r2_ff_wake.txt
