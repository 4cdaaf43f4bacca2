#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "ff_geglu or geglu" 2>&1 | tail -3
for d in 0 1 2 4 7 31; do
  echo "== MOFA_FF_DEBUG=$d"
  MOFA_FF_DEBUG=$d timeout 120 python tools/prof_gemm_case.py ff_fused 20 2>&1 | tail -1
done | tee gpurun_out/r2_ff_debug2.txt
timeout 120 python tools/prof_gemm_case.py ff_unfused 20 2>&1 | tail -1
