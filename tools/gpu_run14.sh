#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_kernels2_gpu.py tests/test_cmp_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 4
for c in geglu320 proj320res qkv320; do timeout 120 python tools/prof_gemm_case.py $c 5; done 2>&1 | tee gpurun_out/gemm_cases.log
timeout 300 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail.log 2>&1; grep -v "gemm_" gpurun_out/step_detail.log | head -24
