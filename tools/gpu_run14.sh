#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_kernels2_gpu.py -q -m gpu -p no:cacheprovider -x -k "gemm or conv" 2>&1 | tail -n 3
for c in proj320res proj320rb ff2 qkv320 geglu320; do timeout 120 python tools/prof_gemm_case.py $c 5; done 2>&1 | tee gpurun_out/gemm_cases.log
timeout 300 python tools/profile_step.py --steps 3 --warmup 1 --profile > gpurun_out/step_profile.log 2>&1; cat gpurun_out/step_profile.log
