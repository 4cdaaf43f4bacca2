#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_kernels2_gpu.py tests/test_engine_gpu.py tests/test_keypoint_gpu.py tests/test_cmp_gpu.py tests/test_clip_engine.py -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests16.log 2>&1; echo "tests: $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/gpu_tests16.log | tail -12
for c in qkv320 proj320res proj320rb; do python tools/prof_gemm_case.py $c 10; done
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2k.txt 2>&1; echo "profile: $?"
head -40 gpurun_out/step_detail_r2k.txt | cut -c1-140
