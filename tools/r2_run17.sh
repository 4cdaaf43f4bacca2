#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "ff_geglu or short_k" > gpurun_out/gpu_tests17.log 2>&1; echo "ff tests: $?"; grep -E "passed|failed|FAILED|Error|mismatch" gpurun_out/gpu_tests17.log | tail -8
timeout 120 python tools/prof_gemm_case.py ff_fused 10; timeout 120 python tools/prof_gemm_case.py ff_unfused 10
timeout 600 python -m pytest tests/test_engine_gpu.py tests/test_keypoint_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests17b.log 2>&1; echo "engine tests: $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/gpu_tests17b.log | tail -8
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2l.txt 2>&1; echo "profile: $?"
head -24 gpurun_out/step_detail_r2l.txt | cut -c1-140
