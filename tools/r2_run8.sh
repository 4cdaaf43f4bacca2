#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_keypoint_gpu.py tests/test_cmp_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests13.log 2>&1; echo "gpu tests: $?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/gpu_tests13.log | tail -12
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2i.txt 2>&1; echo "profile: $?"
head -44 gpurun_out/step_detail_r2i.txt | cut -c1-140
