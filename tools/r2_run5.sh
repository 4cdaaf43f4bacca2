#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_kernels2_gpu.py tests/test_engine_gpu.py tests/test_keypoint_gpu.py tests/test_cmp_gpu.py tests/test_graph_step_gpu.py tests/test_clip_engine.py tests/test_fullsize_parity_gpu.py::test_traj_step_576x1024_elementwise -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests5.log 2>&1; echo "gpu tests: $?"
grep -E "passed|failed|FAILED|Error|error" gpurun_out/gpu_tests5.log | tail -20
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2b.txt 2>&1; echo "profile: $?"
head -64 gpurun_out/step_detail_r2b.txt
