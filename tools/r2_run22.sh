#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_keypoint_gpu.py tests/test_graph_step_gpu.py -m gpu -q -p no:cacheprovider > gpurun_out/gpu_tests22b.log 2>&1; echo "engine tests: $?"; grep -E "passed|failed|FAILED|Error" gpurun_out/gpu_tests22b.log | tail -8
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2n.txt 2>&1; echo "profile: $?"
head -64 gpurun_out/step_detail_r2n.txt | cut -c1-150
