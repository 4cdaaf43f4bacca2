#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/gpu_tests11.log 2>&1; echo "gpu tests: $?"
grep -E "^\[fullsize\]|passed|failed|FAILED|Error|error" gpurun_out/gpu_tests11.log | tail -20
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2g.txt 2>&1; echo "profile: $?"
head -40 gpurun_out/step_detail_r2g.txt | cut -c1-140
