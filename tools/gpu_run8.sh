#!/bin/bash
# keypoint / hybrid GPU parity + per-shape profile of the denoise step
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_keypoint_gpu.py -q -m gpu -p no:cacheprovider -x > gpurun_out/t_keypoint.log 2>&1
echo "keypoint tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 15 gpurun_out/t_keypoint.log
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail.log 2>&1; echo "detail exit $?" | tee -a gpurun_out/summary.txt; head -n 70 gpurun_out/step_detail.log
