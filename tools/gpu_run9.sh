#!/bin/bash
# attention v3 / temporal attention (mma.sync) / GroupNorm + LayerNorm rewrites: parity, then per-shape step profile
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attn or norm" > gpurun_out/t_k1.log 2>&1
echo "attn/norm kernel tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/t_k1.log
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_keypoint_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_engine.log 2>&1
echo "engine+keypoint tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/t_engine.log
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail.log 2>&1; echo "detail exit $?" | tee -a gpurun_out/summary.txt; head -n 45 gpurun_out/step_detail.log
timeout 300 python tools/profile_step.py --steps 3 --warmup 1 --profile > gpurun_out/step_profile.log 2>&1; cat gpurun_out/step_profile.log
