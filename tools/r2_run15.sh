#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "geglu or statistics" > gpurun_out/gpu_tests15.log 2>&1; echo "tests: $?"; tail -3 gpurun_out/gpu_tests15.log
for c in temporal320 temporal320_stats conv320_stats; do python tools/prof_gemm_case.py $c 10; done
timeout 600 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2j.txt 2>&1; echo "profile: $?"
head -30 gpurun_out/step_detail_r2j.txt | cut -c1-140
