#!/bin/bash
# quick validation of GEMM epilogue changes + CMP kernels/network
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
i=0
for grp in "test_gemm"; do
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -p no:cacheprovider > gpurun_out/t_gemm.log 2>&1
  echo "gemm tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/t_gemm.log
done
timeout 600 python -m pytest tests/test_cmp_gpu.py tests/test_kernels2_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_cmp.log 2>&1
echo "cmp tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 12 gpurun_out/t_cmp.log
for c in qkv320 geglu320 proj320res ff2; do timeout 120 python tools/prof_gemm_case.py $c 5; done 2>&1 | tee gpurun_out/gemm_cases.log
timeout 600 python tools/profile_step.py --steps 3 --warmup 1 --profile > gpurun_out/step_profile.log 2>&1; echo "step exit $?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/step_profile.log
