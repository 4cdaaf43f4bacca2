#!/bin/bash
# Runs each GPU kernel test group in its own process (a hang or sticky CUDA error in one group
# must not take the others down), then the kernel micro-benchmarks.  Logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
i=0
for grp in "test_gemm_linear_plain" "test_gemm_linear_epilogue or test_gemm_linear_split_k or test_gemm_geglu" \
           "test_gemm_conv3x3" "test_gemm_temporal3" "test_attn_spatial" "test_attn_temporal" \
           "groupnorm or layernorm or axpy or linear_small or softsplat or cfg_euler"; do
  i=$((i+1))
  timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "$grp" -p no:cacheprovider > gpurun_out/t_$i.log 2>&1
  echo "group $i ($grp): exit $?" | tee -a gpurun_out/summary.txt
  tail -n 3 gpurun_out/t_$i.log
done
if [ "$1" != "nobench" ]; then
  timeout 900 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1
  echo "bench exit $?" | tee -a gpurun_out/summary.txt
  cat gpurun_out/bench_kernels.log | tail -n 60
fi
