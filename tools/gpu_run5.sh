#!/bin/bash
# validation of the rewritten GEMM epilogue + native VAE decoder, then step profile and bench
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash tools/gpu_check.sh nobench
timeout 300 python -m pytest tests/test_kernels2_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/t_8.log 2>&1
echo "kernels2 exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/t_8.log
for c in qkv320 geglu320 proj320res ff2; do timeout 120 python tools/prof_gemm_case.py $c 5; done 2>&1 | tee gpurun_out/gemm_cases.log
timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "microbench exit $?" | tee -a gpurun_out/summary.txt
grep -E "linear_|conv3x3|temporal3" gpurun_out/bench_kernels.log
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/engine_tests.log 2>&1
echo "engine tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/engine_tests.log
timeout 600 python tools/profile_step.py --steps 3 --warmup 1 --profile > gpurun_out/step_profile.log 2>&1; echo "step exit $?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/step_profile.log
timeout 900 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench2.json; tail -n 5 gpurun_out/bench2.err
