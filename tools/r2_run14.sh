#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "geglu or statistics" > gpurun_out/gpu_tests14.log 2>&1; echo "tests: $?"; tail -3 gpurun_out/gpu_tests14.log
for c in temporal320 temporal320_stats conv320_stats; do python tools/prof_gemm_case.py $c 10; done
for d in 1 2 3; do MOFA_GN_DEBUG=$d python tools/prof_gemm_case.py temporal320_stats 10 | sed "s/^/debug=$d /"; done
python tools/prof_gemm_case.py geglu320 10
