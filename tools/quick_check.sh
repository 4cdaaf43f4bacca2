#!/bin/bash
# short GPU check after a kernel change: GEMM + norm + attention kernel tests, engine tests, two GEMM cases, step time
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 3
for c in proj320rb proj320res; do timeout 120 python tools/prof_gemm_case.py $c 5; done
timeout 300 python tools/profile_step.py --steps 3 --warmup 1 --profile 2>&1 | tail -n 6
