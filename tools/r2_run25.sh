#!/bin/bash
mkdir -p gpurun_out
for v in 2 0 1 2 0; do MOFA_GN_APPLY=$v timeout 100 python tools/bench_gn.py 2>&1 | tail -4; done | tee gpurun_out/r2_gn_apply_prefetch.txt
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "groupnorm or gn" 2>&1 | tail -2
