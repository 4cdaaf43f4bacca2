#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/prof_attn_case.py
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attn_spatial" 2>&1 | tail -n 3
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial2 -s 1 -c 1 -o gpurun_out/prof_attn_v3 -f python tools/prof_attn_case.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
