#!/bin/bash
mkdir -p gpurun_out
echo "default"; timeout 120 python tools/prof_attn_case.py
echo "hsum"; MOFA_ATTN_POLY=-1 timeout 120 python tools/prof_attn_case.py
echo "spin"; MOFA_ATTN_IDLE_NS=0 timeout 120 python tools/prof_attn_case.py
echo "hsum+spin"; MOFA_ATTN_POLY=-1 MOFA_ATTN_IDLE_NS=0 timeout 120 python tools/prof_attn_case.py
echo "idle100"; MOFA_ATTN_IDLE_NS=100 timeout 120 python tools/prof_attn_case.py
echo "default again"; timeout 120 python tools/prof_attn_case.py
MOFA_ATTN_POLY=-1 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attn_spatial" 2>&1 | tail -n 2
