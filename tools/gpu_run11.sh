#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attn_spatial" 2>&1 | tail -n 5
echo "split=1 handoff=1"; timeout 120 python tools/prof_attn_case.py
echo "split=1 handoff=0"; MOFA_ATTN_HANDOFF=0 timeout 120 python tools/prof_attn_case.py
echo "split=2"; MOFA_ATTN_SPLIT=2 timeout 120 python tools/prof_attn_case.py
MOFA_ATTN_SPLIT=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attn_spatial" 2>&1 | tail -n 3
