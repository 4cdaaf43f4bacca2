#!/bin/bash
mkdir -p gpurun_out
echo "split=1 handoff=1 poly=0"; MOFA_ATTN_SPLIT=1 timeout 120 python tools/prof_attn_case.py
for poly in 0 4 2; do
  echo "split=2 poly=$poly"; MOFA_ATTN_SPLIT=2 MOFA_ATTN_POLY=$poly timeout 120 python tools/prof_attn_case.py
done
MOFA_ATTN_SPLIT=2 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attn_spatial" 2>&1 | tail -n 3
MOFA_ATTN_SPLIT=2 MOFA_ATTN_POLY=4 timeout 300 python -m pytest tests/test_kernels_gpu.py -q -m gpu -p no:cacheprovider -k "attn_spatial" 2>&1 | tail -n 3
MOFA_ATTN_SPLIT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial2 -s 1 -c 1 -o gpurun_out/prof_attn_v4 -f python tools/prof_attn_case.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
