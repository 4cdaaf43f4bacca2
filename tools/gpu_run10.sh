#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial2 -s 1 -c 1 -o gpurun_out/prof_attn_v5 -f python tools/prof_attn_case.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
MOFA_ATTN_SPLIT=2 timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial2 -s 1 -c 1 -o gpurun_out/prof_attn_v5s2 -f python tools/prof_attn_case.py > gpurun_out/ncu_attn2.log 2>&1; echo "ncu attn exit $?"
