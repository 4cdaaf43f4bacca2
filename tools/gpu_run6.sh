#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_qkv320_v2 -f python tools/prof_gemm_case.py qkv320 1 > gpurun_out/ncu_qkv.log 2>&1; echo "ncu qkv exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_spatial2 -s 1 -c 1 -o gpurun_out/prof_attn_v2 -f python tools/prof_attn_case.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
