#!/bin/bash
mkdir -p gpurun_out
for c in qkv320 geglu320 proj320res ff2; do timeout 120 python tools/prof_gemm_case.py $c 5; done 2>&1 | tee gpurun_out/gemm_cases.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_qkv320 -f python tools/prof_gemm_case.py qkv320 1 > gpurun_out/ncu_qkv.log 2>&1; echo "ncu qkv exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_geglu320 -f python tools/prof_gemm_case.py geglu320 1 > gpurun_out/ncu_geglu.log 2>&1; echo "ncu geglu exit $?"
