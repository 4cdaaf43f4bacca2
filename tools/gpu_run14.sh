#!/bin/bash
mkdir -p gpurun_out
for c in geglu320 proj320res proj320rb qkv320 ff2; do timeout 120 python tools/prof_gemm_case.py $c 5; done 2>&1 | tee gpurun_out/gemm_cases.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_geglu320_v3 -f python tools/prof_gemm_case.py geglu320 1 > gpurun_out/ncu_geglu.log 2>&1; echo "ncu geglu exit $?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/prof_proj320rb_v3 -f python tools/prof_gemm_case.py proj320rb 1 > gpurun_out/ncu_proj.log 2>&1; echo "ncu proj exit $?"
