#!/bin/bash
# round-2 evidence under ncu (numbers printed by a run under ncu are never bench values): launch list of one denoise step,
# DRAM traffic of the GEMM launches, one --set full capture each of the GEGLU K=320 GEMM and the spatial attention
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r2_launches.csv python tools/profile_step.py --steps 1 --warmup 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc --clock-control none -c 600 --csv --log-file gpurun_out/r2_gemm_traffic.csv python tools/profile_step.py --steps 1 --warmup 0 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_ncu_geglu320 -f python tools/prof_gemm_case.py geglu320 > gpurun_out/ncu_geglu.log 2>&1; echo "ncu geglu exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 2 -c 1 -o gpurun_out/r2_ncu_conv320 -f python tools/prof_gemm_case.py conv320 > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_spatial2 -s 2 -c 1 -o gpurun_out/r2_ncu_attn -f python tools/prof_attn_case.py > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?"
ls -la gpurun_out/*.ncu-rep
