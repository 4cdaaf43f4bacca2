#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ff_geglu_kernel -s 2 -c 1 -o gpurun_out/r2_ncu_ff_fused -f python tools/prof_gemm_case.py ff_fused 3 > gpurun_out/ncu_ff.log 2>&1; echo "ncu ff exit $?"
tail -3 gpurun_out/ncu_ff.log; ls -la gpurun_out/*.ncu-rep
