#!/bin/bash
# round-end evidence: bench (both arms), ncu launch list of one denoise step, DRAM traffic of the GEMM launches
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json; tail -n 2 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref exit $?"; cat gpurun_out/bench_ref.json
timeout 200 python tools/profile_step.py --steps 3 --warmup 1 2>&1 | tail -n 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 1 --warmup 0 > gpurun_out/ncu_launches.log 2>&1; echo "ncu launch list exit $?"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:gemm_tc --clock-control none -c 500 --csv --log-file gpurun_out/gemm_traffic.csv python tools/profile_step.py --steps 1 --warmup 0 > gpurun_out/ncu_traffic.log 2>&1; echo "ncu traffic exit $?"
