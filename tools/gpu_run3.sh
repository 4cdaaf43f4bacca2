#!/bin/bash
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
bash tools/gpu_check.sh nobench
timeout 600 python tools/bench_kernels.py > gpurun_out/bench_kernels.log 2>&1; echo "microbench exit $?" | tee -a gpurun_out/summary.txt
grep -E "linear_|conv3x3|temporal3|attn_spatial" gpurun_out/bench_kernels.log
MOFA_ATTN_V1=1 timeout 200 python - > gpurun_out/attn_v1.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'tools'); sys.path.insert(0,'.')
PY
timeout 900 python -m pytest tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/engine_tests.log 2>&1
echo "engine tests exit $?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/engine_tests.log
timeout 600 python tools/profile_step.py --steps 3 --warmup 1 --profile > gpurun_out/step_profile.log 2>&1; echo "step exit $?" | tee -a gpurun_out/summary.txt; tail -n 6 gpurun_out/step_profile.log
# ncu: launch list of one denoise step (cold-cache, serialised: compare shares)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches.csv python tools/profile_step.py --steps 1 > gpurun_out/ncu_list.log 2>&1; echo "ncu list exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 150 -c 3 -o gpurun_out/prof_gemm -f python tools/profile_step.py --steps 1 > gpurun_out/ncu_gemm.log 2>&1; echo "ncu gemm exit $?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:attn_spatial -s 2 -c 1 -o gpurun_out/prof_attn -f python tools/profile_step.py --steps 1 > gpurun_out/ncu_attn.log 2>&1; echo "ncu attn exit $?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out/*.ncu-rep
