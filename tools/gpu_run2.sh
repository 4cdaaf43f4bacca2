#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/engine_tests.log 2>&1
echo "engine tests exit $?" | tee gpurun_out/summary2.txt
tail -n 15 gpurun_out/engine_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary2.txt; tail -n 3 gpurun_out/smoke.log
timeout 600 python tools/profile_step.py --steps 3 --warmup 1 --profile > gpurun_out/step_profile.log 2>&1; echo "step exit $?" | tee -a gpurun_out/summary2.txt; cat gpurun_out/step_profile.log | tail -n 12
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench1.json 2> gpurun_out/bench1.err; echo "bench exit $?" | tee -a gpurun_out/summary2.txt; cat gpurun_out/bench1.json; tail -n 5 gpurun_out/bench1.err
