#!/bin/bash
# full validation on a B200 box: all GPU tests, smoke, bench, step profiles (gpurun -- bash tools/validate_gpu.sh)
mkdir -p gpurun_out
rm -f gpurun_out/summary.txt
timeout 1500 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1
echo "pytest -m gpu exit $?" | tee -a gpurun_out/summary.txt; tail -n 4 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit $?" | tee -a gpurun_out/summary.txt; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?" | tee -a gpurun_out/summary.txt; cat gpurun_out/bench.json; tail -n 3 gpurun_out/bench.err
timeout 300 python tools/profile_step.py --steps 3 --warmup 1 --profile > gpurun_out/step_profile.log 2>&1; cat gpurun_out/step_profile.log
timeout 300 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail.log 2>&1; head -n 30 gpurun_out/step_detail.log
