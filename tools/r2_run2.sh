#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/gpu_tests2.log 2>&1; echo "gpu tests: $?"
grep -E "^\[fullsize\]|passed|failed|FAILED|Error|error" gpurun_out/gpu_tests2.log | tail -30
timeout 900 python bench.py --steps 3 --warmup 2 > gpurun_out/bench_quick2.json 2> gpurun_out/bench_quick2.err; echo "bench: $?"
tail -c 2500 gpurun_out/bench_quick2.json; tail -5 gpurun_out/bench_quick2.err
