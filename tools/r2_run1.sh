#!/bin/bash
# round-2 first GPU pass: reference softsplat kernel outputs, the whole -m gpu suite (incl. full-size parity), quick bench
mkdir -p gpurun_out
python -m oracle.make_softsplat_ref --run > gpurun_out/softsplat_run.log 2>&1; echo "softsplat ref run: $?"
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu tests: $?"
grep -E "^\[fullsize\]|passed|failed|FAILED|Error" gpurun_out/gpu_tests.log | tail -40
timeout 900 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench: $?"
tail -c 3000 gpurun_out/bench_quick.json
