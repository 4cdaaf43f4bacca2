#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fullsize_props_gpu.py tests/test_engine_gpu.py -q -m gpu -p no:cacheprovider 2>&1 | tail -n 6
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["phase_ms_last_clip"], d["clocks"])
PY
tail -n 2 gpurun_out/bench_quick.err
