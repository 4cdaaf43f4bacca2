#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_engine_gpu.py tests/test_flow_post.py tests/test_cmp_gpu.py -q -m gpu -p no:cacheprovider -x 2>&1 | tail -n 4
timeout 900 python bench.py --steps 1 --warmup 1 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d["value"], d["e2e"], d["config"]["phase_ms_last_clip"], d["roofline"]["achieved"], d["cpu_baseline"]["seconds_per_sample"])
PY
tail -n 3 gpurun_out/bench_quick.err
