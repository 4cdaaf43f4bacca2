#!/bin/bash
# full validation: every GPU test, smoke, default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -rP -p no:cacheprovider > gpurun_out/gpu_tests_full.log 2>&1; echo "gpu tests: $?"
grep -E "^\[fullsize\]|passed|failed|FAILED|Error|error" gpurun_out/gpu_tests_full.log | tail -16
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 1200 python bench.py --steps 4 --warmup 3 > gpurun_out/bench_r2_v1.json 2> gpurun_out/bench_r2_v1.err; echo "bench: $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_r2_v1.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'], d['roofline']['achieved'], d['roofline']['frac'], d['clocks'])
    print(d['engine']['phase_ms_last_clip']); print(d['other_configs']); print(d['cpu_baseline'])
except Exception as e: print('parse failed', e)
PY
tail -3 gpurun_out/bench_r2_v1.err
