#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_clip_engine.py tests/test_peer_gather_gpu.py tests/test_graph_step_gpu.py tests/test_engine_gpu.py tests/test_sparse_flow.py tests/test_softsplat_ref_gpu.py -m gpu -q -rP -p no:cacheprovider > gpurun_out/gpu_tests3.log 2>&1; echo "gpu tests: $?"
grep -E "^\[fullsize\]|passed|failed|FAILED|Error|error" gpurun_out/gpu_tests3.log | tail -30
timeout 1200 python bench.py --steps 2 --warmup 1 > gpurun_out/bench_quick3.json 2> gpurun_out/bench_quick3.err; echo "bench: $?"
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_quick3.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','e2e','other_configs','gpu_launches')})
    print(d['engine']['phase_ms_last_clip'], d['roofline']['achieved'])
except Exception as e:
    print('parse failed', e)
PY
tail -5 gpurun_out/bench_quick3.err
