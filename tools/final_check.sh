#!/bin/bash
# last check of the round: every GPU test, smoke, one timed clip
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -p no:cacheprovider -x > gpurun_out/t_all.log 2>&1; echo "pytest -m gpu exit $?"; tail -n 3 gpurun_out/t_all.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench exit $?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print("value", d["value"], "e2e", d["e2e"]["value"], d["config"]["phase_ms_last_clip"], d["clocks"])
PY
tail -n 2 gpurun_out/bench_quick.err
