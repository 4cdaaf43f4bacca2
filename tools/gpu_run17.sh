#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/run_variants.py --steps 2 2>&1 | tee gpurun_out/variants.log | tail -n 12
