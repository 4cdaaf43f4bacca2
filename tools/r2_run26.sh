#!/bin/bash
mkdir -p gpurun_out
timeout 100 python tools/profile_step.py --steps 2 --warmup 1 --detail > gpurun_out/step_detail_r2p.txt 2>&1; echo "profile: $?"
head -4 gpurun_out/step_detail_r2p.txt | cut -c1-150
