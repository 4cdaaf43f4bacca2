#!/bin/bash
mkdir -p gpurun_out
for pv in 14 15 16 14; do
  echo -n "HANDOFF=0 POLY=$pv "
  MOFA_ATTN_POLY=$pv timeout 100 python tools/prof_attn_case.py 2>&1 | tail -1
done | tee gpurun_out/r2_attn_variants4.txt
