#!/bin/bash
mkdir -p gpurun_out
for sp in 1 2; do for pv in 14 13 -1; do
  echo -n "SPLIT=$sp POLY=$pv "
  MOFA_ATTN_SPLIT=$sp MOFA_ATTN_POLY=$pv timeout 100 python tools/prof_attn_case.py 2>&1 | tail -1
done; done | tee gpurun_out/r2_attn_variants5.txt
