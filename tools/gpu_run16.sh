#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/host_profile.py > gpurun_out/host_profile.log 2>&1; head -n 60 gpurun_out/host_profile.log
