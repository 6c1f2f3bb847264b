#!/bin/bash
mkdir -p gpurun_out
V="${1:-7 8}"
timeout 300 python tools/prof_fwd.py $V > gpurun_out/prof_fwd_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_fwd_quad|radon_fwd_strip|quad_build" -o gpurun_out/prof_fwd_r2 -f python tools/prof_fwd.py $V > gpurun_out/ncu_fwd.log 2>&1
tail -3 gpurun_out/ncu_fwd.log
