#!/bin/bash
mkdir -p gpurun_out
V="${1:-1 3}"
timeout 300 python tools/prof_adj.py $V > gpurun_out/prof_adj_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_adj_tile" -o gpurun_out/prof_adj_r2 -f python tools/prof_adj.py $V > gpurun_out/ncu_adj.log 2>&1
tail -3 gpurun_out/ncu_adj.log
