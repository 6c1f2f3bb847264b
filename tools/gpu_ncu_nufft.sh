#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_nufft_one.py 2 > gpurun_out/prof_nufft_plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ff_" -c 4 -f -o gpurun_out/prof_nufft_r2 python tools/prof_nufft_one.py 2 > gpurun_out/ncu_nufft_full.log 2>&1
tail -2 gpurun_out/ncu_nufft_full.log
