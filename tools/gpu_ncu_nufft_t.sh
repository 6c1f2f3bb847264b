#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_nufft_one.py 2 > gpurun_out/prof_nufft_plain.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_nufft.csv python tools/prof_nufft_one.py 2 > gpurun_out/ncu_nufft.log 2>&1
