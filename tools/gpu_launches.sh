#!/bin/bash
mkdir -p gpurun_out
export PDU_BENCH_AUTOTUNE=0 PDU_BENCH_GRAPH=0
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/launches.csv
