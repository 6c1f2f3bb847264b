#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_radon.py -m gpu -q -n 1 --max-worker-restart 60 --timeout 300 -rfE > gpurun_out/pytest_radon.log 2>&1; grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest_radon.log | tail -12
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1; cat gpurun_out/ops.log
export PDU_BENCH_AUTOTUNE=0 PDU_BENCH_GRAPH=0
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/prof_ops.py 1 > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_fwd_strip|filter_tc_kernel" -c 8 -o gpurun_out/prof_fwd2 python tools/prof_ops.py 1 > gpurun_out/ncu_full2.log 2>&1
ls -la gpurun_out | tail -8
