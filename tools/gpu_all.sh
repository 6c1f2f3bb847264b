#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 60 --timeout 300 -rfE > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|^E  " gpurun_out/pytest.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cut -c1-260 gpurun_out/bench.json; tail -2 gpurun_out/bench.err
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1; grep -E "variant 1|variant 5|fan|nufft" gpurun_out/ops.log
timeout 600 python tools/prof_nufft.py 5 > gpurun_out/nufft.log 2>&1; cat gpurun_out/nufft.log
timeout 300 python tools/prof_nufft.py 1 > gpurun_out/plain3.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_nufft.csv python tools/prof_nufft.py 1 > gpurun_out/ncu_nufft.log 2>&1
