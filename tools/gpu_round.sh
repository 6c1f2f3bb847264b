#!/bin/bash
# tests (isolated workers) -> bench -> op timings -> ncu launch list -> ncu full captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 60 --timeout 600 -rfE > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1; cat gpurun_out/ops.log
export PDU_BENCH_AUTOTUNE=0 PDU_BENCH_GRAPH=0
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_launches.log 2>&1
timeout 300 python tools/prof_ops.py 1 > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_fwd_quad|quad_build|radon_adj_tile" -c 12 -f -o gpurun_out/prof_ops python tools/prof_fwd.py -1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_adj_tile" -c 4 -f -o gpurun_out/prof_adj python tools/prof_adj.py -1 > gpurun_out/ncu_full_adj.log 2>&1
# DRAM traffic as it is inside a step (caches not flushed between kernels): the cell tensors come from L2
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"radon_fwd_quad|quad_build|radon_adj_tile" --csv --log-file gpurun_out/traffic_warm.csv python tools/prof_fwd.py -1 > gpurun_out/ncu_traffic.log 2>&1
ls -la gpurun_out | tail -20
cat gpurun_out/bench.json | cut -c1-300; tail -3 gpurun_out/bench.err
