#!/bin/bash
# exactly what the driver runs at round end, plus the profile refresh
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest_driver.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_driver.log; tail -3 gpurun_out/pytest_driver.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err; tail -1 gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json
timeout 900 python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1
timeout 600 python tools/prof_nufft.py 5 > gpurun_out/nufft.log 2>&1
timeout 900 python tools/sweep.py gpurun_out/r01_sweep.md > gpurun_out/sweep.log 2>&1
timeout 600 python tools/parity_report.py gpurun_out/r01_parity.md > gpurun_out/parity.log 2>&1
bash tools/gpu_launches.sh > /dev/null 2>&1
# one ncu --set full capture per operator (each script has already exited 0 above or is re-run plain first)
NCU="ncu --set full --clock-control none --import-source on -f"
timeout 300 python tools/prof_fwd.py -1 > gpurun_out/plain_fwd.log 2>&1 && \
timeout 900 $NCU -k regex:"radon_fwd_quad|quad_build" -c 4 -o gpurun_out/prof_ops python tools/prof_fwd.py -1 > gpurun_out/ncu_full.log 2>&1
timeout 300 python tools/prof_adj.py -1 > gpurun_out/plain_adj.log 2>&1 && \
timeout 900 $NCU -k regex:"radon_adj_tile" -c 2 -o gpurun_out/prof_adj python tools/prof_adj.py -1 > gpurun_out/ncu_full_adj.log 2>&1
timeout 300 python tools/prof_fan.py > gpurun_out/plain_fan.log 2>&1 && \
timeout 900 $NCU -k regex:"radon_fwd_quad|radon_adj_tile" -c 4 -o gpurun_out/prof_fan python tools/prof_fan.py > gpurun_out/ncu_full_fan.log 2>&1
timeout 300 python tools/prof_ops.py 1 > gpurun_out/plain2.log 2>&1 && \
timeout 900 $NCU -k regex:"filter_tc_kernel" -c 2 -o gpurun_out/prof_filter python tools/prof_ops.py 1 > gpurun_out/ncu_full_filter.log 2>&1
timeout 300 python tools/prof_nufft_one.py -1 > gpurun_out/plain_nufft.log 2>&1 && \
timeout 900 $NCU -k regex:"ff_|interp_|crop_apod|transpose_kdata" -c 10 -o gpurun_out/prof_nufft python tools/prof_nufft_one.py -1 > gpurun_out/ncu_full_nufft.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"radon_fwd_quad|quad_build|radon_adj_tile" --csv --log-file gpurun_out/traffic_warm.csv python tools/prof_fwd.py -1 > gpurun_out/ncu_traffic.log 2>&1
ls -la gpurun_out | tail -30
