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
timeout 300 python tools/prof_ops.py 1 > gpurun_out/plain2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_fwd_quad|quad_build|radon_adj_tile|filter_tc_kernel" -c 10 -f -o gpurun_out/prof_ops python tools/prof_ops.py 1 > gpurun_out/ncu_full.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ff_|interp_|crop_apod|transpose_kdata" -c 10 -f -o gpurun_out/prof_nufft python tools/prof_nufft_one.py -1 > gpurun_out/ncu_full_nufft.log 2>&1
ls -la gpurun_out | tail -30
