#!/bin/bash
# r02 ncu --set full captures: projector (cell kernel), backprojector, fused NUFFT forward kernels, generic adjoint kernels
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py -1 > gpurun_out/plain_fwd.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_fwd_quad|quad_build" -c 2 -f -o gpurun_out/r02_fwd python tools/prof_fwd.py -1 > gpurun_out/ncu_r02_fwd.log 2>&1
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/plain_nufft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fz_|ff_rows_adj|ff_cols_adj|interp_adj_csrT|crop_apod|transpose_kdata|interp_adj_csr_long" -s 9 -c 9 -f -o gpurun_out/r02_nufft python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/ncu_r02_nufft.log 2>&1
# DRAM traffic of the projector inside a step (caches not flushed between kernels)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"radon_fwd_quad|quad_build" --csv --log-file gpurun_out/r02_traffic_warm.csv python tools/prof_fwd.py -1 > gpurun_out/ncu_traffic.log 2>&1
ls -la gpurun_out/r02_*
