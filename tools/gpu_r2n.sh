#!/bin/bash
# final ncu --set full captures of the kernels changed last: fused NUFFT forward, fan-beam backprojector, generic adjoint row pass
mkdir -p gpurun_out
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/plain_nufft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fz_cols_fwd|fz_rows_fwd|fz_combine" -s 3 -c 3 -f -o gpurun_out/r02_nufft_fused_fwd python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/ncu_r02_nufft_fused_fwd.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ff_rows_adj|ff_cols_adj|interp_adj_csrT|crop_apod|transpose_kdata|interp_adj_csr_long" -s 6 -c 6 -f -o gpurun_out/r02_nufft python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/ncu_r02_nufft.log 2>&1
timeout 300 python tools/prof_fan.py > gpurun_out/plain_fan.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_adj_tile" -s 1 -c 1 -f -o gpurun_out/r02_fan_adj python tools/prof_fan.py > gpurun_out/ncu_r02_fan_adj.log 2>&1
ls -la gpurun_out/r02_nufft_fused_fwd.ncu-rep gpurun_out/r02_nufft.ncu-rep gpurun_out/r02_fan_adj.ncu-rep
