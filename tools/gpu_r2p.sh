#!/bin/bash
# ncu --set full of the generic adjoint at 64 planes after the 4-lanes-per-cell gather (long rows inside the same launch)
mkdir -p gpurun_out
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/plain_nufft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ff_rows_adj|ff_cols_adj|interp_adj_csrT|crop_apod|transpose_kdata" -s 5 -c 5 -f -o gpurun_out/r02_nufft_adj2 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/ncu_r02_nufft_adj2.log 2>&1
ls -la gpurun_out/r02_nufft_adj2.ncu-rep
