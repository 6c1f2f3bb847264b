#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/plain_nufft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fz_cols_fwd|fz_rows_fwd|fz_combine" -s 3 -c 3 -f -o gpurun_out/r02_nufft_fused_fwd python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/ncu_r02_nufft_fused_fwd.log 2>&1
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b2" > gpurun_out/plain_nufft2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fz_rows_adj|fz_cols_adj" -s 2 -c 2 -f -o gpurun_out/r02_nufft_fused_adj python tools/prof_nufft.py 1 "cfg4 320^2 c8 b2" > gpurun_out/ncu_r02_nufft_fused_adj.log 2>&1
timeout 600 python tools/prof_ops.py 7 > gpurun_out/ops.log 2>&1
timeout 600 python tools/prof_nufft.py 7 all --variants > gpurun_out/nufft_variants.log 2>&1
for w in cfg1 cfg4; do
  timeout 300 python tools/prof_mri_step.py $w 3 > gpurun_out/plain_mri_$w.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_mri_$w.csv python tools/prof_mri_step.py $w 3 > gpurun_out/ncu_mri_$w.log 2>&1
done
ls -la gpurun_out/r02_* gpurun_out/launches_mri*
