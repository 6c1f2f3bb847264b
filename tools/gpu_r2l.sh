#!/bin/bash
# r02 final profiler captures (each only after the same command ran plain with exit 0)
mkdir -p gpurun_out
timeout 300 python tools/prof_fwd.py -1 > gpurun_out/plain_fwd.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_fwd_quad|quad_build|quad_boxes" -c 3 -f -o gpurun_out/r02_fwd python tools/prof_fwd.py -1 > gpurun_out/ncu_r02_fwd.log 2>&1
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"radon_fwd_quad|quad_build|quad_boxes" --csv --log-file gpurun_out/r02_traffic_warm.csv python tools/prof_fwd.py -1 > gpurun_out/ncu_traffic.log 2>&1
timeout 300 python tools/prof_fan.py > gpurun_out/plain_fan.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"radon_adj_tile|radon_fwd_quad" -s 3 -c 2 -f -o gpurun_out/r02_fan python tools/prof_fan.py > gpurun_out/ncu_r02_fan.log 2>&1
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/plain_nufft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fz_|ff_rows_adj|ff_cols_adj|interp_adj_csrT|crop_apod|transpose_kdata|interp_adj_csr_long" -s 9 -c 9 -f -o gpurun_out/r02_nufft python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/ncu_r02_nufft.log 2>&1
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b2" > gpurun_out/plain_nufft2.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"fz_rows_adj|fz_cols_adj" -s 2 -c 2 -f -o gpurun_out/r02_nufft_fused_adj python tools/prof_nufft.py 1 "cfg4 320^2 c8 b2" > gpurun_out/ncu_r02_nufft_fused_adj.log 2>&1
for w in cfg1 cfg4; do
  timeout 300 python tools/prof_mri_step.py $w 3 > gpurun_out/plain_mri_$w.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_mri_$w.csv python tools/prof_mri_step.py $w 3 > gpurun_out/ncu_mri_$w.log 2>&1
done
export PDU_BENCH_AUTOTUNE=0 PDU_BENCH_GRAPH=0
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launches.log 2>&1
unset PDU_BENCH_AUTOTUNE PDU_BENCH_GRAPH
timeout 900 python tools/parity_report.py gpurun_out/r02_parity.md > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -3 gpurun_out/parity.log
ls -la gpurun_out/r02_* gpurun_out/launches*.csv
