#!/bin/bash
# usage: gpu_ncu_full.sh <kernel regex> <out name> <count> <cmd...>
mkdir -p gpurun_out
RE=$1; OUT=$2; CNT=$3; shift 3
timeout 300 "$@" > gpurun_out/plain_$OUT.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -c $CNT -f -o gpurun_out/$OUT "$@" > gpurun_out/ncu_$OUT.log 2>&1
tail -3 gpurun_out/ncu_$OUT.log; ls -la gpurun_out/$OUT.ncu-rep
