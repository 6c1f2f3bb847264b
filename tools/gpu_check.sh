#!/bin/bash
# quick gate: every GPU test, then the operator timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest.log 2>&1; tail -2 gpurun_out/pytest.log
timeout 600 python tools/prof_nufft.py 5 > gpurun_out/nufft.log 2>&1; grep -E "default|r01 one|pad \+ cuFFT" gpurun_out/nufft.log
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1; grep -E "radon_fwd|radon_adj" gpurun_out/ops.log
