#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nufft.py tests/test_gpu_radon.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_b.log 2>&1; tail -2 gpurun_out/pytest_b.log
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1; grep -E "fan512 (fwd|adj)" gpurun_out/ops.log
timeout 900 python tools/train_bench.py > gpurun_out/train_n1.md 2> gpurun_out/train_n1.err; cat gpurun_out/train_n1.md | tail -4
