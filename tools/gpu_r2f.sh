#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_radon.py tests/test_gpu_golden.py -m gpu -q -n 1 --max-worker-restart 30 --timeout 600 -rfE > gpurun_out/pytest_radon.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_radon.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_radon.log | tail -20
timeout 600 python tools/prof_ops.py 7 > gpurun_out/ops.log 2>&1; grep -E "radon_fwd|fan512 fwd" gpurun_out/ops.log
