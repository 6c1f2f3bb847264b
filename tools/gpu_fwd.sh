#!/bin/bash
# forward-projector A/B: parity tests of every variant, then timings
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_radon.py -m gpu -q -x --timeout 300 -rfE > gpurun_out/pytest_radon.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_radon.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_radon.log | tail -12
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1; cat gpurun_out/ops.log
