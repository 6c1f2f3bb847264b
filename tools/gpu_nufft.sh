#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nufft.py -m gpu -q -x --timeout 300 -rfE > gpurun_out/pytest_nufft.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nufft.log
grep -E "^(FAILED|ERROR)|passed|failed|assert" gpurun_out/pytest_nufft.log | tail -12
timeout 600 python tools/prof_nufft.py 5 > gpurun_out/nufft.log 2>&1; cat gpurun_out/nufft.log
