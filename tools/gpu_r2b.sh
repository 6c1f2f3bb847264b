#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nufft.py tests/test_gpu_model.py -m gpu -q -n 1 --max-worker-restart 20 --timeout 600 -rfE > gpurun_out/pytest_nufft.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nufft.log
grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/pytest_nufft.log | tail -30
timeout 600 python tools/prof_nufft.py 5 all --variants > gpurun_out/nufft.log 2>&1; cat gpurun_out/nufft.log | tail -80
