#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 30 --timeout 600 -rfE > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -20
timeout 600 python tools/prof_nufft.py 5 all --pg > gpurun_out/nufft.log 2>&1; grep -v "^    " gpurun_out/nufft.log | grep -v "fused, 2 planes" | tail -40
