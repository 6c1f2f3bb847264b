#!/bin/bash
# r02 first gate: every GPU test (isolated workers), smoke(), a short bench, operator timings
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 60 --timeout 600 -rfE -s > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed|benched-size" gpurun_out/pytest.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -12 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
cut -c1-600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1; cat gpurun_out/ops.log
timeout 600 python tools/prof_nufft.py 5 > gpurun_out/nufft.log 2>&1; tail -12 gpurun_out/nufft.log
