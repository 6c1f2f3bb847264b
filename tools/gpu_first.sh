#!/bin/bash
# first GPU contact: smoke, parity tests (isolated workers), bench with default and baseline variants
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1; nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
PDU_RADON_FWD_VARIANT=0 PDU_RADON_ADJ_VARIANT=0 timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v0.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_v0.log
timeout 1500 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 60 --timeout 600 -rfE > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
PDU_RADON_FWD_VARIANT=0 PDU_RADON_ADJ_VARIANT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_v0.json 2> gpurun_out/bench_v0.err; echo "bench rc=$?" >> gpurun_out/bench_v0.err
tail -3 gpurun_out/smoke.log gpurun_out/smoke_v0.log; tail -15 gpurun_out/pytest.log; cat gpurun_out/bench.json | cut -c1-1500
