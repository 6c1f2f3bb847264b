#!/bin/bash
# full gate + launch lists (CT step via bench, MRI steps) for profiles/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 30 --timeout 600 -rfE > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -8 gpurun_out/smoke.log
timeout 600 python tools/prof_nufft.py 5 all > gpurun_out/nufft.log 2>&1; grep -v "^    " gpurun_out/nufft.log | tail -16
for w in cfg1 cfg4; do
  timeout 300 python tools/prof_mri_step.py $w 3 > gpurun_out/plain_mri_$w.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_mri_$w.csv python tools/prof_mri_step.py $w 3 > gpurun_out/ncu_mri_$w.log 2>&1
done
export PDU_BENCH_AUTOTUNE=0 PDU_BENCH_GRAPH=0
timeout 600 python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/*.csv
