#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_nufft.py tests/test_gpu_model.py -m gpu -q -n 1 --max-worker-restart 30 --timeout 600 -rfE > gpurun_out/pytest_mri.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_mri.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest_mri.log | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_nocpu.json 2> gpurun_out/bench_nocpu.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_nocpu.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'])
for k,v in d['operators'].items(): print(k, round(v['ms']*1e3,1),'us', round(v['hbm_frac'],4))
for k,v in d['extras'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
PY
for w in cfg1; do
  timeout 300 python tools/prof_mri_step.py $w 3 > gpurun_out/plain_mri_$w.log 2>&1 && \
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_mri_$w.csv python tools/prof_mri_step.py $w 3 > gpurun_out/ncu_mri_$w.log 2>&1
done
