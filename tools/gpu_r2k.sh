#!/bin/bash
# r02 final gate (no profiler): every GPU test, smoke(), both bench arms, operator / NUFFT timings, the configs[4] sweep
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 60 --timeout 600 -rfE > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -6 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_n1.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'cpu', d.get('cpu_baseline',{}).get('value'))
print('roofline', {k:d['roofline'][k] for k in ('achieved','frac','avg_launch_ms','gsamples_per_s','kernel')})
for k,v in d['operators'].items(): print(k, round(v['ms']*1e3,1),'us', round(v['hbm_frac'],4))
for k,v in d['extras'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
r=json.load(open('gpurun_out/bench_reference.json')); print('reference', r.get('value'), r.get('unit'), r.get('cpu_baseline'))
PY
timeout 600 python tools/prof_ops.py 7 > gpurun_out/ops.log 2>&1; cat gpurun_out/ops.log
timeout 600 python tools/prof_nufft.py 7 all --variants > gpurun_out/nufft_variants.log 2>&1; grep -v "^    " gpurun_out/nufft_variants.log | grep -E "default" 
timeout 900 python tools/sweep.py gpurun_out/r02_sweep.md > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"; tail -3 gpurun_out/sweep.log
