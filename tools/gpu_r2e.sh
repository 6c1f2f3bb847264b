#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -n 1 --max-worker-restart 30 --timeout 600 -rfE > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log
grep -E "^(FAILED|ERROR)|passed|failed" gpurun_out/pytest.log | tail -20
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -5 gpurun_out/bench.err
python - <<'PY'
import json
try:
    d=json.load(open('gpurun_out/bench.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e'])
    print('roofline', {k:d['roofline'][k] for k in ('kernel','achieved','frac','avg_launch_ms')})
    for k,v in d['operators'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='kernel'})
    for k,v in d['extras'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
    print(d.get('cpu_baseline'))
except Exception as e: print('bench parse failed', e)
PY
