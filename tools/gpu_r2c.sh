#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_nufft.py -m gpu -q -n 1 --max-worker-restart 20 --timeout 600 -rfE > gpurun_out/pytest_nufft.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_nufft.log
grep -E "^(FAILED|ERROR)|passed|failed|Error|error" gpurun_out/pytest_nufft.log | tail -30
timeout 600 python tools/prof_nufft.py 5 all --pg > gpurun_out/nufft.log 2>&1; grep -v "^    " gpurun_out/nufft.log | tail -80
timeout 300 python tools/prof_nufft.py 2 "cfg4 320^2 c8 b8" > gpurun_out/plain_nufft.log 2>&1 && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"fz_|crop_apod" --csv --log-file gpurun_out/launches_nufft2.csv python tools/prof_nufft.py 2 "cfg4 320^2 c8 b8" > gpurun_out/ncu_nufft2.log 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_nufft2.csv')))
hdr=None
agg=collections.OrderedDict()
for r in rows:
    if 'Kernel Name' in r: hdr=r; continue
    if hdr is None or len(r)!=len(hdr): continue
    d=dict(zip(hdr,r))
    k=d['Kernel Name'][:60]; m=d['Metric Name']; v=float(d['Metric Value'].replace(',',''))
    agg.setdefault(k,{}).setdefault(m,[]).append(v)
for k,v in agg.items():
    print(k, {m:(round(sum(x)/len(x),1),len(x)) for m,x in v.items()})
PY
