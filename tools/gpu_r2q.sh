#!/bin/bash
# final r02 gate: driver-style tests / smoke / bench / reference arm, operator timings, sweep, parity table
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/ -x -q -m gpu > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest.log; tail -2 gpurun_out/pytest.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"
timeout 900 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"
python - <<'PY'
import json
b = json.loads(open('gpurun_out/bench_n1.json').read().strip().splitlines()[-1])
print({k: b[k] for k in ('value', 'ms_per_step', 'gpu_launches')}, 'e2e', b['e2e']['value'], 'cpu', b['cpu_baseline']['value'])
for k, v in b.get('operators', {}).items():
    print(k, v)
for k, v in b.get('extras', {}).items():
    print(k, {a: c for a, c in v.items() if a in ('ms_per_step', 'slices_per_s')})
PY
timeout 600 python tools/prof_ops.py 5 > gpurun_out/ops.log 2>&1
timeout 600 python tools/prof_nufft.py 7 all --variants > gpurun_out/nufft_variants.log 2>&1; grep -c . gpurun_out/nufft_variants.log
timeout 900 python tools/sweep.py gpurun_out/r02_sweep.md > gpurun_out/sweep.log 2>&1; echo "sweep rc=$?"
timeout 600 python tools/parity_report.py gpurun_out/r02_parity.md > gpurun_out/parity.log 2>&1; echo "parity rc=$?"; tail -12 gpurun_out/r02_parity.md
timeout 300 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/plain_nufft.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ff_rows_adj|ff_cols_adj|interp_adj_csrT|crop_apod|transpose_kdata" -s 5 -c 5 -f -o gpurun_out/r02_nufft_adj2 python tools/prof_nufft.py 1 "cfg4 320^2 c8 b8" > gpurun_out/ncu_r02_nufft_adj2.log 2>&1
ls -la gpurun_out/r02_nufft_adj2.ncu-rep
