#!/bin/bash
# driver-style N-GPU launch: bash tools/gpu_n8.sh <N>
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'])
for k,v in d['extras'].items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items() if a!='workload'})
PY
tail -3 gpurun_out/bench_n$N.err
