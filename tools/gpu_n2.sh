#!/bin/bash
# driver-style 2-GPU launch of both arms
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"
tail -c 6000 gpurun_out/bench_n2.json; tail -5 gpurun_out/bench_n2.err
