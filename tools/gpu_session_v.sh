#!/bin/bash
# round-2 GPU session V (8 GPUs): the N = 8 headline line (weak + strong + samplers + strict f64) on the final build.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 600 $T --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/v_bench_n8.json 2> gpurun_out/v_bench_n8.err
tail -n 3 gpurun_out/v_bench_n8.err
