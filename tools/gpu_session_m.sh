#!/bin/bash
# round-2 GPU session M (= G on the final build) (8 GPUs, one box): the configurations BASELINE.json names at their GPU counts --
# C3 headline at N=8 (weak line + `strong` block: n_chains = 4000 partitioned over 8 ranks), C5 with the PT driver
# (512 chains over 8 ranks), C4 (2000 chains over 4 ranks) -- plus N=2 / N=4 of the headline for the scaling table.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 900 $T --nproc-per-node 8 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/m_bench_n8.json 2> gpurun_out/m_bench_n8.err
timeout 600 $T --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --config c5 --steps 20 --warmup 5 --no-trace-writer > gpurun_out/m_bench_c5_n8.json 2> gpurun_out/m_bench_c5_n8.err
timeout 600 $T --nproc-per-node 4 --master-port 29513 bench.py --gpus 4 --config c4 --steps 20 --warmup 5 --no-trace-writer > gpurun_out/m_bench_c4_n4.json 2> gpurun_out/m_bench_c4_n4.err
timeout 600 $T --nproc-per-node 4 --master-port 29514 bench.py --gpus 4 --steps 20 --warmup 5 --no-trace-writer --no-strict-f64 > gpurun_out/m_bench_n4.json 2> gpurun_out/m_bench_n4.err
timeout 600 $T --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 20 --warmup 5 --no-trace-writer --no-strict-f64 > gpurun_out/m_bench_n2.json 2> gpurun_out/m_bench_n2.err
timeout 300 $T --nproc-per-node 8 --master-port 29516 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/m_bench_ref_n8.json 2> gpurun_out/m_bench_ref_n8.err
tail -2 gpurun_out/g_*.err
