#!/bin/bash
# round-2 GPU session Q (1 GPU): patch-chunk size in the small-population regime (500 / 250 / 64 chains per GPU): with fewer
# chains than resident warps several (target, chunk) groups are in flight at once and share L2.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline --no-strict-f64 --no-trace-writer --steps 30 --warmup 5"
for ch in 29 20 15 10 7 5; do
  BEATGPU_CHUNK=$ch timeout 300 $B --chains 500 > gpurun_out/q_chunk${ch}_500.json 2> gpurun_out/q_chunk${ch}_500.err
done
for ch in 29 15 7; do
  BEATGPU_CHUNK=$ch timeout 300 $B --chains 1000 > gpurun_out/q_chunk${ch}_1000.json 2> gpurun_out/q_chunk${ch}_1000.err
  BEATGPU_CHUNK=$ch timeout 300 $B --chains 128 > gpurun_out/q_chunk${ch}_128.json 2> gpurun_out/q_chunk${ch}_128.err
done
ls gpurun_out/q_* | wc -l
