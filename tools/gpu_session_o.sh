#!/bin/bash
# round-2 GPU session O (1 GPU): nearest-neighbour stack kernel variants: 8 patches in flight at the 5-CTA register budget,
# 4 in flight at 7 and 10 CTAs per SM.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --interpolation nearest_neighbor --no-cpu-baseline --no-strict-f64 --no-trace-writer --steps 20 --warmup 5"
for occ in 5 7 10; do
  BEATGPU_CHUNK_OCC=$occ timeout 300 $B > gpurun_out/o_nn_occ${occ}.json 2> gpurun_out/o_nn_occ${occ}.err
  BEATGPU_CHUNK_OCC=$occ timeout 300 $B --chains 500 > gpurun_out/o_nn_occ${occ}_500.json 2> gpurun_out/o_nn_occ${occ}_500.err
done
timeout 600 python -m pytest tests -m gpu -q -x -k "nn_exp or nearest or stack_rand or golden" > gpurun_out/o_pytest.log 2>&1
tail -2 gpurun_out/o_pytest.log
