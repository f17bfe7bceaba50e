#!/bin/bash
# round-2 GPU session T (1 GPU): chunks of up to 64 patches (a lane plans two): 29 / 34 / 40 / 50 / 64 patches per chunk at
# 4000 and 500 chains, f32 and f64; blocking + parity tests.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --no-cpu-baseline --no-strict-f64 --no-trace-writer --steps 20 --warmup 5"
for ch in 29 34 40 50 64; do
  BEATGPU_CHUNK=$ch timeout 300 $B > gpurun_out/t_chunk${ch}_4000.json 2> gpurun_out/t_chunk${ch}_4000.err
  BEATGPU_CHUNK=$ch timeout 300 $B --chains 500 > gpurun_out/t_chunk${ch}_500.json 2> gpurun_out/t_chunk${ch}_500.err
done
for ch in 29 34 40 50; do
  BEATGPU_CHUNK=$ch timeout 300 $B --store f64 > gpurun_out/t_chunk${ch}_f64.json 2> gpurun_out/t_chunk${ch}_f64.err
done
timeout 300 $B > gpurun_out/t_default_4000.json 2> gpurun_out/t_default_4000.err
timeout 900 python -m pytest tests -m gpu -q -x -k "blocking or execution_modes or fused_loglike or fuzz" > gpurun_out/t_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/t_pytest.log
tail -2 gpurun_out/t_pytest.log
