#!/bin/bash
# round-2 GPU session W (1 GPU): per-(chain, patch) plan cache on / off: parity tests, then C3 multilinear at 4000 and 500
# chains, nearest neighbour, C5.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_blocking.py tests/test_gpu_errors.py tests/test_gpu_fuzz.py tests/test_gpu_fullsize.py -m gpu -q -x > gpurun_out/w_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/w_pytest.log
B="python bench.py --no-cpu-baseline --no-strict-f64 --no-trace-writer --steps 20 --warmup 5"
for pc in 1 0; do
  BEATGPU_PLAN_CACHE=$pc timeout 300 $B > gpurun_out/w_pc${pc}_4000.json 2> gpurun_out/w_pc${pc}_4000.err
  BEATGPU_PLAN_CACHE=$pc timeout 300 $B --chains 500 > gpurun_out/w_pc${pc}_500.json 2> gpurun_out/w_pc${pc}_500.err
  BEATGPU_PLAN_CACHE=$pc timeout 300 $B --interpolation nearest_neighbor > gpurun_out/w_pc${pc}_nn.json 2> gpurun_out/w_pc${pc}_nn.err
  BEATGPU_PLAN_CACHE=$pc timeout 300 $B --config c5 > gpurun_out/w_pc${pc}_c5.json 2> gpurun_out/w_pc${pc}_c5.err
done
tail -3 gpurun_out/w_pytest.log
