#!/bin/bash
# round-2 GPU session H (1 GPU): rupture sweep with several chains per warp -- bit-exactness tests, then the step at 4000
# and 500 chains per GPU with the packing on and off, launch list.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K='regex:chain_sweep|gf_stack|misfit|sum_like'
timeout 1500 python -m pytest tests -m gpu -q -x -k "sweep or Sweeper or fuzz or parity or ops_protocol or sampler" > gpurun_out/h_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/h_pytest.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strict-f64 --no-trace-writer"
for pack in 1 0; do
  BEATGPU_SWEEP_PACK=$pack timeout 300 $B > gpurun_out/h_pack${pack}_4000.json 2> gpurun_out/h_pack${pack}_4000.err
  BEATGPU_SWEEP_PACK=$pack timeout 300 $B --chains 500 > gpurun_out/h_pack${pack}_500.json 2> gpurun_out/h_pack${pack}_500.err
  BEATGPU_SWEEP_PACK=$pack timeout 300 $B --config c5 > gpurun_out/h_pack${pack}_c5.json 2> gpurun_out/h_pack${pack}_c5.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/h_launches_4000.csv \
    $B --steps 2 --warmup 3 > gpurun_out/h_ncu_list.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/h_launches_500.csv \
    $B --chains 500 --steps 2 --warmup 3 > gpurun_out/h_ncu500_list.out 2>&1
tail -3 gpurun_out/h_pytest.log
