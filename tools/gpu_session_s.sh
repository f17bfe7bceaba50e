#!/bin/bash
# round-2 GPU session S (1 GPU): default L2 budget 0.6: blocking tests, C3BIG / C3 / C3-f64 lines.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests -m gpu -q -x -k "blocking or fullsize" > gpurun_out/s_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/s_pytest.log
B="python bench.py --no-cpu-baseline --no-trace-writer --steps 20 --warmup 5"
timeout 300 $B --config c3big --no-strict-f64 > gpurun_out/s_bench_c3big.json 2> gpurun_out/s_bench_c3big.err
timeout 300 $B > gpurun_out/s_bench_c3.json 2> gpurun_out/s_bench_c3.err
timeout 900 ncu --set full --clock-control none -k regex:'gf_stack_chunk' -s 3 -c 1 -o gpurun_out/s_stack_c3big -f \
    $B --config c3big --no-strict-f64 --steps 2 --warmup 3 > gpurun_out/s_ncu_c3big.out 2>&1
tail -2 gpurun_out/s_pytest.log
