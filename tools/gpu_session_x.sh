#!/bin/bash
# round-2 GPU session X (1 GPU): the shipping build once more -- full GPU suite, default bench line, ncu capture of the
# stack + misfit kernels for profiles/stack_kernel_traffic.json.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/x_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/x_pytest.log
timeout 600 python bench.py > gpurun_out/x_bench_n1.json 2> gpurun_out/x_bench_n1.err
B="python bench.py --no-cpu-baseline --no-strict-f64 --no-trace-writer --steps 2 --warmup 3"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'gf_stack_chunk|misfit_warp|plan_cache' -s 9 -c 3 -o gpurun_out/x_stack_misfit_4000 -f $B > gpurun_out/x_ncu_full.out 2>&1
tail -n 3 gpurun_out/x_pytest.log
