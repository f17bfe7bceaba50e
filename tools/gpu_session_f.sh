#!/bin/bash
# round-2 GPU session F (1 GPU): full GPU suite (new STFs, laplacian GEMM, GEMM shapes), C4 line + launch list.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K='regex:chain_sweep|gf_stack|misfit|sum_like|dgemm|finish|laplacian|residual_from|geodetic'
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/f_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/f_pytest.log
timeout 600 python bench.py --config c4 --steps 20 --warmup 5 --no-cpu-baseline --no-trace-writer > gpurun_out/f_bench_c4.json 2> gpurun_out/f_bench_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/f_launches_c4.csv \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/f_ncu_c4_list.out 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
tail -5 gpurun_out/f_pytest.log
