#!/bin/bash
# round-2 GPU session E (1 GPU): GEMM re-check (DMMA order), launch lists of our kernels only, full capture of the stack
# kernel at 500 chains per GPU (the 8-GPU regime of the named configuration).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K='regex:chain_sweep|gf_stack|misfit|sum_like|dgemm|finish|laplacian|residual_from|geodetic'
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-trace-writer"
timeout 900 python -m pytest tests -m gpu -q -x -k "dense or geodetic or mvn or misfit or joint" > gpurun_out/e_pytest_gemm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/e_pytest_gemm.log
timeout 600 python bench.py --config c2llk --noise dense $Q > gpurun_out/e_bench_c2dense.json 2> gpurun_out/e_bench_c2dense.err
timeout 600 python bench.py --config c4 $Q > gpurun_out/e_bench_c4.json 2> gpurun_out/e_bench_c4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dgemm_tile' -s 2 -c 1 -o gpurun_out/e_dgemm_c2dense -f \
    python bench.py --config c2llk --noise dense --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/e_ncu_c2dense_full.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 60 --csv --log-file gpurun_out/e_launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/e_ncu_c3_list.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/e_launches_c4.csv \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/e_ncu_c4_list.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/e_launches_c5.csv \
    python bench.py --config c5 --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/e_ncu_c5_list.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gf_stack_chunk -s 3 -c 1 -o gpurun_out/e_stack_500 -f \
    python bench.py --chains 500 --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/e_ncu500_full.out 2>&1
ls -la gpurun_out | tail -12
