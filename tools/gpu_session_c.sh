#!/bin/bash
# round-2 GPU session C (1 GPU): full GPU test suite after the session-B fixes, the default bench line, the C4 / C5 /
# C2 lines, launch lists and full ncu captures of the secondary kernels (sweep, misfit, DMMA GEMM).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -q --durations=10 > gpurun_out/c_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/c_pytest.log
timeout 900 python bench.py > gpurun_out/c_bench_n1.json 2> gpurun_out/c_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c_bench_ref.json 2> gpurun_out/c_bench_ref.err
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-trace-writer"
timeout 600 python bench.py --config c4 $Q > gpurun_out/c_bench_c4.json 2> gpurun_out/c_bench_c4.err
timeout 600 python bench.py --config c5 $Q > gpurun_out/c_bench_c5.json 2> gpurun_out/c_bench_c5.err
timeout 600 python bench.py --config c2llk --noise dense $Q > gpurun_out/c_bench_c2dense.json 2> gpurun_out/c_bench_c2dense.err
# launch lists (serialised, cold cache: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c_launches_c3.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/c_ncu_c3_list.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/c_launches_c4.csv \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/c_ncu_c4_list.out 2>&1
# full captures: one launch of each secondary kernel
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'chain_sweep|misfit_kernel|sum_like' -s 9 -c 3 -o gpurun_out/c_secondary_c3 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/c_ncu_c3_full.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dgemm_tile' -s 8 -c 2 -o gpurun_out/c_dgemm_c4 -f \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/c_ncu_c4_full.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dgemm_tile' -s 2 -c 1 -o gpurun_out/c_dgemm_c2dense -f \
    python bench.py --config c2llk --noise dense --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/c_ncu_c2dense_full.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gf_stack_chunk -s 3 -c 1 -o gpurun_out/c_stack_4000 -f \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/c_ncu_stack_full.out 2>&1
# cuBLAS DGEMM yardstick (not a product path): what the FP64 tensor pipe delivers on this box
timeout 300 python tools/dgemm_yardstick.py > gpurun_out/c_dgemm_yardstick.json 2> gpurun_out/c_dgemm_yardstick.err
ls -la gpurun_out | tail -30
