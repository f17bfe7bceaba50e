#!/bin/bash
# round-2 GPU session U (= L on the final build) (1 GPU): the state that ships -- full GPU suite, default bench line, C4 / C5 / C2 lines, launch
# lists, full ncu captures of the stack + misfit kernels (-> profiles/stack_kerneu_traffic.json), compute-sanitizer on the
# kernels that changed in round 2 (DMMA GEMM ring, packed sweep, warp misfit pass, STFs).
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K='regex:chain_sweep|gf_stack|misfit|sum_like|dgemm|finish|laplacian|residuau_from|geodetic'
timeout 2400 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/u_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/u_pytest.log
timeout 900 python bench.py > gpurun_out/u_bench_n1.json 2> gpurun_out/u_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/u_bench_ref.json 2> gpurun_out/u_bench_ref.err
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-trace-writer"
timeout 600 python bench.py --config c4 $Q > gpurun_out/u_bench_c4.json 2> gpurun_out/u_bench_c4.err
timeout 600 python bench.py --config c5 $Q > gpurun_out/u_bench_c5.json 2> gpurun_out/u_bench_c5.err
timeout 600 python bench.py --interpolation nearest_neighbor $Q --no-strict-f64 > gpurun_out/u_bench_nn.json 2> gpurun_out/u_bench_nn.err
B="python bench.py --no-cpu-baseline --no-strict-f64 --no-trace-writer --steps 2 --warmup 3"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/u_launches_4000.csv $B > gpurun_out/u_ncu_list.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/u_launches_500.csv $B --chains 500 > gpurun_out/u_ncu500_list.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 120 --csv --log-file gpurun_out/u_launches_c4.csv $B --config c4 > gpurun_out/u_ncu_c4_list.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gf_stack_chunk|misfit_warp' -s 6 -c 2 -o gpurun_out/u_stack_misfit_4000 -f $B > gpurun_out/u_ncu_full.out 2>&1
timeout 900 compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -x -k "execution_modes or blocking or sweep_packed" > gpurun_out/u_sanitizer_memcheck.log 2>&1
tail -3 gpurun_out/u_pytest.log gpurun_out/u_sanitizer_memcheck.log
