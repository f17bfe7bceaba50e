#!/bin/bash
# round-2 GPU session D (1 GPU): pipelined DMMA GEMM (cp.async ring, 128x64 tiles): parity tests, C2-dense / C4 lines, ncu;
# reference arm with the full 17-node host library.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
free -g | head -2 > gpurun_out/d_host.txt; nproc >> gpurun_out/d_host.txt; cat /sys/fs/cgroup/memory.max >> gpurun_out/d_host.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -x -k "dense or geodetic or mvn or misfit or fuzz or joint" > gpurun_out/d_pytest_gemm.log 2>&1
echo "pytest rc=$?" >> gpurun_out/d_pytest_gemm.log
Q="--steps 20 --warmup 5 --no-cpu-baseline --no-trace-writer"
timeout 600 python bench.py --config c2llk --noise dense $Q > gpurun_out/d_bench_c2dense.json 2> gpurun_out/d_bench_c2dense.err
timeout 600 python bench.py --config c4 $Q > gpurun_out/d_bench_c4.json 2> gpurun_out/d_bench_c4.err
timeout 600 python bench.py --noise dense $Q --no-strict-f64 > gpurun_out/d_bench_c3dense.json 2> gpurun_out/d_bench_c3dense.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dgemm_tile' -s 2 -c 1 -o gpurun_out/d_dgemm_c2dense -f \
    python bench.py --config c2llk --noise dense --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/d_ncu_c2dense_full.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dgemm_tile' -s 8 -c 2 -o gpurun_out/d_dgemm_c4 -f \
    python bench.py --config c4 --steps 2 --warmup 3 --no-cpu-baseline --no-trace-writer > gpurun_out/d_ncu_c4_full.out 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/d_bench_ref.json 2> gpurun_out/d_bench_ref.err
ls -la gpurun_out | tail -12
