#!/bin/bash
# round-2 GPU session R (1 GPU): ncu of the stack kernel on the C3BIG shape (4x larger per-patch library block: L2-derived
# chunk of 7 patches) and the bench line of that shape on the final build.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --config c3big --no-cpu-baseline --no-strict-f64 --no-trace-writer"
timeout 300 $B --steps 20 --warmup 5 > gpurun_out/r_bench_c3big.json 2> gpurun_out/r_bench_c3big.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gf_stack_chunk' -s 3 -c 1 -o gpurun_out/r_stack_c3big -f \
    $B --steps 2 --warmup 3 > gpurun_out/r_ncu_c3big_full.out 2>&1
BEATGPU_CHUNK=29 timeout 900 ncu --set full --clock-control none -k regex:'gf_stack_chunk' -s 3 -c 1 -o gpurun_out/r_stack_c3big_chunk29 -f \
    $B --steps 2 --warmup 3 > gpurun_out/r_ncu_c3big_chunk29.out 2>&1
BEATGPU_CHUNK=29 timeout 300 $B --steps 20 --warmup 5 > gpurun_out/r_bench_c3big_chunk29.json 2> gpurun_out/r_bench_c3big_chunk29.err
ls gpurun_out/r_*
