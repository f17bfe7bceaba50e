#!/bin/bash
# round-2 GPU session N (1 GPU): nearest-neighbour stack kernel -- where is its time (ncu), occupancy variants; the
# example script with the trace writer on.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --interpolation nearest_neighbor --no-cpu-baseline --no-strict-f64 --no-trace-writer"
for occ in 5 6 7; do
  BEATGPU_CHUNK_OCC=$occ timeout 300 $B --steps 20 --warmup 5 > gpurun_out/n_nn_occ${occ}.json 2> gpurun_out/n_nn_occ${occ}.err
  BEATGPU_CHUNK_OCC=$occ timeout 300 $B --steps 20 --warmup 5 --chains 500 > gpurun_out/n_nn_occ${occ}_500.json 2> gpurun_out/n_nn_occ${occ}_500.err
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gf_stack_chunk' -s 3 -c 1 -o gpurun_out/n_stack_nn -f \
    $B --steps 2 --warmup 3 > gpurun_out/n_ncu_nn_full.out 2>&1
rm -rf /tmp/beat_trace_example
timeout 600 python examples/smc_c3_synthetic.py --small --chains 256 --steps 12 --trace-dir /tmp/beat_trace_example > gpurun_out/n_example.log 2>&1
ls /tmp/beat_trace_example | head -5 >> gpurun_out/n_example.log; ls /tmp/beat_trace_example/stage_1 | wc -l >> gpurun_out/n_example.log
tail -5 gpurun_out/n_example.log
