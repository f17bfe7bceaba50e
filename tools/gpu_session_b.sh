#!/bin/bash
# round-2 GPU session B (1 GPU): full GPU test suite, chunk-kernel occupancy variants (new address arithmetic), L2 budget sweep,
# e2e upload split on/off, launch list at 500 chains, probes (fixed), ncu of the batched-TMA probe.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strict-f64 --no-trace-writer"
for occ in 5 6 7; do
  BEATGPU_CHUNK_OCC=$occ timeout 300 $B > gpurun_out/b_occ${occ}_4000.json 2> gpurun_out/b_occ${occ}_4000.err
  BEATGPU_CHUNK_OCC=$occ timeout 300 $B --chains 500 > gpurun_out/b_occ${occ}_500.json 2> gpurun_out/b_occ${occ}_500.err
done
for fr in 0.3 0.5 0.6; do
  BEATGPU_L2_FRAC=$fr timeout 300 $B --store f64 > gpurun_out/b_f64_l2frac${fr}.json 2> gpurun_out/b_f64_l2frac${fr}.err
  BEATGPU_L2_FRAC=$fr timeout 300 $B --config c3big > gpurun_out/b_c3big_l2frac${fr}.json 2> gpurun_out/b_c3big_l2frac${fr}.err
done
timeout 300 $B --store f64 > gpurun_out/b_f64_l2frac0.4.json 2> gpurun_out/b_f64_l2frac0.4.err
BEATGPU_SPLIT_H2D=0 timeout 300 $B > gpurun_out/b_split0.json 2> gpurun_out/b_split0.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'chain_sweep|gf_stack|misfit|sum_like' -c 40 --csv --log-file gpurun_out/b_launches_500.csv \
    python bench.py --chains 500 --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/b_ncu500_list.out 2>&1
timeout 600 python tools/probe_gather.py gpurun_out/b_probe_gather.json --quick > /dev/null 2> gpurun_out/b_probe_gather.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:probe_tma_batch -s 1 -c 1 -o gpurun_out/b_probe_mode3 -f \
    python tools/probe_gather.py --quick --mode=3 > gpurun_out/b_ncu_probe3.out 2>&1
timeout 2400 python -m pytest tests -m gpu -q --durations=20 > gpurun_out/b_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/b_pytest.log
ls -la gpurun_out | tail -40
