#!/bin/bash
# round-2 GPU session I (1 GPU): warp-per-item misfit pass, conditional sweep packing: full GPU suite, bench lines,
# launch lists, ncu of the misfit pass.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
K='regex:chain_sweep|gf_stack|misfit|sum_like'
timeout 2400 python -m pytest tests -m gpu -q --durations=5 > gpurun_out/i_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/i_pytest.log
B="python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strict-f64 --no-trace-writer"
for mw in 1 0; do
  BEATGPU_MISFIT_WARP=$mw timeout 300 $B > gpurun_out/i_mw${mw}_4000.json 2> gpurun_out/i_mw${mw}_4000.err
  BEATGPU_MISFIT_WARP=$mw timeout 300 $B --chains 500 > gpurun_out/i_mw${mw}_500.json 2> gpurun_out/i_mw${mw}_500.err
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/i_launches_4000.csv \
    $B --steps 2 --warmup 3 > gpurun_out/i_ncu_list.out 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 40 --csv --log-file gpurun_out/i_launches_500.csv \
    $B --chains 500 --steps 2 --warmup 3 > gpurun_out/i_ncu500_list.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'misfit_warp' -s 3 -c 1 -o gpurun_out/i_misfit_warp -f \
    $B --steps 2 --warmup 3 > gpurun_out/i_ncu_misfit_full.out 2>&1
tail -3 gpurun_out/i_pytest.log
