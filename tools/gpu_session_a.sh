#!/bin/bash
# round-2 GPU session A (1 GPU): regression + new parity tests, the new bench line, row-alignment experiment, probes,
# launch list + full ncu capture in the 500-chains-per-GPU regime.  Everything lands in gpurun_out/.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1700 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/a_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/a_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench_n1.json 2> gpurun_out/a_bench_n1.err
for al in 16 128 512; do
  BEATGPU_ROW_ALIGN_BYTES=$al timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/a_bench_align${al}.json 2> gpurun_out/a_bench_align${al}.err
  BEATGPU_ROW_ALIGN_BYTES=$al timeout 300 python bench.py --chains 500 --steps 20 --warmup 5 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/a_bench500_align${al}.json 2> gpurun_out/a_bench500_align${al}.err
done
timeout 300 python bench.py --config c3big --steps 20 --warmup 5 --no-cpu-baseline --no-trace-writer > gpurun_out/a_bench_c3big.json 2> gpurun_out/a_bench_c3big.err
timeout 600 python tools/probe_gather.py gpurun_out/a_probe_gather.json --quick > /dev/null 2> gpurun_out/a_probe_gather.err
# 500 chains per GPU: launch list and one full capture of the stack kernel
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/a_launches_500.csv \
    python bench.py --chains 500 --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/a_ncu500_list.out 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gf_stack_chunk -s 3 -c 1 -o gpurun_out/a_stack_500 -f \
    python bench.py --chains 500 --steps 2 --warmup 3 --no-cpu-baseline --no-strict-f64 --no-trace-writer > gpurun_out/a_ncu500_full.out 2>&1
ls -la gpurun_out | tail -30
