#!/bin/bash
# round-2 GPU session J (1 GPU): asynchronous trace writer (device-packed records, page-locked buffers, writer threads):
# GPU sampler tests, the full default bench line, full-size nearest-neighbour parity.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -x -k "sampler or backend or nearest_neighbor or recorder or errors" > gpurun_out/j_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j_pytest.log
timeout 900 python bench.py > gpurun_out/j_bench_n1.json 2> gpurun_out/j_bench_n1.err
df -h /tmp | tail -1 > gpurun_out/j_tmpfs.txt
tail -3 gpurun_out/j_pytest.log
