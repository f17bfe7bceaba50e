#!/bin/bash
# round-2 GPU session P (1 GPU): full-size parity of configs 4 and 5.
set -x
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_fullsize_c45.py -m gpu -q --durations=3 > gpurun_out/p_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/p_pytest.log
tail -15 gpurun_out/p_pytest.log
