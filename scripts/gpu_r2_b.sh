#!/bin/bash
mkdir -p gpurun_out
timeout 120 tests/probe/tc_rates > gpurun_out/tc_rates.txt 2>&1; echo "rates rc=$?"
head -8 gpurun_out/tc_rates.txt
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "matched|passed|failed|FAILED|engine" gpurun_out/pytest_gpu.log | head -60
