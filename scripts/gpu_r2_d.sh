#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcx.py -q -s > gpurun_out/pytest_tcx.log 2>&1; echo "pytest rc=$?"
grep -E "layer|engine|passed|failed|FAILED|Error|error|assert" gpurun_out/pytest_tcx.log | head -70
