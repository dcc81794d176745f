#!/bin/bash
# Round-end session: the GPU suite, smoke, then the ncu evidence + bench line of scripts/gpu_profile.sh.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
bash scripts/gpu_profile.sh > gpurun_out/gpu_profile.log 2>&1; echo "profile rc=$?"
tail -c 600 gpurun_out/bench_final.err
