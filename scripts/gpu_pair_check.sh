#!/bin/bash
# Guarded first run of the cta_group::2 layers: targeted parity test under a short timeout, then the full round.
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -k "hardnet or nets" > gpurun_out/pair_pytest.log 2>&1; rc=$?
echo "pair pytest rc=$rc"; tail -15 gpurun_out/pair_pytest.log
if [ $rc -ne 0 ]; then
  echo "pair path failed; rerunning with AG_NO_PAIR=1"
  AG_NO_PAIR=1 timeout 240 python -m pytest tests/test_gpu_parity.py -x -q -k "hardnet or nets" 2>&1 | tail -5
  exit 1
fi
AG_NO_PAIR=1 timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nopair.json 2> gpurun_out/bench_nopair.err; tail -c 1500 gpurun_out/bench_nopair.json
SKIP_NCU=1 timeout 1200 bash scripts/gpu_round.sh
