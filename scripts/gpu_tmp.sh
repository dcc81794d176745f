#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 120 tests/probe/_bin/tc_rates > gpurun_out/tc_rates.txt 2>&1; head -12 gpurun_out/tc_rates.txt | tail -5
for n in 1184 48000; do
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_xprof.so timeout 300 python scripts/role_prof_x.py $n > gpurun_out/role_prof_x_$n.txt 2>&1; echo "roleprof $n rc=$?"
done
grep "==\|issuer" gpurun_out/role_prof_x_48000.txt | cut -c1-150 | head -24
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
