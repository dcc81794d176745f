#!/bin/bash
# r02 session 1: tcgen05 rate probe + baseline / variant bench at batch 16, 1, 64
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 tests/probe/tc_rates > gpurun_out/tc_rates.txt 2>&1; echo "rates rc=$?"
cat gpurun_out/tc_rates.txt
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_base.json 2> gpurun_out/bench_base.err; echo "base rc=$?"
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_exp.so timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_exp.json 2> gpurun_out/bench_exp.err; echo "exp rc=$?"
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_exp.so timeout 300 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_exp.log 2>&1; echo "pytest exp rc=$?"; tail -3 gpurun_out/pytest_exp.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --batch 1 > gpurun_out/bench_b1.json 2> gpurun_out/bench_b1.err; echo "b1 rc=$?"
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --batch 64 > gpurun_out/bench_b64.json 2> gpurun_out/bench_b64.err; echo "b64 rc=$?"
python - <<'PY'
import json
for n in ("base","exp","b1","b64"):
    try:
        d=json.load(open("gpurun_out/bench_%s.json"%n))
        print(n, round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
    except Exception as e:
        print(n, "failed", e)
PY
