#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_cur.json 2> gpurun_out/bench_cur.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_cur.json"))
print("cur", round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
PY
for v in sw2off; do
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_$v.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_tcx.py -q -s -k "end_to_end or benchmark_configs or full_size or tcx_nets" > gpurun_out/pytest_$v.log 2>&1; echo "$v pytest rc=$?"
grep -E "vs oracle|vs reference|engine tc2|passed|failed" gpurun_out/pytest_$v.log | sed 's/max|dA|.*max|ddesc|/ddesc/' | head -24
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_$v.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_$v.json 2> gpurun_out/bench_$v.err; echo "bench $v rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_$v.json"))
print("$v", round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step", [ (k,v) for k,v in d["roofline"]["launches_ms"] if k.startswith("tcx")][-5:])
PY
done
