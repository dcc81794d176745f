#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tcx.py -q -s > gpurun_out/pytest_tcx.log 2>&1; echo "pytest rc=$?"
grep -E "engine|passed|failed|FAILED|raw heads|Error" gpurun_out/pytest_tcx.log | head -20
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200_xprof.so timeout 300 python scripts/role_prof_x.py 48000 2>&1 | tee gpurun_out/role_prof_x.txt | grep -A5 "slot 0\|==" | cut -c1-200 | head -40
for v in ""; do
AFFNET_B200_LIB=$PWD/affnet_b200/lib/libaffnet_b200$v.so timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_tc2$v.json 2> gpurun_out/bench_tc2$v.err; echo "bench$v rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_tc2$v.json"))
print("$v", round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
print([ (k,v) for k,v in d["roofline"]["launches_ms"] if k.startswith("tc")])
PY
done
