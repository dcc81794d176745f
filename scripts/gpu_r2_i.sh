#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu.log
for i in 1 2; do
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_i$i.json 2> gpurun_out/bench_i$i.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/bench_i$i.json"))
print(round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
print([ (k,round(v,4)) for k,v in d["roofline"]["launches_ms"] if k.startswith("blur") or k.startswith("tcx_first")][:12])
PY
done
