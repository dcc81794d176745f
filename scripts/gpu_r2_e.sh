#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "matched|passed|failed|FAILED|engine" gpurun_out/pytest_gpu.log | head -60
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tc2.json 2> gpurun_out/bench_tc2.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_tc2.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_tc2.json"))
print(round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
print([ (k,v) for k,v in d["roofline"]["launches_ms"] if k.startswith("tc")])
for k,v in (d.get("extra") or {}).items(): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a!="e2e"}, v.get("e2e",{}).get("value"))
PY
