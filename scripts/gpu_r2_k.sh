#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
grep -E "matched|passed|failed|FAILED|engine|raw heads|Error" gpurun_out/pytest_gpu.log | head -60
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --config 5 > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "bench c5 rc=$?"; tail -2 gpurun_out/bench_c5.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_c5.json"))
print(round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), d["dtype"][:60], {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
PY
