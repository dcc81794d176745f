#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/pytest_gpu.log
for cfg in 2 5; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras --config $cfg > gpurun_out/bench_sel$cfg.json 2> gpurun_out/bench_sel$cfg.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_sel$cfg.err
python - <<PY
import json
d=json.load(open("gpurun_out/bench_sel$cfg.json"))
print(round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
PY
done
