#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -q -s -k "fused_octave or gaussian_blur" > gpurun_out/pytest_pyr.log 2>&1; echo "pytest rc=$?"
grep -E "passed|failed|FAILED|Error|assert" gpurun_out/pytest_pyr.log | head -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_pyr.json 2> gpurun_out/bench_pyr.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_pyr.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_pyr.json"))
print(round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), {k:round(v,3) for k,v in d["roofline"]["stages_ms"].items()})
print([ (k,v) for k,v in d["roofline"]["launches_ms"] if not k.startswith("tc")])
PY
