#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "pipeline_batched or full_size" > gpurun_out/pytest_pipe.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_pipe.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "bench2 rc=$?"; tail -3 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_2gpu.json"))
print(round(d["value"],1), "Mpix/s", round(d["ms_per_step"],3), "ms/step e2e", round(d["e2e"]["value"],1), d.get("exchange"))
for k,v in (d.get("extra") or {}).items(): print(k, {a:(round(b,2) if isinstance(b,float) else b) for a,b in v.items() if a not in("e2e","workload","config")}, v.get("e2e",{}).get("value"))
PY
