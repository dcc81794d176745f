#!/bin/bash
# N-GPU check (default 2): the bench under torchrun with a tight timeout (weak scaling, overlapped all-gather).
N=${1:-2}
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 \
    > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
d=json.loads(open('gpurun_out/bench_${N}gpu.json').read().strip().splitlines()[-1])
print('N=$N value %.1f e2e %.1f ms/step %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step']), d['config']['parallelism'], d['clocks'])
PY
