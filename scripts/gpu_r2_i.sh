#!/bin/bash
mkdir -p gpurun_out
cat > /tmp/t.py <<'PY'
import sys, torch
sys.path[:0] = ["tests", "oracle", "."]
from helpers import load_weights
from affnet_b200.architectures import AffNetFast
W = load_weights()
a = AffNetFast(PS=32); a.load_state_dict(W["affnet"]); a = a.eval().cuda()
P = (torch.rand(450, 1, 32, 32) * 255).cuda()
print(a(P).sum().item())
PY
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/t.py > gpurun_out/sanitizer.log 2>&1; echo "rc=$?"
grep -v "^=========     Host Frame\|^=========         in \|^=========     Saved host" gpurun_out/sanitizer.log | head -60
