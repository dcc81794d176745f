"""Developer experiment (CPU, no GPU): how many bits do the residual ("lo") planes of AffNet / OriNet activations need?
fp64 forward of the two nets on the golden patches with the output of chosen layers quantised as fp16 only | fp16 + fp16 residual (what the
engine stores) | fp16 + fp8 residual (e5m2, or e4m3 scaled by 2^10).  Output: max / mean deviation of the head outputs from the unquantised
forward.  AffNet needs <= 5e-5 (OriNet amplifies its error 15x towards the 1e-3 LAF contract).    python scripts/emu_residual_bits.py"""
import sys, torch, torch.nn.functional as F
import os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import affnet_oracle as O
from helpers import gold, load_weights
W = load_weights()
torch.set_num_threads(8)
z = gold("graf_crop.npz")
g = torch.Generator().manual_seed(8)
P = torch.cat([torch.from_numpy(z["aff_patches"])[:300], torch.rand(50, 1, 32, 32, generator=g) * 255]).double()
for name, cfgname in (("affnet", "AFF"), ("orinet", "ORI")):
    sd = {k: v.double() for k, v in W[name].items()}
    cfg = [(1,16,1),(16,16,1),(16,32,2),(32,32,1),(32,64,2),(64,64,1)]
    def run(quant_after, q):
        x = O.input_norm(P.float()).double()
        for li, (i, (cin, cout, stride)) in enumerate(zip(O.CONV_IDX, cfg)):
            x = F.conv2d(x, sd["features.%d.weight" % i], stride=stride, padding=1)
            m = sd["features.%d.running_mean" % (i + 1)].view(1, -1, 1, 1); v = sd["features.%d.running_var" % (i + 1)].view(1, -1, 1, 1)
            x = F.relu((x - m) / torch.sqrt(v + O.BN_EPS))
            if li + 1 in quant_after: x = q(x)
        # head conv (8x8 -> 1x1) + tanh
        hi = [k for k in sd if k.endswith(".weight") and sd[k].dim() == 4][-1]
        y = F.conv2d(x, sd[hi], bias=sd.get(hi.replace("weight", "bias")))
        return torch.tanh(y).flatten(1)
    ident = lambda x: x
    def q_fp16(x): return x.half().double()
    def q_hi_lo16(x):
        h = x.half().double(); return h + (x - h).half().double()
    def q_hi_lo8(x):
        h = x.half().double(); return h + (x - h).float().to(torch.float8_e5m2).double()
    def q_hi_lo8m3(x):
        h = x.half().double(); r = (x - h).float() * 1024.0; return h + r.to(torch.float8_e4m3fn).double() / 1024.0
    ref = run((), ident)
    for label, layers, q in (("L2 out fp16 only", (2,), q_fp16), ("L2 out hi+lo16", (2,), q_hi_lo16), ("L2 out hi+lo e5m2", (2,), q_hi_lo8),
                             ("L2 out hi+lo e4m3*1024", (2,), q_hi_lo8m3),
                             ("all layers hi+lo16", (1,2,3,4,5), q_hi_lo16), ("L2,L4 hi+lo e5m2, rest lo16", None, None)):
        if layers is None:
            def run2():
                x = O.input_norm(P.float()).double()
                for li, (i, (cin, cout, stride)) in enumerate(zip(O.CONV_IDX, cfg)):
                    x = F.conv2d(x, sd["features.%d.weight" % i], stride=stride, padding=1)
                    m = sd["features.%d.running_mean" % (i + 1)].view(1, -1, 1, 1); v = sd["features.%d.running_var" % (i + 1)].view(1, -1, 1, 1)
                    x = F.relu((x - m) / torch.sqrt(v + O.BN_EPS))
                    if li + 1 in (2, 4): x = q_hi_lo8(x)
                    elif li + 1 in (1, 3, 5): x = q_hi_lo16(x)
                hi = [k for k in sd if k.endswith(".weight") and sd[k].dim() == 4][-1]
                return torch.tanh(F.conv2d(x, sd[hi], bias=sd.get(hi.replace("weight", "bias")))).flatten(1)
            out = run2()
        else:
            out = run(layers, q)
        print("%-8s %-28s max |d out| %.3e   mean %.3e" % (name, label, (out - ref).abs().max().item(), (out - ref).abs().mean().item()))
