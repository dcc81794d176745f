"""Diagnostic (GPU box): where do the hand-crafted orientation bins of the CUDA kernel differ from the oracle's?  For the reference-extracted
19x19 patches of tests/golden/handcrafted.npz: per mismatching patch the oracle's top-2 margin, and how many pixels fall into another bin when
the same formula runs with CUDA's atan2f (torch on the GPU) instead of the CPU's."""
import math
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")]
import affnet_oracle as O  # noqa: E402
from helpers import gold  # noqa: E402
from affnet_b200.HandCraftedModules import OrientationDetector  # noqa: E402


def hist(P, dev):
    x = P.to(dev)
    xp = F.pad(x, (1, 1, 0, 0), "replicate"); gx = 0.5 * xp[:, :, :, :-2] - 0.5 * xp[:, :, :, 2:]
    yp = F.pad(x, (0, 0, 1, 1), "replicate"); gy = 0.5 * yp[:, :, :-2, :] - 0.5 * yp[:, :, 2:, :]
    gk = 10.0 * torch.from_numpy(O.circular_gauss_kernel(19).astype(np.float32)).to(dev)
    mag = torch.sqrt(gx * gx + gy * gy + 1e-10) * gk
    ori = torch.atan2(gy, gx)
    o_big = 36.0 * (ori + 1.0 * math.pi) / (2.0 * math.pi)
    bo0 = torch.floor(o_big); wo1 = o_big - bo0; bo0 = bo0 % 36
    wo0 = (1.0 - wo1) * mag
    bins = torch.stack([((bo0 == i).float() * wo0).mean(dim=(1, 2, 3)) for i in range(36)], dim=1)
    sm = F.conv1d(bins.view(-1, 1, 36), torch.tensor([[[0.33, 0.34, 0.33]]], device=dev), padding=1).view(-1, 36)
    return ori.cpu(), o_big.cpu(), bo0.cpu(), sm.cpu(), gx.cpu(), gy.cpu()


z = gold("handcrafted.npz")
P = torch.from_numpy(z["patches19"])
ref = torch.from_numpy(z["angle"])
ours = OrientationDetector(patch_size=19)(P.cuda()).cpu()
d = torch.atan2(torch.sin(ours - ref), torch.cos(ours - ref)).abs()
bad = (d > 1e-5).nonzero().view(-1)
oc, obc, bc, smc, gx, gy = hist(P, "cpu")
og, obg, bg, smg, _, _ = hist(P, "cuda")
print("patches %d, kernel != reference: %d ; torch-on-GPU argmax != reference: %d" % (P.size(0), bad.numel(), int((smg.argmax(1) != smc.argmax(1)).sum())))
diffpix = (bc != bg)
print("pixels whose bin differs CPU vs CUDA torch: %d of %d ; of those with gx==0 or gy==0: %d" % (int(diffpix.sum()), diffpix.numel(), int((diffpix & ((gx == 0) | (gy == 0))).sum())))
if diffpix.any():
    i = diffpix.nonzero()[:8]
    for r in i:
        t = tuple(r.tolist())
        print("  gx %.6g gy %.6g  ori cpu %.9g gpu %.9g  o_big cpu %.9g gpu %.9g" % (gx[t], gy[t], oc[t], og[t], obc[t], obg[t]))
top = smc.topk(2, dim=1).values
for b in bad.tolist():
    print("patch %d: ref bin %d ours-angle %.4f ref-angle %.4f, oracle top-2 margin %.3e, pixels in other bin (CUDA atan2): %d" % (
        b, int(smc[b].argmax()), ours[b], ref[b], float((top[b, 0] - top[b, 1]) / top[b, 0]), int(diffpix[b].sum())))
