"""Developer check (CPU): why can ONE keypoint of the hand-crafted-orientation test take a far-away bin when AffNet changes by 1e-5?
The reference accumulates only the LOWER bin weight (1 - frac) * magnitude per pixel (HandCraftedModules.py:168-190), so a pixel whose
orientation sits on a bin boundary moves its whole weight between two bins under an arbitrarily small change.  For the given oracle
keypoints of graf img1 (K = 3000, do_ori): distance of each pixel to the nearest bin boundary and its weight relative to the best bin.
    python scripts/hcori_boundary_check.py [oracle keypoint indices]"""
import sys, math, torch, numpy as np, torch.nn.functional as F
import os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import affnet_oracle as O
from helpers import gold, gray_from_rgb, load_weights
W = load_weights(); torch.set_num_threads(8)
f = gold("graf_full.npz")
oL, _, st = O.detect(gray_from_rgb(f["rgb"]), W["affnet"], None, 3000, do_ori=True, debug=True)
P = st["debug"]["ori"]["patches"]
for idx in ([int(a) for a in sys.argv[1:]] or [2707, 2738]):
    x = P[idx:idx + 1]; PS = x.size(2)
    xp = F.pad(x, (1, 1, 0, 0), "replicate"); gx = 0.5 * xp[:, :, :, :-2] - 0.5 * xp[:, :, :, 2:]
    yp = F.pad(x, (0, 0, 1, 1), "replicate"); gy = 0.5 * yp[:, :, :-2, :] - 0.5 * yp[:, :, 2:, :]
    gk = 10.0 * torch.from_numpy(O.circular_gauss_kernel(PS).astype(np.float32))
    mag = torch.sqrt(gx * gx + gy * gy + 1e-10) * gk
    o_big = 36.0 * (torch.atan2(gy, gx) + math.pi) / (2 * math.pi)
    frac = o_big - torch.floor(o_big)
    dist = torch.minimum(frac, 1 - frac).view(-1)          # distance of the pixel's orientation to a bin boundary (in bins)
    sm = O.orientation_hist_bins(x)[0]
    top = sm.topk(3)
    share = (mag.view(-1) / PS / PS) / top.values[0]      # a pixel's full weight relative to the best smoothed bin
    order = dist.argsort()[:6]
    print("keypoint %d: top bins %s values %s" % (idx, top.indices.tolist(), ["%.4e" % v for v in top.values.tolist()]))
    print("   pixels closest to a bin boundary: dist(bins) / weight share of the best bin / bin:", [("%.1e" % dist[i].item(), "%.3f" % share[i].item(), int(torch.floor(o_big).view(-1)[i].item()) % 36) for i in order.tolist()])
    print("   largest single-pixel shares:", ["%.3f" % v for v in share.topk(5).values.tolist()], " patch range %.1f..%.1f" % (x.min().item(), x.max().item()))
