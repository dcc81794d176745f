"""Shared test helpers (fixtures loading, keypoint matching)."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    return np.load(os.path.join(GOLD, name))


def load_weights():
    """-> dict(affnet=sd, orinet=sd, hardnet=sd) of float32 torch tensors (checkpoint key names)."""
    z = gold("weights.npz")
    out = {"affnet": {}, "orinet": {}, "hardnet": {}}
    for k in z.files:
        net, key = k.split("/", 1)
        out[net][key] = torch.from_numpy(z[k])
    return out


def gray_from_rgb(rgb):
    """hesaffnet.py:35-39: mean over RGB -> float32 [1,1,H,W]."""
    return torch.from_numpy(np.mean(rgb, axis=2).astype(np.float32)).view(1, 1, rgb.shape[0], rgb.shape[1])


def match_keypoints(LA, LB, tol_px=0.05):
    """Greedy one-to-one matching of LAF centres. Returns (idxA, idxB) of matched pairs."""
    a = LA[:, :, 2].double().numpy(); b = LB[:, :, 2].double().numpy()
    from scipy.spatial import cKDTree
    tree = cKDTree(b)
    d, j = tree.query(a, k=1)
    ok = d <= tol_px
    ia = np.nonzero(ok)[0]; ib = j[ok]
    # scale check disambiguates same-centre detections from different levels
    keep = []
    used = set()
    for x, y in zip(ia, ib):
        if y in used:
            continue
        sa = abs(np.linalg.det(LA[x, :, :2].double().numpy())) ** 0.5
        sb = abs(np.linalg.det(LB[y, :, :2].double().numpy())) ** 0.5
        if abs(sa - sb) <= 0.05 * max(sa, 1.0):
            keep.append((x, y)); used.add(y)
    if not keep:
        return np.zeros(0, int), np.zeros(0, int)
    k = np.array(keep)
    return k[:, 0], k[:, 1]
