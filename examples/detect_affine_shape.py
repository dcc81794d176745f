#!/usr/bin/env python
"""Counterpart of the reference's examples/just_shape/detect_affine_shape.py (BASELINE.json configs[0]):

    python examples/detect_affine_shape.py patches.png out.txt

`patches.png` is a column of square grayscale patches (w x n*w); every patch is resized to 32x32, scaled to 0..1 and run
through AffNetFast; the output rows are `a11 0 a21 a22` (%10.5f), detect_affine_shape.py:36-70.
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from affnet_b200.architectures import AffNetFast  # noqa: E402


def main():
    import cv2
    src, dst = sys.argv[1], sys.argv[2]
    weights = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "weights.npz")
    img = cv2.imread(src, 0)
    h, w = img.shape
    n = h // w
    patches = np.stack([cv2.resize(img[i * w:(i + 1) * w], (32, 32), interpolation=cv2.INTER_LINEAR) for i in range(n)])
    x = torch.from_numpy(patches.astype(np.float32) / 255.0).view(n, 1, 32, 32).cuda()
    net = AffNetFast(PS=32)
    if weights.endswith(".npz"):
        z = np.load(weights)
        net.load_state_dict({k.split("/", 1)[1]: torch.from_numpy(z[k]) for k in z.files if k.startswith("affnet/")})
    else:
        net.load_state_dict(torch.load(weights, map_location="cpu")["state_dict"])
    net = net.eval().cuda()
    with torch.no_grad():
        A = net(x).cpu().numpy().reshape(n, 4)
    np.savetxt(dst, A, delimiter=" ", fmt="%10.5f")
    print("%d shapes -> %s" % (n, dst))


if __name__ == "__main__":
    main()
