"""Utils.py counterparts: GaussianBlur (Utils.py:150-166) and line_prepender (Utils.py:177-182)."""
import torch
import torch.nn as nn

from . import _lib as L


class GaussianBlur(nn.Module):
    """k = int(6 sigma + 1) | 1 taps at linspace(-k/2, k/2, k), replicate padding; x: CUDA float32 [B,1,h,w]."""

    def __init__(self, sigma=1.6):
        super().__init__()
        self.sigma = float(sigma)

    def forward(self, x):
        x = L.f32c(x, "x")
        if x.dim() != 4 or x.size(1) != 1:
            raise L.AffnetB200Error("GaussianBlur expects [B,1,h,w]")
        out = torch.empty_like(x)
        L.check(L.lib().ag_gaussian_blur(L.ptr(x), L.ptr(out), x.size(0), x.size(2), x.size(3), self.sigma, L.stream_ptr()))
        return out


def line_prepender(filename, line):
    with open(filename, "r+") as f:
        content = f.read()
        f.seek(0, 0)
        f.write(line.rstrip("\r\n") + "\n" + content)
