"""HandCraftedModules.py counterparts on the hot path: ScalePyramid (:13-56) and HessianResp (:58-78)."""
import ctypes as C

import torch
import torch.nn as nn

from . import _lib as L


class ScalePyramid(nn.Module):
    def __init__(self, nLevels=3, init_sigma=1.6, border=5):
        super().__init__()
        self.nLevels, self.init_sigma, self.b = nLevels, init_sigma, border

    def build(self, x):
        """x CUDA float32 [B,1,H,W] -> (plan, flat pyramid buffer)."""
        x = L.f32c(x, "x")
        if x.dim() != 4 or x.size(1) != 1:
            raise L.AffnetB200Error("ScalePyramid expects [B,1,H,W]")
        plan = L.make_plan(x.size(0), x.size(2), x.size(3), self.nLevels, self.init_sigma, self.b)
        buf = torch.empty(plan.total_floats, dtype=torch.float32, device=x.device)
        L.check(L.lib().ag_pyramid_build(C.byref(plan), L.ptr(x), L.ptr(buf), L.stream_ptr()))
        return plan, buf

    @staticmethod
    def views(plan, buf):
        """Reference-shaped outputs: pyr[o][l] = tensor [B,1,h,w] (views into buf), sigmas[o][l], pix_dists[o][l]."""
        pyr, sigmas, pix = [], [], []
        for o in range(plan.n_octaves):
            h, w = plan.h[o], plan.w[o]
            n = plan.B * h * w
            pyr.append([buf[plan.level_offset[o][l]:plan.level_offset[o][l] + n].view(plan.B, 1, h, w) for l in range(plan.n_levels)])
            sigmas.append([plan.sigma[o][l] for l in range(plan.n_levels)])
            pix.append([plan.pix_dist[o]] * plan.n_levels)
        return pyr, sigmas, pix

    def forward(self, x):
        plan, buf = self.build(x)
        return self.views(plan, buf)


class HessianResp(nn.Module):
    def forward(self, x, scale):
        x = L.f32c(x, "x")
        out = torch.empty_like(x)
        L.check(L.lib().ag_hessian_response(L.ptr(x), L.ptr(out), x.size(0) * x.size(1), x.size(2), x.size(3), float(scale), 0.0, L.stream_ptr()))
        return out


def _gauss_window(PS, sigma, scale, device):
    """CircularGaussKernel(kernlen=PS, sigma=sigma) * scale as a device tensor (Utils.py:92-114)."""
    import numpy as np
    buf = np.empty(PS * PS, np.float32)
    L.check(L.lib().ag_circular_gauss_kernel(PS, float(sigma) if sigma else 0.0, buf.ctypes.data_as(C.c_void_p)))
    return (torch.from_numpy(buf) * scale).to(device)


class OrientationDetector(nn.Module):
    """HandCraftedModules.py:133-192: dominant gradient orientation of a PS x PS patch (36 bins)."""

    def __init__(self, mrSize=3.0, patch_size=None):
        super().__init__()
        self.PS = 32 if patch_size is None else patch_size
        self.mrSize = mrSize
        self.num_ang_bins = 36
        self._gk = None

    def forward(self, x, return_rot_matrix=False):
        x = L.f32c(x, "patches")
        if x.dim() != 4 or x.size(1) != 1 or x.size(2) != self.PS or x.size(3) != self.PS:
            raise L.AffnetB200Error("expected patches of shape [n,1,%d,%d]" % (self.PS, self.PS))
        if self._gk is None or self._gk.device != x.device:
            self._gk = _gauss_window(self.PS, None, 10.0, x.device)
        n = x.size(0)
        ang = torch.empty(n, dtype=torch.float32, device=x.device)
        if n:
            L.check(L.lib().ag_orientation_hist(L.ptr(x), n, self.PS, L.ptr(self._gk), L.ptr(ang), L.stream_ptr()))
        if return_rot_matrix:
            c, s = torch.cos(ang).view(-1, 1, 1), torch.sin(ang).view(-1, 1, 1)
            return torch.cat([torch.cat([c, s], dim=2), torch.cat([-s, c], dim=2)], dim=1)
        return ang


class AffineShapeEstimator(nn.Module):
    """HandCraftedModules.py:81-132: one Baumberg step (second-moment matrix -> A), up-is-up rectified."""

    def __init__(self, threshold=0.001, patch_size=19):
        super().__init__()
        self.threshold = threshold
        self.PS = patch_size
        self._gk = None

    def forward(self, x, *unused):
        x = L.f32c(x, "patches")
        if x.dim() != 4 or x.size(1) != 1 or x.size(2) != self.PS or x.size(3) != self.PS:
            raise L.AffnetB200Error("expected patches of shape [n,1,%d,%d]" % (self.PS, self.PS))
        if self._gk is None or self._gk.device != x.device:
            self._gk = _gauss_window(self.PS, (self.PS / 2) / 3.0, 1.0, x.device)
        n = x.size(0)
        A = torch.empty(n, 2, 2, dtype=torch.float32, device=x.device)
        if n:
            L.check(L.lib().ag_baumberg_shape(L.ptr(x), n, self.PS, L.ptr(self._gk), L.ptr(A), L.stream_ptr()))
        return A
