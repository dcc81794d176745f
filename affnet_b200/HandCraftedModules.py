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
