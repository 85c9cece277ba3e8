"""models.losses of the reference (/root/reference/src/models/losses.py), same names and
semantics.  These two are elementwise host-level helpers the reference's Optimizer.compute_loss
calls on device tensors; the fused kernel loner_render_loss computes the same quantities inside
the engine (loner_b200/csrc/render.cu)."""
import math

import torch


def img_to_mse(x, y):
    return torch.mean((x - y) ** 2)


def mse_to_psnr(x):
    return -10.0 * torch.log(x) / math.log(10.0)


def get_weights_gt(sampled_depth, gt_depth, eps, norm=True):
    """Truncated Gaussian target around the measured depth (losses.py:29-51)."""
    sigma = eps / 3
    lo = (gt_depth - eps - gt_depth) / sigma
    hi = (gt_depth + eps - gt_depth) / sigma

    def cdf(x):
        return 0.5 * (1 + torch.erf(x / math.sqrt(2)))

    x = (sampled_depth - gt_depth) / sigma
    w = (1.0 / math.sqrt(2 * math.pi)) * torch.exp(-0.5 * x ** 2) / sigma / (cdf(hi) - cdf(lo))
    inside = ((sampled_depth - (gt_depth - eps)) > 0) & (((gt_depth + eps) - sampled_depth) > 0)
    w = w * inside.to(w.dtype)
    if norm:
        w = w / (w.sum(dim=1, keepdim=True) + 1e-6)
    return w


def get_logits_grad(z_vals, depth, eps=2, l_free=0.25, l_occ=2.5):
    """Occupancy-grid pseudo-gradient (losses.py:54-62); heaviside(0) = 0."""
    x = z_vals - depth
    return l_free * (x < -eps).to(x.dtype) - l_occ * ((x > -eps) & (x < eps)).to(x.dtype)
