"""models.rendering_tcnn of the reference (/root/reference/src/models/rendering_tcnn.py).
`render_rays` keeps the reference's signature and result dict, but between the ray rows and the
result it is ONE autograd node: fused sampler -> loner_mlp_fwd (points formed in registers) ->
loner_render_fwd, with loner_render_bwd -> loner_mlp_bwd -> loner_points_bwd as its backward."""
import torch

from loner_b200 import ops
from models.nerf_tcnn import device_grad_scale


def sample_pdf(bins, weights, N_importance, det=False, eps=1e-5):
    """Inverse-CDF sampling (rendering_tcnn.py:18-67).  Kept for API compatibility: the fused
    sampler kernel (loner_sample_ogm) contains this computation and is what render_rays uses."""
    n_rays, nb = weights.shape
    w = weights + eps
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros_like(pdf[:, :1]), torch.cumsum(pdf, -1)], -1)
    if det:
        u = torch.linspace(0, 1, N_importance, device=bins.device).expand(n_rays, N_importance)
    else:
        u = torch.rand(n_rays, N_importance, device=bins.device)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below, above = (inds - 1).clamp(min=0), inds.clamp(max=nb)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return b0 + (u - c0) / denom * (b1 - b0)


def inference(model, xyz_, dir_, sigma_only=False, netchunk=32768, detach_sigma=True, meshing=False):
    """rendering_tcnn.py:149-187 (netchunk is ignored: the kernel is tiled internally)."""
    n_rays, n_samples = xyz_.shape[0:2]
    out = model(xyz_.reshape(-1, 3).contiguous(), None if sigma_only else dir_, sigma_only, detach_sigma)
    return out if meshing else out.view(n_rays, n_samples, -1)


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays, params, z, sigma_module, raw_noise_std, seed, noise):
        n, S = z.shape
        r = rays.detach().contiguous().float()
        need = params.requires_grad or rays.requires_grad
        sigma, acts = sigma_module.fwd(n * S, need, rays=r, z=z)
        w, d, o, v = ops.render_fwd(sigma, z, r, noise=noise, raw_noise_std=raw_noise_std, seed=seed)
        ctx.save_for_backward(r, z, sigma)
        ctx.m, ctx.acts, ctx.std, ctx.seed, ctx.noise = sigma_module, acts, raw_noise_std, seed, noise
        ctx.want_rays = rays.requires_grad
        return d, w, o, v

    @staticmethod
    def backward(ctx, gd, gw, go, gv):
        r, z, sigma = ctx.saved_tensors
        n, S = z.shape
        c = lambda t: None if t is None else t.contiguous().float()
        d_sigma, d_rays = ops.render_bwd(sigma, z, r, ctx.noise, ctx.std, ctx.seed, c(gw), c(gd), c(go), c(gv))
        m = ctx.m
        d_params = torch.zeros_like(m.params)
        cs = device_grad_scale(d_sigma)            # device scalar: no host sync in the backward
        d_pos = m.bwd(n * S, (d_sigma * cs).view(-1), ctx.acts, 1.0, d_params, ctx.want_rays, rays=r, z=z)
        d_params.div_(cs)
        if ctx.want_rays:
            ops.points_bwd(d_pos.div_(cs), z, d_rays)
        return (d_rays if ctx.want_rays else None), d_params, None, None, None, None, None


_render_calls = [0]


def render_rays(rays, ray_sampler, nerf_model, ray_range, scale_factor, N_samples=64, retraw=False, perturb=0,
                white_bkgd=False, raw_noise_std=0., netchunk=32768, num_colors=3, sigma_only=False, DEBUG=False,
                detach_sigma=True, return_variance=False):
    """rendering_tcnn.py:192-267, sigma-only (LiDAR) path."""
    if not sigma_only:
        raise NotImplementedError("camera rendering is disabled in the reference (optimizer.py:433-434)")
    z_vals = ray_sampler.get_samples(rays, N_samples, perturb)
    _render_calls[0] += 1
    sm = nerf_model._model_sigma
    # parity hook: tests may attach `injected = dict(u1=, u2=, noise=)` to the sampler to replay the reference's draws
    inj = getattr(ray_sampler, "injected", None) or {}
    noise = inj.get("noise")
    depth, weights, opacity, variance = _RenderFn.apply(rays, sm.params, z_vals, sm, float(raw_noise_std),
                                                        12345 + _render_calls[0],
                                                        None if noise is None else noise.to(z_vals.device).contiguous())
    result = {'rgb_fine': torch.tensor([-1.]), 'depth_fine': depth, 'weights_fine': weights,
              'opacity_fine': opacity}
    if return_variance:
        result["variance"] = variance
    if retraw:
        result['samples_fine'] = z_vals
        result['points_fine'] = rays[:, None, 0:3] + rays[:, None, 3:6] * z_vals[:, :, None]
    return result
