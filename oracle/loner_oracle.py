"""TEST INFRASTRUCTURE — CPU restatement (torch, fp32) of LONER's mapping hot path.

Every function follows one reference function, cited as file:line under /root/reference/src.
This file is self-contained (it travels to the GPU box, where /root/reference does not exist)
and is PINNED by tests/test_oracle_vs_golden.py against tests/golden/*.npz, which
oracle/make_golden.py minted by executing the reference's own Python in this container
(oracle/ref_harness.py).  Two sub-parts stay "parity unpinned" because the reference delegates
them to un-vendored packages with no test of its own: the tcnn network arithmetic
(oracle/tcnn_standin.py) and pytorch3d's axis-angle maths (oracle/p3d_standin.py).

Only tests/, bench.py's cpu_baseline / --impl reference leg and __graft_entry__.smoke() may
import this module; the product path never does.
"""
import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

from . import hashgrid_standin, p3d_standin, tcnn_standin


# ----------------------------------------------------------------------------- poses / rays
def pose6_to_matrix(pose6: torch.Tensor) -> torch.Tensor:
    """[6]=[t, axis-angle] -> 4x4.  common/pose_utils.py:288-302 (tensor_to_transform)."""
    R = p3d_standin.axis_angle_to_matrix(pose6[3:][None])[0]
    top = torch.cat([R, pose6[:3, None]], dim=1)
    bottom = torch.tensor([[0.0, 0.0, 0.0, 1.0]], dtype=top.dtype)
    return torch.cat([top, bottom], dim=0)


def get_far_val(o: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """Exit distance from the cube [-1,1]^3.  common/ray_utils.py:31-60 with no_nan=True."""
    d = d + 1e-15
    t_neg = (-1.0 - o) / d
    t_pos = (1.0 - o) / d
    per_axis = torch.maximum(t_neg.clamp(min=0), t_pos.clamp(min=0))
    return per_axis.min(dim=1, keepdim=True)[0]


def build_lidar_rays(directions, distances, idx, pose, ray_range, scale, shift):
    """common/ray_utils.py:269-322.  directions [3,M], distances [M], idx [n] int64, pose 4x4.
    Returns rays [n',13], depths [n'], keep-mask [n] (rows with far > near + 1/scale)."""
    depths = distances[idx] / scale
    dirs = directions[:, idx]
    origin = (pose[:3, 3] + shift) / scale
    n = idx.shape[0]
    origins = origin.tile(n, 1)
    rd = (pose[:3, :3] @ dirs).T
    rd = rd / torch.norm(rd, dim=1, keepdim=True)
    near = (ray_range[0] / scale) * torch.ones_like(origins[:, :1])
    far_range = (ray_range[1] / scale) * torch.ones_like(origins[:, :1])
    far = torch.minimum(far_range, get_far_val(origins, rd))
    rays = torch.cat([origins, rd, -rd, torch.zeros_like(origins[:, :2]), near, far], dim=1)
    keep = (far > near + 1.0 / scale)[:, 0]
    return rays[keep], depths[keep], keep


# ----------------------------------------------------------------------------- sampling
def stratified(near, far, H, perturb, u):
    """First half of models/ray_sampling.py:59-73 (same as UniformRaySampler :22-43)."""
    t = torch.linspace(0, 1, H)
    z = near * (1 - t) + far * t
    if perturb > 0:
        mid = 0.5 * (z[:, :-1] + z[:, 1:])
        upper = torch.cat([mid, z[:, -1:]], -1)
        lower = torch.cat([z[:, :1], mid], -1)
        z = lower + (upper - lower) * (perturb * u)
    return z


def uniform_samples(rays, S, perturb, u=None):
    """models/ray_sampling.py:22-43."""
    return stratified(rays[:, -2:-1], rays[:, -1:], S, perturb, u)


def ogm_interpolate(grid, pts):
    """models/model_tcnn.py:124-131: trilinear, align_corners=False, zero padding;
    query (x,y,z) reads grid[z,y,x]."""
    n, s, _ = pts.shape
    return F.grid_sample(grid, pts.reshape(1, 1, n, s, 3), mode="bilinear",
                         align_corners=False).reshape(n, s)


def sample_pdf(bins, weights, n_importance, u, eps=1e-5):
    """models/rendering_tcnn.py:18-67 with det=False and the uniform draws `u` injected."""
    n_rays, nb = weights.shape
    w = weights + eps
    pdf = w / w.sum(-1, keepdim=True)
    cdf = torch.cat([torch.zeros(n_rays, 1), torch.cumsum(pdf, -1)], -1)
    u = u.contiguous()
    inds = torch.searchsorted(cdf, u, right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=nb)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < eps, torch.ones_like(denom), denom)
    return b0 + (u - c0) / denom * (b1 - b0)


def ogm_samples(rays, grid, S, perturb, u1, u2):
    """models/ray_sampling.py:53-92.  u1 [N,S/2] (stratified jitter), u2 [N,S/2] (inverse CDF)."""
    o, d = rays[:, 0:3], rays[:, 3:6]
    H = S // 2
    z = stratified(rays[:, -2:-1], rays[:, -1:], H, perturb, u1)
    pts = o[:, None, :] + d[:, None, :] * z[:, :, None]
    logits = ogm_interpolate(grid, pts)
    probs = 1.0 / (1 + torch.exp(-logits))
    probs = 2 * (probs.clamp(min=0.5, max=1.0) - 0.5)
    mid = 0.5 * (z[:, :-1] + z[:, 1:])
    z_imp = sample_pdf(mid, probs[:, 1:-1], H, u2)
    return torch.sort(torch.cat([z, z_imp], -1), -1)[0]


# ----------------------------------------------------------------------------- network
@dataclass
class NetSpec:
    """Sigma head (models/nerf_tcnn.py:35-38): Frequency encoding + bias-free ReLU MLP, or - `hash` given -
    the shipped configuration, multiresolution hash encoding + MLP (cfg/nerf_config/default_nerf_hash.yaml).
    Flat params in tcnn's order: network matrices, then the hash table."""
    n_frequencies: int = 10
    n_neurons: int = 256
    n_hidden_layers: int = 4
    precision: str = "fp16"
    hash: object = None            # oracle.hashgrid_standin.HashGridSpec or None
    shapes: list = field(init=False)

    def __post_init__(self):
        width = self.hash.n_output_dims if self.hash is not None else 3 * 2 * self.n_frequencies
        e_pad = (width + 15) // 16 * 16
        self.e_pad = e_pad
        self.shapes = tcnn_standin.mlp_layer_shapes(e_pad, self.n_neurons, self.n_hidden_layers, 16)

    @property
    def n_network_params(self):
        return sum(a * b for a, b in self.shapes)

    @property
    def n_params(self):
        return self.n_network_params + (self.hash.n_params if self.hash is not None else 0)


def sigma_net(pos, params, spec: NetSpec):
    """models/nerf_tcnn.py:59-78 with sigma_only=True: pos in [-1,1] -> sigma [P]."""
    x = (pos + 1) / 2
    if spec.hash is not None:
        enc = hashgrid_standin.hashgrid_encode(x, params[spec.n_network_params:], spec.hash, spec.precision)
        if enc.shape[1] < spec.e_pad:
            enc = torch.cat([enc, torch.ones(enc.shape[0], spec.e_pad - enc.shape[1])], dim=1)
    else:
        enc = tcnn_standin.frequency_encode(x, spec.n_frequencies, pad_to=16)
    out = tcnn_standin.mlp_forward(enc, params[:spec.n_network_params], spec.shapes, spec.precision)
    sigma = out[:, 0]
    finfo = torch.finfo(torch.float16)
    if not torch.isfinite(sigma).all():
        sigma = sigma.nan_to_num(posinf=finfo.max, neginf=finfo.min)
    return sigma


# ----------------------------------------------------------------------------- render
def raw2outputs(sigma, z, rays_d, noise, far):
    """models/rendering_tcnn.py:93-145 (sigma_only, far given, ret_var).  sigma,z,noise [N,S]."""
    deltas = z[:, 1:] - z[:, :-1]
    deltas = torch.cat([deltas, 1e10 * torch.ones_like(deltas[:, :1])], -1)
    deltas = deltas * torch.norm(rays_d[:, None, :], dim=-1)
    alphas = 1 - torch.exp(-deltas * torch.relu(sigma + noise))
    shifted = torch.cat([torch.ones_like(alphas[:, :1]), 1.0 - alphas + 1e-10], -1)
    weights = alphas * torch.cumprod(shifted, -1)[:, :-1]
    opacity = weights.sum(-1)
    z_app = torch.cat([z, far], dim=-1)
    w_app = torch.cat([weights, 1 - weights.sum(dim=1, keepdim=True)], dim=1)
    depth = (w_app * z_app).sum(-1)
    variance = (weights * (depth.view(-1, 1) - z) ** 2).sum(dim=1)
    return depth, weights, opacity, variance


def render_rays(rays, z, params, spec, noise):
    """models/rendering_tcnn.py:192-267 (sigma-only)."""
    o, d = rays[:, 0:3], rays[:, 3:6]
    xyz = o[:, None, :] + d[:, None, :] * z[:, :, None]
    sigma = sigma_net(xyz.reshape(-1, 3), params, spec).view(z.shape)
    depth, weights, opacity, variance = raw2outputs(sigma, z, d, noise, rays[:, -1:])
    return dict(depth_fine=depth, weights_fine=weights, opacity_fine=opacity, variance=variance,
                samples_fine=z, points_fine=xyz, sigma=sigma)


# ----------------------------------------------------------------------------- loss
@dataclass
class LossCfg:
    """cfg/model_config/default_model_config.yaml:40-58."""
    min_depth_eps: float = 0.5
    min_js: float = 1.0
    max_js: float = 10.0
    alpha: float = 1.0
    los_lambda: float = 1000.0
    depthloss_lambda: float = 0.005
    loss_selection: str = "L1_JS"      # L1_JS | L2_JS | L1_LOS | L2_LOS
    fixed_eps: float = 3.0             # the *_LOS margin for this iteration (optimizer.py:516-521)


def _kl(m1, s1, m2, s2):
    """mapping/optimizer.py:614-621."""
    return torch.log(s2 / s1) + (s1 * s1 + (m1 - m2) ** 2) / (2 * s2 * s2) - 0.5


def js_divergence(m1, s1, m2, s2):
    """mapping/optimizer.py:623-626."""
    mm = 0.5 * (m1 + m2)
    sm = 0.5 * torch.sqrt(s1 ** 2 + s2 ** 2)
    return 0.5 * _kl(m1, s1, mm, sm) + 0.5 * _kl(m2, s2, mm, sm)


def get_weights_gt(s, gt, eps):
    """models/losses.py:29-51 (norm=True)."""
    sg = eps / 3
    a = (gt - eps - gt) / sg
    b = (gt + eps - gt) / sg
    cdf = lambda x: 0.5 * (1 + torch.erf(x / math.sqrt(2)))
    pdf = (1.0 / math.sqrt(2 * math.pi)) * torch.exp(-0.5 * ((s - gt) / sg) ** 2)
    w = pdf / sg / (cdf(b) - cdf(a))
    zero = torch.zeros_like(s)
    w = torch.heaviside(s - (gt - eps), zero) * torch.heaviside((gt + eps) - s, zero) * w
    return w / (w.sum(dim=1, keepdim=True) + 1e-6)


def compute_loss(rays, depths, res, scale, cfg: LossCfg):
    """mapping/optimizer.py:437-595, loss_selection L1_JS.  depths [N] in cube units."""
    gt = depths.reshape(-1, 1)
    far = rays[:, -1]
    transparent = (gt > far[:, None])[:, 0]
    opaque = (gt > 0)[:, 0] & ~transparent
    s = res["samples_fine"] * scale
    G = gt * scale
    w = res["weights_fine"]
    wsum = w.sum(1)
    mean = (s * w).sum(1) / (wsum + 1e-10)
    var = ((s - mean[:, None]) ** 2 * w).sum(1) / (wsum + 1e-10) + 1e-10
    std = torch.sqrt(var)
    js = js_divergence(G, cfg.min_depth_eps / 3.0, mean[:, None], std[:, None]).squeeze(-1)
    js_raw = js.detach().clone()
    depth_m = res["depth_fine"][:, None] * scale
    depth_loss = F.mse_loss(depth_m[opaque, 0], G[opaque, 0])
    js_c = js.detach().clone()
    js_c[js_c < cfg.min_js] = 0
    js_c[js_c > cfg.max_js] = cfg.max_js
    if cfg.loss_selection.endswith("JS"):
        eps_dyn = (cfg.min_depth_eps * (1 + cfg.alpha * js_c))[:, None]            # optimizer.py:497-503
    else:
        eps_dyn = torch.full_like(G, cfg.fixed_eps)                                  # optimizer.py:516-523
    w_gt = get_weights_gt(s.detach(), G, eps_dyn)
    w_gt[~opaque, :] = 0
    los = F.l1_loss(w, w_gt) if cfg.loss_selection.startswith("L1") else F.mse_loss(w, w_gt)   # optimizer.py:568-574
    opacity_loss = (res["opacity_fine"][opaque] - 1).abs().mean()
    loss = cfg.depthloss_lambda * depth_loss + cfg.los_lambda * los + opacity_loss
    return dict(loss=loss, depth_loss=depth_loss, los_loss=los, opacity_loss=opacity_loss,
                eps_dynamic=eps_dyn[:, 0], js=js_raw, std=std.detach(), mean=mean.detach(),
                weights_gt=w_gt, opaque=opaque, depth_eps_mean=float(eps_dyn.mean()))


def depth_l1_metric(depth_fine, depths, scale, ray_range):
    """analysis/compute_l1_depth.py:59-64: L1 on rays with r0 < gt < r1 - 0.25 (metres)."""
    gt = depths * scale
    ok = (gt > ray_range[0]) & (gt < ray_range[1] - 0.25)
    return F.l1_loss(depth_fine[ok] * scale, gt[ok])


# ----------------------------------------------------------------------------- occupancy grid
def get_logits_grad(s, gt, eps=2.0, l_free=0.25, l_occ=2.5):
    """models/losses.py:54-62 (heaviside(0)=0)."""
    x = s - gt
    return l_free * (x < -eps).float() - l_occ * ((x > -eps) & (x < eps)).float()


def occupancy_step(grid, points, s, gt, lr):
    """mapping/optimizer.py:598-609: one SGD step of the logit grid with the pseudo-gradient."""
    g = grid.detach().clone().requires_grad_(True)
    logits = ogm_interpolate(g, points.detach())
    logits.backward(gradient=get_logits_grad(s, gt))
    return (g - lr * g.grad).detach()


# ----------------------------------------------------------------------------- whole step
def adam_update(p, g, m, v, step, lr, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam defaults as constructed at mapping/optimizer.py:257-267."""
    m = b1 * m + (1 - b1) * g
    v = b2 * v + (1 - b2) * g * g
    mh = m / (1 - b1 ** step)
    vh = v / (1 - b2 ** step)
    return p - lr * mh / (vh.sqrt() + eps), m, v


def mapping_iteration(scans, poses6, idx_per_kf, params, spec, grid, S, scale, shift, ray_range,
                      perturb, u1, u2, noise, cfg: LossCfg, sampler="OGM", sky=None):
    """One pass of the hot loop mapping/optimizer.py:276-380 up to loss.backward():
    ray build per keyframe -> concat -> sample -> net -> render -> loss -> grads.
    poses6: list of [6] tensors (leaf, may require grad); params: flat fp32 (requires grad).
    sky: optional (sky_dirs per keyframe [3,Ks], sky_idx per keyframe [n_sky]) - the keyframe's sky rays are
    appended after its lidar rays, at distance ray_range[1] + 1 and built from the DETACHED pose
    (mapping/keyframe.py:87-99, common/sensors.py:162-167, mapping/optimizer.py:299-305)."""
    rays_l, depth_l = [], []
    for k, (sc, p6, idx) in enumerate(zip(scans, poses6, idx_per_kf)):
        r, dep, _ = build_lidar_rays(sc.ray_directions, sc.distances, idx, pose6_to_matrix(p6),
                                     ray_range, scale, shift)
        rays_l.append(r)
        depth_l.append(dep)
        if sky is not None and sky[1][k] is not None:
            dirs = sky[0][k]
            r, dep, _ = build_lidar_rays(dirs, torch.full_like(dirs[0], ray_range[1] + 1), sky[1][k],
                                         pose6_to_matrix(p6.detach()), ray_range, scale, shift)
            rays_l.append(r)
            depth_l.append(dep)
    rays = torch.cat(rays_l).float()
    depths = torch.cat(depth_l).float()
    with torch.no_grad():
        if sampler == "OGM":
            z = ogm_samples(rays, grid, S, perturb, u1, u2)
        else:
            z = uniform_samples(rays, S, perturb, u1)
    res = render_rays(rays, z, params, spec, noise)
    out = compute_loss(rays, depths, res, scale, cfg)
    return rays, depths, res, out
