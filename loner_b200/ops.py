"""One Python function per C-ABI entry point: allocates outputs with torch (the library never
allocates), passes raw device pointers and the current CUDA stream.  Plumbing only."""
import ctypes
import math

import torch

from . import lib as L

RAY_COLS = 13
FLAG_VALID, FLAG_OPAQUE = 1, 2


def _f32(t):
    assert t.dtype == torch.float32 and t.is_cuda, "expected a float32 CUDA tensor"
    return t.contiguous()


NET_SINGLE_CTA, NET_STASH_DZL, NET_ONE_ISSUER, NET_DGRAD_ONE_ISSUER, NET_STASH_AL, NET_WG_PLAN_BYTES = 1, 2, 32, 64, 128, 256   # LONER_NET_*
HASH_SCALAR = 1                            # LONER_HASH_*
DEFAULT_NET_FLAGS = 0                       # production: CTA pairs, dZ_L rebuilt inside wgrad, dW_out folded (no A_L stash)


class Net:
    """Sigma-head description (Frequency encoding + bias-free ReLU MLP) mirrored from the
    reference's nerf_config keys (models/nerf_tcnn.py:29-38)."""

    def __init__(self, n_frequencies=10, n_neurons=256, n_hidden_layers=4, flags=None):
        """flags: LONER_NET_* bits of include/loner_b200.h (kernel A/B variants); None = DEFAULT_NET_FLAGS."""
        self.flags = DEFAULT_NET_FLAGS if flags is None else int(flags)
        self.c = L.NetT(int(n_frequencies), int(n_neurons), int(n_hidden_layers), self.flags)
        self.n_frequencies, self.n_neurons, self.n_hidden_layers = int(n_frequencies), int(n_neurons), int(n_hidden_layers)
        lib = L.load()
        self.param_count = lib.loner_mlp_param_count(ctypes.byref(self.c))
        if self.param_count < 0:
            raise RuntimeError(f"unsupported sigma network {n_frequencies=} {n_neurons=} {n_hidden_layers=} "
                               "(kernels implement Frequency<=10, width 64/128/256, 1..8 hidden layers)")
        self.packed_bytes = lib.loner_mlp_packed_bytes(ctypes.byref(self.c))
        self.e_pad = (6 * self.n_frequencies + 15) // 16 * 16

    def ref(self):
        return ctypes.byref(self.c)

    @property
    def stashes_last_activation(self):
        """False when the forward leaves A_L out of the stash: wgrad then derives dW_out from the last hidden layer's
        weight-gradient partials (mlp.cu `wgrad_folds_out`); the slot stays in the layout, unwritten."""
        return self.n_hidden_layers < 2 or bool(self.flags & (NET_STASH_DZL | NET_STASH_AL))

    def act_bytes(self, P):
        return L.load().loner_mlp_act_bytes(self.ref(), P)

    def bwd_scratch_bytes(self, P):
        return L.load().loner_mlp_bwd_scratch_bytes(self.ref(), P)

    def layer_shapes(self):
        s = [(self.n_neurons, self.e_pad)]
        s += [(self.n_neurons, self.n_neurons)] * (self.n_hidden_layers - 1)
        s.append((16, self.n_neurons))
        return s


class HashNet:
    """The reference's shipped sigma head: multiresolution hash encoding + 1 x 64 MLP
    (cfg/nerf_config/default_nerf_hash.yaml `pos_encoding_sigma` / `sigma_network`, models/nerf_tcnn.py:35-38).
    Flat fp32 params in tcnn's order: W1 [64, e_pad] | W_2 .. W_L [64, 64] | W_out [16, 64] | table [entries, 2]."""

    def __init__(self, n_levels=16, n_features_per_level=2, log2_hashmap_size=18, base_resolution=16,
                 per_level_scale=2.0, n_neurons=64, n_hidden_layers=1, flags=0):
        """flags: LONER_HASH_* bits (HASH_SCALAR = round 1's CUDA-core head, kept for A/B)."""
        self.flags = int(flags)
        self.c = L.HashNetT(int(n_levels), int(n_features_per_level), int(log2_hashmap_size), int(base_resolution),
                            float(per_level_scale), int(n_neurons), int(n_hidden_layers), self.flags)
        lib = L.load()
        self.param_count = lib.loner_hash_param_count(ctypes.byref(self.c))
        if self.param_count < 0:
            raise RuntimeError(f"unsupported hash-grid sigma network {n_levels=} {n_features_per_level=} "
                               f"{log2_hashmap_size=} {n_neurons=} {n_hidden_layers=} (kernels implement <=16 levels "
                               "x 2 features, 1-4 hidden layers of 64 neurons)")
        self.n_levels, self.n_neurons, self.n_hidden_layers = int(n_levels), int(n_neurons), int(n_hidden_layers)
        self.e_pad = (2 * self.n_levels + 15) // 16 * 16
        self.table_entries = lib.loner_hash_table_entries(ctypes.byref(self.c))
        self.packed_bytes = lib.loner_hash_packed_bytes(ctypes.byref(self.c))
        self.n_network_params = sum(o * i for o, i in self.layer_shapes())

    def ref(self):
        return ctypes.byref(self.c)

    def bwd_scratch_bytes(self, P):
        return L.load().loner_hash_bwd_scratch_bytes(self.ref(), P)

    def layer_shapes(self):
        return ([(self.n_neurons, self.e_pad)] + [(self.n_neurons, self.n_neurons)] * (self.n_hidden_layers - 1) +
                [(16, self.n_neurons)])


def hash_pack(net: HashNet, params, packed=None):
    if packed is None:
        packed = torch.empty(net.packed_bytes, device=params.device, dtype=torch.uint8)
    L.check(L.load().loner_hash_pack(net.ref(), L.ptr(_f32(params)), L.ptr(packed), L.stream_ptr()), "loner_hash_pack")
    return packed


def hash_fwd(net: HashNet, packed, P, pos=None, rays=None, z=None, sigma=None):
    if sigma is None:
        sigma = torch.empty(P, device=packed.device, dtype=torch.float32)
    S = z.shape[1] if z is not None else 1
    L.check(L.load().loner_hash_fwd(net.ref(), L.ptr(packed), L.ptr(pos), L.ptr(rays), L.ptr(z), S, P, L.ptr(sigma),
                                    L.stream_ptr()), "loner_hash_fwd")
    return sigma


def hash_bwd(net: HashNet, packed, P, d_sigma, grad_scale, d_params, pos=None, rays=None, z=None, want_dpos=False,
             scratch=None):
    dev = packed.device
    if scratch is None:
        scratch = torch.empty(max(net.bwd_scratch_bytes(P), 16), device=dev, dtype=torch.uint8)
    d_pos = torch.empty(P, 3, device=dev, dtype=torch.float32) if want_dpos else None
    S = z.shape[1] if z is not None else 1
    L.check(L.load().loner_hash_bwd(net.ref(), L.ptr(packed), L.ptr(pos), L.ptr(rays), L.ptr(z), S, P,
                                    L.ptr(_f32(d_sigma)), float(grad_scale), L.ptr(d_params), L.ptr(d_pos),
                                    L.ptr(scratch), L.stream_ptr()), "loner_hash_bwd")
    return d_pos


def pack_points(ray_directions, distances):
    """[3,M] + [M] -> [M,4] float4 rows (dx,dy,dz,dist): one 16-byte load per picked ray."""
    return torch.cat([ray_directions.t(), distances[:, None]], dim=1).contiguous().float()


KF_DETACHED = 0x40000000
KF_MASK = 0x3FFFFFFF
MLP_BWD_LAUNCHES = 3          # dgrad, wgrad (incl. dW_out), partial-sum reduce
PICK_RANDOM, PICK_FIXED, PICK_MASK = 0, 1, 2


def pick_segments(segs, device):
    """[(kf, mode, base, size, map_off, out_begin)] -> device array of loner_pick_seg_t (uint8 tensor)."""
    arr = (L.PickSegT * len(segs))(*[L.PickSegT(*[int(v) for v in s]) for s in segs])
    raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
    return raw.to(device)


def ray_pick(segs_dev, n_segs, index_map, seed, n, ray_kf=None, ray_point=None):
    dev = segs_dev.device
    if ray_kf is None:
        ray_kf = torch.empty(n, device=dev, dtype=torch.int32)
    if ray_point is None:
        ray_point = torch.empty(n, device=dev, dtype=torch.int64)
    L.check(L.load().loner_ray_pick(L.ptr(segs_dev), n_segs, L.ptr(index_map), seed, n, L.ptr(ray_kf), L.ptr(ray_point),
                                    L.stream_ptr()), "loner_ray_pick")
    return ray_kf, ray_point


def ray_build(points, ray_kf, ray_point, poses12, shift, scale, ray_range, counters=None):
    n = ray_point.shape[0]
    dev = points.device
    rays = torch.empty(n, RAY_COLS, device=dev, dtype=torch.float32)
    depths = torch.empty(n, device=dev, dtype=torch.float32)
    flags = torch.empty(n, device=dev, dtype=torch.uint8)
    sh = L.host_floats(shift)
    L.check(L.load().loner_ray_build(L.ptr(points), L.ptr(ray_kf), L.ptr(ray_point), n, L.ptr(_f32(poses12)),
                                     poses12.shape[0], sh, float(scale), float(ray_range[0]), float(ray_range[1]),
                                     L.ptr(rays), L.ptr(depths), L.ptr(flags), L.ptr(counters), L.stream_ptr()),
            "loner_ray_build")
    return rays, depths, flags


def ray_build_bwd(points, ray_kf, ray_point, poses12, shift, scale, ray_range, d_rays, out=None):
    """out: optional zeroed [K,12] buffer the gradient is accumulated into."""
    d_poses = torch.zeros_like(poses12) if out is None else out
    sh = L.host_floats(shift)
    L.check(L.load().loner_ray_build_bwd(L.ptr(points), L.ptr(ray_kf), L.ptr(ray_point), ray_point.shape[0],
                                         L.ptr(_f32(poses12)), poses12.shape[0], sh, float(scale),
                                         float(ray_range[1]), L.ptr(_f32(d_rays)), L.ptr(d_poses), L.stream_ptr()),
            "loner_ray_build_bwd")
    return d_poses


def sample_uniform(rays, S, perturb, u=None, seed=0):
    n = rays.shape[0]
    z = torch.empty(n, S, device=rays.device, dtype=torch.float32)
    L.check(L.load().loner_sample_uniform(L.ptr(_f32(rays)), n, S, float(perturb), L.ptr(u), seed, L.ptr(z),
                                          L.stream_ptr()), "loner_sample_uniform")
    return z


def sample_ogm(rays, grid, S, perturb, u1=None, u2=None, seed=0):
    n = rays.shape[0]
    V = grid.shape[-1]
    z = torch.empty(n, S, device=rays.device, dtype=torch.float32)
    L.check(L.load().loner_sample_ogm(L.ptr(_f32(rays)), n, S, float(perturb), L.ptr(_f32(grid)), V, L.ptr(u1),
                                      L.ptr(u2), seed, L.ptr(z), L.stream_ptr()), "loner_sample_ogm")
    return z


def mlp_pack(net: Net, params, packed=None):
    if packed is None:
        packed = torch.empty(net.packed_bytes, device=params.device, dtype=torch.uint8)
    L.check(L.load().loner_mlp_pack(net.ref(), L.ptr(_f32(params)), L.ptr(packed), L.stream_ptr()), "loner_mlp_pack")
    return packed


def mlp_fwd(net: Net, packed, P, pos=None, rays=None, z=None, stash=False, sigma=None, acts=None):
    dev = packed.device
    if sigma is None:
        sigma = torch.empty(P, device=dev, dtype=torch.float32)
    if stash and acts is None:
        acts = torch.empty(net.act_bytes(P), device=dev, dtype=torch.uint8)
    S = z.shape[1] if z is not None else 1
    L.check(L.load().loner_mlp_fwd(net.ref(), L.ptr(packed), L.ptr(pos), L.ptr(rays), L.ptr(z), S, P, L.ptr(sigma),
                                   L.ptr(acts) if stash else None, L.stream_ptr()), "loner_mlp_fwd")
    return sigma, acts


def mlp_bwd(net: Net, packed, P, d_sigma, acts, grad_scale, d_params, pos=None, rays=None, z=None, want_dpos=False,
            scratch=None):
    dev = packed.device
    if scratch is None:
        scratch = torch.empty(net.bwd_scratch_bytes(P), device=dev, dtype=torch.uint8)
    d_pos = torch.empty(P, 3, device=dev, dtype=torch.float32) if want_dpos else None
    S = z.shape[1] if z is not None else 1
    L.check(L.load().loner_mlp_bwd(net.ref(), L.ptr(packed), L.ptr(pos), L.ptr(rays), L.ptr(z), S, P,
                                   L.ptr(_f32(d_sigma)), L.ptr(acts), float(grad_scale), L.ptr(d_params),
                                   L.ptr(d_pos), L.ptr(scratch), L.stream_ptr()), "loner_mlp_bwd")
    return d_pos


def mlp_dgrad(net: Net, packed, P, d_sigma, acts, grad_scale, scratch, pos=None, rays=None, z=None, want_dpos=False):
    d_pos = torch.empty(P, 3, device=packed.device, dtype=torch.float32) if want_dpos else None
    S = z.shape[1] if z is not None else 1
    L.check(L.load().loner_mlp_dgrad(net.ref(), L.ptr(packed), L.ptr(pos), L.ptr(rays), L.ptr(z), S, P,
                                     L.ptr(_f32(d_sigma)), L.ptr(acts), float(grad_scale), L.ptr(d_pos),
                                     L.ptr(scratch), L.stream_ptr()), "loner_mlp_dgrad")
    return d_pos


def mlp_wgrad(net: Net, packed, P, d_sigma, acts, grad_scale, d_params, scratch):
    L.check(L.load().loner_mlp_wgrad(net.ref(), L.ptr(packed), P, L.ptr(_f32(d_sigma)), L.ptr(acts),
                                     float(grad_scale), L.ptr(d_params), L.ptr(scratch), L.stream_ptr()),
            "loner_mlp_wgrad")


def render_fwd(sigma, z, rays, noise=None, raw_noise_std=0.0, seed=0, want_weights=True):
    n, S = z.shape
    dev = z.device
    weights = torch.empty(n, S, device=dev, dtype=torch.float32) if want_weights else None
    depth = torch.empty(n, device=dev, dtype=torch.float32)
    opacity = torch.empty(n, device=dev, dtype=torch.float32)
    variance = torch.empty(n, device=dev, dtype=torch.float32)
    L.check(L.load().loner_render_fwd(L.ptr(_f32(sigma)), L.ptr(_f32(z)), L.ptr(_f32(rays)), n, S, L.ptr(noise),
                                      float(raw_noise_std), seed, L.ptr(weights), L.ptr(depth), L.ptr(opacity),
                                      L.ptr(variance), L.stream_ptr()), "loner_render_fwd")
    return weights, depth, opacity, variance


def render_bwd(sigma, z, rays, noise, raw_noise_std, seed, g_weights, g_depth, g_opacity, g_variance):
    n, S = z.shape
    d_sigma = torch.empty(n, S, device=z.device, dtype=torch.float32)
    d_rays = torch.zeros(n, RAY_COLS, device=z.device, dtype=torch.float32)
    L.check(L.load().loner_render_bwd(L.ptr(_f32(sigma)), L.ptr(_f32(z)), L.ptr(_f32(rays)), n, S, L.ptr(noise),
                                      float(raw_noise_std), seed, L.ptr(g_weights), L.ptr(g_depth), L.ptr(g_opacity),
                                      L.ptr(g_variance), L.ptr(d_sigma), L.ptr(d_rays), L.stream_ptr()),
            "loner_render_bwd")
    return d_sigma, d_rays


def render_loss(sigma, z, rays, depths, flags, counts, loss_cfg, noise=None, raw_noise_std=1.0, seed=0,
                loss_acc=None, want_outputs=True, d_rays=None):
    """Returns dict(loss_acc[4], d_sigma, d_rays, and optionally weights/depth/opacity/variance/eps)."""
    n, S = z.shape
    dev = z.device
    if loss_acc is None:
        loss_acc = torch.zeros(4, device=dev, dtype=torch.float32)
    out = dict(loss_acc=loss_acc)
    w = dpt = opa = var = eps = None
    if want_outputs:
        w = torch.empty(n, S, device=dev, dtype=torch.float32)
        dpt = torch.empty(n, device=dev, dtype=torch.float32)
        opa = torch.empty(n, device=dev, dtype=torch.float32)
        var = torch.empty(n, device=dev, dtype=torch.float32)
        eps = torch.empty(n, device=dev, dtype=torch.float32)
    d_sigma = torch.empty(n, S, device=dev, dtype=torch.float32)
    if d_rays is None:
        d_rays = torch.zeros(n, RAY_COLS, device=dev, dtype=torch.float32)
    cfg = L.host_floats(list(loss_cfg) + [0.0] * (9 - len(loss_cfg)))
    L.check(L.load().loner_render_loss(L.ptr(_f32(sigma)), L.ptr(_f32(z)), L.ptr(_f32(rays)), L.ptr(_f32(depths)),
                                       L.ptr(flags), n, S, L.ptr(noise), float(raw_noise_std), seed, L.ptr(counts),
                                       cfg, L.ptr(loss_acc), L.ptr(w), L.ptr(dpt), L.ptr(opa), L.ptr(var), L.ptr(eps),
                                       L.ptr(d_sigma), L.ptr(d_rays), L.stream_ptr()), "loner_render_loss")
    out.update(weights=w, depth=dpt, opacity=opa, variance=var, eps_dyn=eps, d_sigma=d_sigma, d_rays=d_rays)
    return out


def loss_finalize(loss_acc, counts, depthloss_lambda, los_lambda, S, out=None):
    """-> float[6] = loss, mean eps_dyn, depth loss, LOS loss, opacity loss, #valid rays (device, no sync)."""
    if out is None:
        out = torch.empty(6, device=loss_acc.device, dtype=torch.float32)
    L.check(L.load().loner_loss_finalize(L.ptr(loss_acc), L.ptr(counts), float(depthloss_lambda), float(los_lambda), int(S),
                                         L.ptr(out), L.stream_ptr()), "loner_loss_finalize")
    return out


def points_bwd(d_pos, z, d_rays):
    n, S = z.shape
    L.check(L.load().loner_points_bwd(L.ptr(_f32(d_pos)), L.ptr(_f32(z)), n, S, L.ptr(d_rays), L.stream_ptr()),
            "loner_points_bwd")
    return d_rays


def adam_step(params, grads, exp_avg, exp_avg_sq, step, lr, beta1=0.9, beta2=0.999, eps=1e-8, grad_unscale=1.0):
    L.check(L.load().loner_adam_step(L.ptr(params), L.ptr(grads), L.ptr(exp_avg), L.ptr(exp_avg_sq), params.numel(),
                                     int(step), float(lr), float(beta1), float(beta2), float(eps),
                                     float(grad_unscale), L.stream_ptr()), "loner_adam_step")


STATUS_ORIGIN_OUTSIDE, STATUS_BAD_POSE_GRAD, STATUS_BAD_POSE = 1, 2, 4      # LONER_STATUS_*


def pose_matrices(poses6, rows, out=None, shift=None, scale=1.0, status=None):
    """poses6 [n_keyframes,6] store, rows int32 [K] -> [K,12] = R row-major | t.  status: device int32 word that
    collects LONER_STATUS_* bits (origin outside the world cube given by shift / scale, non-finite pose)."""
    K = rows.shape[0]
    p12 = torch.empty(K, 12, device=poses6.device, dtype=torch.float32) if out is None else out
    sh = L.host_floats(shift) if shift is not None else None
    L.check(L.load().loner_pose_matrices(L.ptr(_f32(poses6)), L.ptr(rows), K, sh, float(scale), L.ptr(p12),
                                         L.ptr(status), L.stream_ptr()), "loner_pose_matrices")
    return p12


def pose_step(poses6, rows, free_rows, d_poses12, grad6, exp_avg=None, exp_avg_sq=None, steps=None, lr=0.0,
              beta1=0.9, beta2=0.999, eps=1e-8, apply=True, status=None):
    """Chain rule d_poses12 -> grad6 rows of the window; apply: Adam step on the rows with free_rows != 0."""
    L.check(L.load().loner_pose_step(L.ptr(_f32(poses6)), L.ptr(rows), L.ptr(free_rows), rows.shape[0],
                                     L.ptr(_f32(d_poses12)), L.ptr(grad6), L.ptr(exp_avg), L.ptr(exp_avg_sq),
                                     L.ptr(steps), float(lr), float(beta1), float(beta2), float(eps), int(bool(apply)),
                                     L.ptr(status), L.stream_ptr()), "loner_pose_step")
    return grad6


def ogm_grad(rays, z, depths, scale, V, d_grid=None, flags=None):
    n, S = z.shape
    if d_grid is None:
        d_grid = torch.zeros(V, V, V, device=z.device, dtype=torch.float32)
    L.check(L.load().loner_ogm_grad(L.ptr(_f32(rays)), L.ptr(_f32(z)), L.ptr(_f32(depths)), L.ptr(flags), n, S,
                                    float(scale), V, L.ptr(d_grid), L.stream_ptr()), "loner_ogm_grad")
    return d_grid


def sgd_step(x, g, lr):
    L.check(L.load().loner_sgd_step(L.ptr(x), L.ptr(g), x.numel(), float(lr), L.stream_ptr()), "loner_sgd_step")


def default_grad_scale(n_rays, S, los_lambda=1000.0):
    """Power-of-two loss scale that lifts d_sigma (~ los_lambda / (N*S)) into fp16's normal range."""
    return float(2.0 ** (math.ceil(math.log2(max(n_rays * S / max(los_lambda, 1e-6), 1.0))) + 7))
