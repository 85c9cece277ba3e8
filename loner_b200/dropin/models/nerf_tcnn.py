"""models.nerf_tcnn of the reference (/root/reference/src/models/nerf_tcnn.py): DecoupledNeRF with the
sigma head on the hand-written sm_100a kernels: Frequency + MLP on the tensor cores (loner_mlp_fwd /
loner_mlp_bwd) or the shipped HashGrid + 1 x 64 configuration and its 2-4 x 64 variants (loner_hash_fwd /
loner_hash_bwd).  Driven by
the same `nerf_config` keys (nerf_tcnn.py:29-33).  The intensity head exists only as frozen empty modules:
the reference never enables the camera (mapping/optimizer.py:433-434)."""
import torch
import torch.nn as nn

from loner_b200 import engine as _engine
from loner_b200 import ops


def device_grad_scale(g):
    """Power-of-two loss scale that brings the largest upstream gradient to ~2^4, as a DEVICE scalar: the
    backward stays free of host synchronisation (the kernels take grad_scale = 1 and the pre-scaled gradient)."""
    gmax = g.abs().max()
    return torch.clamp(16.0 / (gmax + 1e-30), 1.0, 2.0 ** 24).log2().floor().exp2()


class _SigmaFn(torch.autograd.Function):
    """sigma = MLP(encoding((pos + 1) / 2)) with a hand-rolled backward (params and positions)."""

    @staticmethod
    def forward(ctx, pos, params, module):
        P = pos.shape[0]
        need = params.requires_grad or pos.requires_grad
        posc = pos.detach().contiguous().float()
        sigma, acts = module.fwd(P, need, pos=posc)
        ctx.module, ctx.P, ctx.acts, ctx.pos = module, P, acts, posc
        ctx.want_dpos = pos.requires_grad
        return sigma.view(P, 1)

    @staticmethod
    def backward(ctx, g):
        m = ctx.module
        d_params = torch.zeros_like(m.params)
        g = g.contiguous().view(-1).float()
        c = device_grad_scale(g)
        d_pos = m.bwd(ctx.P, g * c, ctx.acts, 1.0, d_params, ctx.want_dpos, pos=ctx.pos)
        d_params.div_(c)
        if d_pos is not None:
            d_pos.div_(c)
        return d_pos, d_params, None


class SigmaNet(nn.Module):
    """tcnn.NetworkWithInputEncoding(3 -> 1) stand-in: one flat fp32 `params` (nerf_tcnn.py:35-38)."""

    def __init__(self, encoding_config, network_config, seed=1337):
        super().__init__()
        enc = dict(encoding_config)
        nw = dict(network_config)
        otype = enc.get("otype", "Frequency")
        if otype not in ("Frequency", "HashGrid"):
            raise NotImplementedError("sigma-head encoding otype=%r: Frequency and HashGrid are implemented" % otype)
        self.hash = otype == "HashGrid"
        if self.hash:
            self.net = ops.HashNet(int(enc["n_levels"]), int(enc["n_features_per_level"]), int(enc["log2_hashmap_size"]),
                                   int(enc["base_resolution"]), float(enc.get("per_level_scale", 2.0)),
                                   int(nw["n_neurons"]), int(nw["n_hidden_layers"]))
            flat = _engine.xavier_uniform_flat(self.net.layer_shapes(), seed)
            g = torch.Generator().manual_seed(seed + 1)      # tcnn: table ~ U(-1e-4, 1e-4), after the network matrices
            flat = torch.cat([flat, (torch.rand(2 * self.net.table_entries, generator=g) * 2 - 1) * 1e-4])
        else:
            self.net = ops.Net(int(enc.get("n_frequencies", 10)), int(nw["n_neurons"]), int(nw["n_hidden_layers"]))
            flat = _engine.xavier_uniform_flat(self.net.layer_shapes(), seed)
        self.params = nn.Parameter(flat)
        self.n_output_dims = 1
        self.dtype = torch.float16
        self._packed = None
        self._packed_version = -1

    def packed(self):
        if self._packed is None or self._packed_version != self.params._version or self._packed.device != self.params.device:
            pack = ops.hash_pack if self.hash else ops.mlp_pack
            self._packed = pack(self.net, self.params.detach())
            self._packed_version = self.params._version
        return self._packed

    def fwd(self, P, stash, pos=None, rays=None, z=None):
        """-> (sigma [P], activation stash or None).  The hash-grid path never stashes: its backward recomputes."""
        if self.hash:
            return ops.hash_fwd(self.net, self.packed(), P, pos=pos, rays=rays, z=z), None
        return ops.mlp_fwd(self.net, self.packed(), P, pos=pos, rays=rays, z=z, stash=stash)

    def bwd(self, P, d_sigma, acts, scale, d_params, want_dpos, pos=None, rays=None, z=None):
        if self.hash:
            return ops.hash_bwd(self.net, self.packed(), P, d_sigma, scale, d_params, pos=pos, rays=rays, z=z,
                                want_dpos=want_dpos)
        return ops.mlp_bwd(self.net, self.packed(), P, d_sigma, acts, scale, d_params, pos=pos, rays=rays, z=z,
                           want_dpos=want_dpos)

    def forward(self, pos01):
        """pos01 in [0,1] as tcnn receives it (nerf_tcnn.py:63); the kernel takes [-1,1]."""
        return _SigmaFn.apply(pos01 * 2 - 1, self.params, self)


class _Frozen(nn.Module):
    """Parameter-less placeholder for the intensity-head modules (`_pos_encoding`, ...)."""

    def __init__(self, n_output_dims=0):
        super().__init__()
        self.n_output_dims = n_output_dims
        self.dtype = torch.float16

    def forward(self, x):
        return torch.zeros(x.shape[0], self.n_output_dims, device=x.device)


class DecoupledNeRF(nn.Module):
    def __init__(self, cfg, num_colors=3):
        super().__init__()
        self._num_colors = num_colors
        self.cfg = cfg
        self._enable_view_dependence = cfg["enable_view_dependence"]
        self._model_sigma = SigmaNet(cfg["pos_encoding_sigma"], cfg["sigma_network"])
        self._pos_encoding = _Frozen()
        self._dir_encoding = _Frozen() if self._enable_view_dependence else None
        self._model_intensity = _Frozen(num_colors)
        self._max_float = torch.finfo(torch.float16).max
        self._min_float = torch.finfo(torch.float16).min
        self._warn_infinite = True

    def forward(self, pos, dir, sigma_only=False, detach_sigma=True):
        """pos [P,3] in [-1,1] -> sigma [P,1] (nerf_tcnn.py:59-78).  LiDAR-only: sigma_only=True."""
        if not sigma_only:
            raise NotImplementedError("intensity head: the reference never enables the camera "
                                      "(mapping/optimizer.py:433-434); only sigma_only=True is implemented")
        sigma = _SigmaFn.apply(pos, self._model_sigma.params, self._model_sigma)
        if not torch.isfinite(sigma).all():
            if self._warn_infinite:
                print("Warning: Clipping infinite outputs. Will not warn about this again (but it will happen again)")
                self._warn_infinite = False
            sigma = sigma.nan_to_num(posinf=self._max_float, neginf=self._min_float)
        return sigma
