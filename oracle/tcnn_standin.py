"""TEST INFRASTRUCTURE — CPU stand-in for the `tinycudann` torch bindings.

PARITY UNPINNED for this file: `tinycudann` is an un-vendored, un-pinned third-party
dependency of the reference (pip git HEAD, /root/reference/docker/container_dockerhub.Dockerfile:64)
and the reference holds no test or golden vector at this boundary.  What is restated here is
the *published* behaviour of NVlabs/tiny-cuda-nn that the reference relies on at its call
sites (/root/reference/src/models/nerf_tcnn.py:35-55, :59-95):

  * `Encoding(n_input_dims, cfg)` with `otype: Frequency`:
        out[d*2F + 2f + 0] = sin(pi * 2^f * x_d),  out[d*2F + 2f + 1] = cos(pi * 2^f * x_d)
    (tcnn writes the second one as sin(. + pi/2)); encoded width padded up to a multiple of
    16 with the constant 1.0.
  * `Network(n_input_dims, n_output_dims, cfg)`: bias-free MLP, ReLU hidden activations,
    `output_activation: None`, weight matrices `[out, in]` row-major concatenated into ONE flat
    fp32 `params` tensor, output width padded to 16 and sliced back; xavier-uniform init.
  * `NetworkWithInputEncoding` = the two chained, one flat `params`.
  * mixed precision: tcnn keeps fp32 master params and computes with fp16 weights and fp16
    activations.  `precision="fp16"` emulates exactly that rounding (params and every layer
    input rounded to fp16, fp32 accumulation, final layer output left in fp32);
    `precision="fp32"` is plain fp32 everywhere.

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import this module.
"""
import math

import torch
import torch.nn as nn

try:                                   # imported as oracle.tcnn_standin or, by the reference harness, top-level
    from . import hashgrid_standin
except ImportError:                    # pragma: no cover
    import hashgrid_standin

PRECISION = {"mode": "fp32"}   # module-level switch used by the reference harness


def _round_fp16_ste(x: torch.Tensor) -> torch.Tensor:
    """Round to fp16 and back, straight-through for autograd."""
    return x + (x.half().float() - x).detach()


def frequency_encode(x: torch.Tensor, n_frequencies: int, pad_to: int = 16) -> torch.Tensor:
    """x: [P, D] in [0, 1].  Returns [P, pad16(D*2F)] fp32; differentiable w.r.t. x.

    Computed through float64 so that the result is the correctly rounded fp32 value of
    sin/cos(pi * 2^f * x) (the argument 2^f * x is exact in fp32)."""
    P, D = x.shape
    F = n_frequencies
    xd = x.double()
    scales = (2.0 ** torch.arange(F, dtype=torch.float64, device=x.device)) * math.pi  # [F]
    arg = xd[:, :, None] * scales[None, None, :]                   # [P, D, F]
    enc = torch.stack([torch.sin(arg), torch.cos(arg)], dim=-1)    # [P, D, F, 2]
    enc = enc.reshape(P, D * 2 * F).float()
    width = D * 2 * F
    padded = (width + pad_to - 1) // pad_to * pad_to
    if padded > width:
        enc = torch.cat([enc, torch.ones(P, padded - width, dtype=enc.dtype, device=x.device)], dim=1)
    return enc


def _pad16(n):
    return (n + 15) // 16 * 16


class Encoding(nn.Module):
    """tcnn.Encoding stand-in.  Frequency is computed; every other otype (HashGrid,
    SphericalHarmonics: only used by the frozen intensity head, whose output the LiDAR-only
    path discards, nerf_tcnn.py:64) returns zeros of the right width."""

    def __init__(self, n_input_dims, encoding_config, dtype=None):
        super().__init__()
        self.n_input_dims = n_input_dims
        self.cfg = dict(encoding_config)
        ot = self.cfg.get("otype", "Frequency")
        if ot == "Frequency":
            self.n_frequencies = int(self.cfg.get("n_frequencies", 10))
            self.n_output_dims = n_input_dims * 2 * self.n_frequencies
        elif ot == "HashGrid":
            self.hash_spec = hashgrid_standin.HashGridSpec.from_config(self.cfg)
            self.n_output_dims = self.hash_spec.n_output_dims
        elif ot == "SphericalHarmonics":
            self.n_output_dims = int(self.cfg["degree"]) ** 2
        else:
            self.n_output_dims = n_input_dims
        self.otype = ot
        self.params = nn.Parameter(torch.zeros(0))
        self.dtype = torch.float32

    def forward(self, x):
        if self.otype == "Frequency":
            return frequency_encode(x.float(), self.n_frequencies, pad_to=1)
        # stand-alone HashGrid / SphericalHarmonics encodings only feed the frozen intensity head, whose
        # output the LiDAR-only path discards (nerf_tcnn.py:64): zeros of the right width
        return torch.zeros(x.shape[0], self.n_output_dims, dtype=torch.float32, device=x.device)


def xavier_uniform_flat(layer_shapes, seed):
    """Flat fp32 params: concatenation of [out, in] row-major matrices, xavier-uniform."""
    g = torch.Generator().manual_seed(seed)
    chunks = []
    for (n_out, n_in) in layer_shapes:
        bound = math.sqrt(6.0 / (n_in + n_out))
        chunks.append(((torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound).reshape(-1))
    return torch.cat(chunks)


def mlp_layer_shapes(n_in_padded, n_neurons, n_hidden_layers, n_out_padded):
    shapes = [(n_neurons, n_in_padded)]
    for _ in range(n_hidden_layers - 1):
        shapes.append((n_neurons, n_neurons))
    shapes.append((n_out_padded, n_neurons))
    return shapes


def mlp_forward(x, params, shapes, precision):
    """x: [P, n_in_padded] fp32.  Returns [P, n_out_padded] fp32."""
    off = 0
    h = x
    for li, (n_out, n_in) in enumerate(shapes):
        W = params[off:off + n_out * n_in].view(n_out, n_in)
        off += n_out * n_in
        if precision == "fp16":
            W = _round_fp16_ste(W)
            h = _round_fp16_ste(h)
        h = h @ W.t()
        if li < len(shapes) - 1:
            h = torch.relu(h)
    return h


class Network(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, network_config, seed=1337):
        super().__init__()
        cfg = dict(network_config)
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.n_neurons = int(cfg["n_neurons"])
        self.n_hidden_layers = int(cfg["n_hidden_layers"])
        self.in_padded = _pad16(n_input_dims)
        self.out_padded = _pad16(n_output_dims)
        self.shapes = mlp_layer_shapes(self.in_padded, self.n_neurons, self.n_hidden_layers, self.out_padded)
        self.params = nn.Parameter(xavier_uniform_flat(self.shapes, seed))
        self.dtype = torch.float16  # what the reference reads at nerf_tcnn.py:54-55

    def forward(self, x):
        x = x.float()
        if x.shape[1] < self.in_padded:
            x = torch.cat([x, torch.ones(x.shape[0], self.in_padded - x.shape[1], dtype=x.dtype)], dim=1)
        out = mlp_forward(x, self.params, self.shapes, PRECISION["mode"])
        return out[:, :self.n_output_dims]


class NetworkWithInputEncoding(nn.Module):
    def __init__(self, n_input_dims, n_output_dims, encoding_config, network_config, seed=1337):
        super().__init__()
        self.encoding = Encoding(n_input_dims, encoding_config)
        if self.encoding.otype not in ("Frequency", "HashGrid"):
            raise NotImplementedError("tcnn stand-in: sigma-head encoding must be Frequency or HashGrid")
        cfg = dict(network_config)
        self.n_input_dims = n_input_dims
        self.n_output_dims = n_output_dims
        self.n_neurons = int(cfg["n_neurons"])
        self.n_hidden_layers = int(cfg["n_hidden_layers"])
        self.in_padded = _pad16(self.encoding.n_output_dims)
        self.out_padded = _pad16(n_output_dims)
        self.shapes = mlp_layer_shapes(self.in_padded, self.n_neurons, self.n_hidden_layers, self.out_padded)
        flat = xavier_uniform_flat(self.shapes, seed)
        self.n_network_params = flat.numel()
        if self.encoding.otype == "HashGrid":       # tcnn order: network parameters, then the encoding's table
            flat = torch.cat([flat, hashgrid_standin.init_table(self.encoding.hash_spec, seed + 1)])
        self.params = nn.Parameter(flat)
        self.dtype = torch.float16

    def forward(self, x):
        if self.encoding.otype == "HashGrid":
            enc = hashgrid_standin.hashgrid_encode(x.float(), self.params[self.n_network_params:],
                                                   self.encoding.hash_spec, PRECISION["mode"])
            if enc.shape[1] < self.in_padded:
                enc = torch.cat([enc, torch.ones(enc.shape[0], self.in_padded - enc.shape[1])], dim=1)
        else:
            enc = frequency_encode(x.float(), self.encoding.n_frequencies, pad_to=16)
        out = mlp_forward(enc, self.params[:self.n_network_params], self.shapes, PRECISION["mode"])
        return out[:, :self.n_output_dims]
