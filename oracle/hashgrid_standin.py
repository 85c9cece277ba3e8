"""TEST INFRASTRUCTURE — CPU restatement of tiny-cuda-nn's multiresolution hash encoding.

PARITY UNPINNED (same status as oracle/tcnn_standin.py): `tinycudann` is an un-vendored, un-pinned
dependency of the reference (/root/reference/docker/container_dockerhub.Dockerfile:64) and the reference
has no test or golden vector at this boundary.  This file restates the PUBLISHED behaviour of
NVlabs/tiny-cuda-nn `GridEncoding` (grid type Hash, linear interpolation, Instant-NGP's spatial hash)
for the configuration the reference ships as its default sigma encoding
(/root/reference/cfg/nerf_config/default_nerf_hash.yaml: `pos_encoding_sigma`, used at
/root/reference/src/models/nerf_tcnn.py:35-38):

  level l = 0 .. n_levels-1
    scale_l      = 2^(l * log2(per_level_scale)) * base_resolution - 1          (per_level_scale defaults to 2)
    resolution_l = ceil(scale_l) + 1
    entries_l    = min(round_up(resolution_l^3, 8), 2^log2_hashmap_size)        ("hashmap size" of the level)
    pos          = fma(scale_l, x, 0.5),  cell = floor(pos),  frac = pos - cell   (x in [0,1]^3)
    for the 8 corners c:  weight = prod_d (c_d ? frac_d : 1 - frac_d),  p = cell + c
        index = p_x + p_y * res + p_z * res^2          when res^3 <= entries_l  (dense level)
              = (p_x * 1) ^ (p_y * 2654435761) ^ (p_z * 805459861)   (uint32 wrap-around) otherwise
        index %= entries_l
        out[l * F + f] += weight * table[offset_l + index][f]
  parameters: ONE flat table [sum_l entries_l, F]; tcnn initialises it uniformly in [-1e-4, 1e-4], keeps an
  fp32 master copy in the torch binding and computes with fp16 values (`precision="fp16"` here rounds
  the table and the encoded output to fp16, straight-through for autograd).

Only tests/, bench.py's cpu_baseline leg and __graft_entry__.smoke() may import this module."""
import math
from dataclasses import dataclass, field

import torch

PRIMES = (1, 2654435761, 805459861)


def _round_fp16_ste(x):
    return x + (x.half().float() - x).detach()


@dataclass
class HashGridSpec:
    n_levels: int = 16
    n_features_per_level: int = 2
    log2_hashmap_size: int = 18
    base_resolution: int = 16
    per_level_scale: float = 2.0
    scales: list = field(init=False)
    resolutions: list = field(init=False)
    entries: list = field(init=False)
    offsets: list = field(init=False)

    def __post_init__(self):
        self.scales, self.resolutions, self.entries, self.offsets = [], [], [], []
        off = 0
        log2_pls = torch.log2(torch.tensor(self.per_level_scale, dtype=torch.float32))
        for l in range(self.n_levels):
            # grid_scale(): exp2f(level * log2_per_level_scale) * base_resolution - 1.0f   (all fp32)
            scale = float(torch.exp2(torch.tensor(float(l), dtype=torch.float32) * log2_pls) * self.base_resolution - 1.0)
            res = int(math.ceil(scale)) + 1
            n = min(res ** 3, (2 ** 32 - 1) // 2)
            n = (n + 7) // 8 * 8
            n = min(n, 1 << self.log2_hashmap_size)
            self.scales.append(scale)
            self.resolutions.append(res)
            self.entries.append(n)
            self.offsets.append(off)
            off += n
        self.offsets.append(off)

    @classmethod
    def from_config(cls, cfg):
        return cls(n_levels=int(cfg["n_levels"]), n_features_per_level=int(cfg["n_features_per_level"]),
                   log2_hashmap_size=int(cfg["log2_hashmap_size"]), base_resolution=int(cfg["base_resolution"]),
                   per_level_scale=float(cfg.get("per_level_scale", 2.0)))

    @property
    def n_entries(self):
        return self.offsets[-1]

    @property
    def n_params(self):
        return self.n_entries * self.n_features_per_level

    @property
    def n_output_dims(self):
        return self.n_levels * self.n_features_per_level

    def level_is_dense(self, l):
        # grid_index(): the stride walk stops as soon as stride > hashmap size; the hash is used iff it did
        stride, hs = 1, self.entries[l]
        for _ in range(3):
            if stride > hs:
                break
            stride *= self.resolutions[l]
        return not (hs < stride)


def init_table(spec: HashGridSpec, seed: int, scale: float = 1e-4):
    """tcnn initialises the table uniformly in [-1e-4, 1e-4]; fixtures use a larger `scale` so that the
    encoded features (and their gradients) are not numerical noise."""
    g = torch.Generator().manual_seed(seed)
    return (torch.rand(spec.n_params, generator=g) * 2 - 1) * scale


def grid_index(spec: HashGridSpec, l: int, p: torch.Tensor) -> torch.Tensor:
    """p: int64 [P, 3] corner coordinates.  Returns int64 [P] entry index inside level l."""
    hs, res = spec.entries[l], spec.resolutions[l]
    if spec.level_is_dense(l):
        idx = p[:, 0] + p[:, 1] * res + p[:, 2] * res * res
    else:
        m = 0xFFFFFFFF
        idx = ((p[:, 0] * PRIMES[0]) & m) ^ ((p[:, 1] * PRIMES[1]) & m) ^ ((p[:, 2] * PRIMES[2]) & m)
    return idx % hs


def hashgrid_encode(x: torch.Tensor, table: torch.Tensor, spec: HashGridSpec, precision: str = "fp16") -> torch.Tensor:
    """x [P,3] in [0,1]; table flat fp32 [n_params].  Returns [P, n_levels*F] fp32, differentiable w.r.t.
    the table (scatter-add of the interpolation weights) and x (derivative of the linear interpolation)."""
    F = spec.n_features_per_level
    tbl = table.view(-1, F)
    if precision == "fp16":
        tbl = _round_fp16_ste(tbl)
    outs = []
    for l in range(spec.n_levels):
        scale = spec.scales[l]
        # pos = fmaf(scale, x, 0.5f): one rounding, reproduced through float64
        pos = (x.double() * scale + 0.5).float()
        cell = torch.floor(pos)
        frac = pos - cell.detach()          # d frac / d x = scale (floor has zero gradient)
        cell_i = cell.detach().long()
        acc = torch.zeros(x.shape[0], F, dtype=torch.float32)
        for c in range(8):
            bits = [(c >> d) & 1 for d in range(3)]
            w = torch.ones(x.shape[0], dtype=torch.float32)
            for d in range(3):
                w = w * (frac[:, d] if bits[d] else (1.0 - frac[:, d]))
            p = cell_i + torch.tensor(bits, dtype=torch.long)
            idx = grid_index(spec, l, p) + spec.offsets[l]
            acc = acc + w[:, None] * tbl[idx]
        outs.append(acc)
    out = torch.cat(outs, dim=1)
    if precision == "fp16":
        out = _round_fp16_ste(out)          # tcnn writes the encoding as __half
    return out
