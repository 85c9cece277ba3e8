"""Helpers for the GPU parity tests: decode the swizzled tile images the tensor-core kernels
write, and run the oracle's network layer by layer."""
import numpy as np
import torch

from oracle import loner_oracle as orc
from oracle import tcnn_standin


def decode_image(buf_u8: torch.Tensor, nb: int) -> torch.Tensor:
    """[nb*16384] uint8 (one tile image) -> [128, 64*nb] float32."""
    a = buf_u8.cpu().numpy().view(np.float16).reshape(nb, 128, 8, 8)
    out = np.empty((128, nb, 8, 8), dtype=np.float32)
    for r in range(128):
        for j in range(8):
            out[r, :, j, :] = a[:, r, j ^ (r & 7), :]
    return torch.from_numpy(out.reshape(128, nb * 64))


def oracle_layers(pos, params, spec):
    """Returns (enc fp32, [A_1..A_L] fp32 (pre fp16 rounding), sigma) for positions in [-1,1]."""
    x = (pos + 1) / 2
    enc = tcnn_standin.frequency_encode(x, spec.n_frequencies, pad_to=16)
    acts = []
    h = enc
    off = 0
    for li, (n_out, n_in) in enumerate(spec.shapes):
        W = params[off:off + n_out * n_in].view(n_out, n_in)
        off += n_out * n_in
        if spec.precision == "fp16":
            W = W.half().float()
            h = h.half().float()
        h = h @ W.t()
        if li < len(spec.shapes) - 1:
            h = torch.relu(h)
            acts.append(h)
    return enc, acts, h[:, 0]


def relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def norm_relerr(a, b):
    a = torch.as_tensor(a).detach().double().cpu().flatten()
    b = torch.as_tensor(b).detach().double().cpu().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def axis_angle_to_matrix(aa: torch.Tensor) -> torch.Tensor:
    """[...,3] -> [...,3,3] via the unit quaternion, as pytorch3d.transforms does for the reference
    (common/pose_utils.py:294): the autograd reference of loner_pose_matrices / loner_pose_step."""
    angles = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    small = angles.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angles), angles)
    s_over_a = torch.where(small, 0.5 - (angles * angles) / 48, torch.sin(half) / safe)
    q = torch.cat([torch.cos(half), aa * s_over_a], dim=-1)
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def poses6_to_poses12(poses6: torch.Tensor) -> torch.Tensor:
    """[K,6] = [t, axis-angle] -> [K,12] = R row-major | t  (common/pose_utils.py:288-302)."""
    R = axis_angle_to_matrix(poses6[:, 3:])
    return torch.cat([R.reshape(-1, 9), poses6[:, :3]], dim=1)
