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
