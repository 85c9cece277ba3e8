"""CPU check of the identity behind mlp.cu's "fold" (wgrad without an A_L stash): the sigma head has no biases
(models/nerf_tcnn.py:35-38: tcnn FullyFusedMLP), so with mask_L = relu'(Z_L) and

    G[n,k] = sum_s d_sigma[s] * mask_L[s,n] * A_{L-1}[s,k]

the last hidden layer's gradient is dW_{L-1}[n,k] = w_out[n] * G[n,k] and the output layer's gradient is
dW_out[n] = sum_k W_{L-1}[n,k] * G[n,k] - no activation of the last layer is needed.  Checked here against the oracle's
autograd (the same reference the GPU parity tests use), in fp32 (identity to rounding) and with the fp16 roundings of the
kernels (operand = fp16(d_sigma * gscale * c), c = power of two >= max|w_out|)."""
import math

import pytest
import torch

from gpu_util import norm_relerr, oracle_layers
from oracle import loner_oracle as orc
from oracle import tcnn_standin


def _case(W, L, P, precision):
    spec = orc.NetSpec(n_frequencies=10, n_neurons=W, n_hidden_layers=L, precision=precision)
    params = tcnn_standin.xavier_uniform_flat(spec.shapes, 1337) * 1.5
    g = torch.Generator().manual_seed(4)
    pos = torch.rand(P, 3, generator=g) * 1.8 - 0.9
    d_sigma = torch.randn(P, generator=g) * 1e-4
    p_ref = params.clone().requires_grad_(True)
    (orc.sigma_net(pos, p_ref, spec) * d_sigma).sum().backward()
    return spec, params, pos, d_sigma, p_ref.grad


def _split(flat, shapes):
    out, off = [], 0
    for no, ni in shapes:
        out.append(flat[off:off + no * ni].view(no, ni))
        off += no * ni
    return out


@pytest.mark.parametrize("W,L", [(64, 2), (128, 3), (256, 4)])
def test_fold_identity_fp32(W, L):
    spec, params, pos, d_sigma, grad = _case(W, L, 700, "fp32")
    _, acts, _ = oracle_layers(pos, params, spec)
    Ws, Gs = _split(params, spec.shapes), _split(grad, spec.shapes)
    a_prev, mask = acts[L - 2], (acts[L - 1] > 0).float()
    G = (d_sigma[:, None] * mask).t() @ a_prev                    # [n,k]
    w_out = Ws[L][0]
    assert norm_relerr(w_out[:, None] * G, Gs[L - 1]) < 1e-5      # dW_{L-1}
    assert norm_relerr((Ws[L - 1] * G).sum(1), Gs[L][0]) < 1e-5   # dW_out (row 0 of the padded [16,W] matrix)


@pytest.mark.parametrize("W,L", [(128, 2), (256, 4)])
def test_fold_with_the_kernels_fp16_roundings(W, L):
    spec, params, pos, d_sigma, grad = _case(W, L, 3000, "fp16")
    _, acts, _ = oracle_layers(pos, params, spec)
    Ws, Gs = _split(params, spec.shapes), _split(grad, spec.shapes)
    a_prev = acts[L - 2].half().float()                           # the stashed A_{L-1} image
    mask = (acts[L - 1].half() > 0).float()                       # "active" = non-zero fp16 output (relu_mask_word)
    w_out16, w_last16 = Ws[L][0].half().float(), Ws[L - 1].half().float()
    gscale = 2.0 ** 12
    c = 2.0 ** math.ceil(math.log2(float(w_out16.abs().max())))   # pow2_ceil
    operand = (d_sigma * gscale * c).half().float()[:, None] * mask
    G = operand.t() @ a_prev
    dW_last = G * (w_out16 / c)[:, None] / gscale
    dW_out = (w_last16 * G).sum(1) / (gscale * c)
    # same bars as tests/test_gpu_kernels.py::test_mlp_backward_matches_autograd
    assert norm_relerr(dW_last, Gs[L - 1]) < 5e-4 * max(4, 2 * L)
    assert norm_relerr(dW_out, Gs[L][0]) < 5e-4 * max(4, 2 * L)
