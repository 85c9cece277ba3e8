"""GPU: edge cases and full-size properties (BASELINE.json C2 size) that do not need an oracle run:
empty inputs, ragged sample counts, sortedness/range of samples, finiteness and conservation laws
of the renderer, linearity of the backward kernels."""
import os
import sys

import pytest
import torch

from loner_b200 import engine as eng
from loner_b200 import ops, synth

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _rays(n, geom="canteen", seed=0):
    wc = synth.world_cube(geom)
    g = torch.Generator().manual_seed(seed)
    d = torch.randn(n, 3, generator=g)
    d = d / d.norm(dim=1, keepdim=True)
    r = torch.zeros(n, 13)
    r[:, 0:3] = (torch.rand(n, 3, generator=g) - 0.5) * 0.1
    r[:, 3:6] = d
    r[:, 6:9] = -d
    r[:, 11] = 1 / wc.scale_factor
    r[:, 12] = 50 / wc.scale_factor
    return r.to(DEV), wc


def test_empty_inputs_are_accepted():
    rays, wc = _rays(4)
    net = ops.Net(10, 128, 2)
    packed = ops.mlp_pack(net, eng.xavier_uniform_flat(net.layer_shapes(), 1).to(DEV))
    z = ops.sample_uniform(rays[:0], 64, 1.0)
    assert z.shape == (0, 64)
    z = ops.sample_ogm(rays[:0], torch.zeros(100, 100, 100, device=DEV), 64, 1.0)
    assert z.shape == (0, 64)
    sigma, _ = ops.mlp_fwd(net, packed, 0, rays=rays[:0], z=z)
    assert sigma.numel() == 0
    w, d, o, v = ops.render_fwd(sigma.view(0, 64), z, rays[:0])
    assert d.numel() == 0


@pytest.mark.parametrize("S", [96, 130, 256])
def test_ragged_sample_counts(S):
    """S that is not a power of two / not a multiple of 32: sampler pads its sort, the MLP tiles over the
    flat sample index, the renderer masks the tail of each warp pass."""
    n = 37
    rays, wc = _rays(n, seed=3)
    grid = synth.trained_occupancy_grid("canteen")[0, 0].to(DEV)
    if S % 2 == 0:
        z = ops.sample_ogm(rays, grid, S, 1.0, seed=5)
    else:
        z = ops.sample_uniform(rays, S, 1.0, seed=5)
    assert bool((z[:, 1:] >= z[:, :-1]).all())
    assert bool((z >= rays[:, 11:12] - 1e-6).all() and (z <= rays[:, 12:13] + 1e-6).all())
    net = ops.Net(10, 128, 2)
    packed = ops.mlp_pack(net, (eng.xavier_uniform_flat(net.layer_shapes(), 2) * 3).to(DEV))
    sigma, _ = ops.mlp_fwd(net, packed, n * S, rays=rays, z=z)
    pos = (rays[:, None, 0:3] + rays[:, None, 3:6] * z[:, :, None]).reshape(-1, 3).contiguous()
    sigma2, _ = ops.mlp_fwd(net, packed, n * S, pos=pos)
    assert torch.allclose(sigma, sigma2, rtol=0, atol=1e-5)          # (rays, z) mode == explicit positions mode
    w, d, o, v = ops.render_fwd(sigma.view(n, S), z, rays, raw_noise_std=0.0)
    assert torch.isfinite(d).all() and bool((o >= 0).all() and (o <= 1 + 1e-5).all())
    assert torch.allclose(w.sum(1), o, atol=1e-5)
    expect = (w * z).sum(1) + (1 - o) * rays[:, 12]
    assert torch.allclose(d, expect, rtol=1e-5, atol=1e-6)           # depth = sum w z + (1 - A) far


def test_full_size_step_properties():
    """BASELINE C2 size (8192 rays x 512 samples, 4x256): finite loss that decreases, sorted samples,
    weights consistent with opacity, gradient linear in the upstream gradient."""
    wc = synth.world_cube("canteen")
    cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_samples=512)
    e = eng.MappingEngine(cfg)
    scans, poses = synth.make_window("canteen", 1, seed=0)
    e.add_keyframe(scans[0].ray_directions, scans[0].distances, synth.axis_angle_from_yaw_pose(poses[0]))
    e.grid.copy_(synth.trained_occupancy_grid("canteen")[0, 0])
    e.new_phase(False)
    losses = []
    for it in range(12):
        losses.append(float(e.step([0], 8192, want_outputs=(it == 0))))
        if it == 0:
            o = e.last["outs"][0]
            z, w = o["z_vals"], o["weights"]
            assert bool((z[:, 1:] >= z[:, :-1]).all())
            assert torch.allclose(w.sum(1), o["opacity"], atol=2e-5)
            assert torch.isfinite(o["depth"]).all() and torch.isfinite(e.d_params).all()
            assert int(e.last["counters"][0]) == 8192
    assert all(l == l and l < 1e6 for l in losses)
    assert losses[-1] < losses[0]                                     # Adam on the fused gradients reduces the loss
    # linearity of the MLP backward in d_sigma (size-independent property of the wgrad/dgrad kernels)
    net, P = e.net, 4096 * 8
    pos = (torch.rand(P, 3, device=DEV) * 1.6 - 0.8).contiguous()
    sigma, acts = ops.mlp_fwd(net, e.packed, P, pos=pos, stash=True)
    g1 = torch.randn(P, device=DEV) * 1e-4
    g2 = torch.randn(P, device=DEV) * 1e-4
    outs = []
    for g in (g1, g2, g1 + g2):
        dp = torch.zeros(net.param_count, device=DEV)
        ops.mlp_bwd(net, e.packed, P, g, acts, 2.0 ** 12, dp, pos=pos)
        outs.append(dp)
    err = float((outs[0] + outs[1] - outs[2]).norm() / outs[2].norm())
    assert err < 2e-3, err                                            # fp16 rounding of the scaled gradients only


def test_ray_selection_strategies():
    """FIXED = arange(n), MASK = random picks among the scan mask (optimizer.py:286-296)."""
    wc = synth.world_cube("canteen")
    scans, poses = synth.make_window("canteen", 1, seed=0, n_beams=16, n_azimuth=256)
    M = scans[0].distances.shape[0]
    mask = torch.zeros(M, dtype=torch.bool)
    mask[100:164] = True
    for strategy in ("FIXED", "MASK", "RANDOM"):
        cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_neurons=128,
                               n_hidden_layers=2, n_samples=64, rays_selection=strategy)
        e = eng.MappingEngine(cfg)
        e.add_keyframe(scans[0].ray_directions, scans[0].distances, synth.axis_angle_from_yaw_pose(poses[0]), mask=mask)
        _, rp = e._pick_rays([0], 256)
        if strategy == "FIXED":
            assert torch.equal(rp.cpu(), torch.arange(256))
        elif strategy == "MASK":
            assert int(rp.min()) >= 100 and int(rp.max()) < 164
        else:
            assert int(rp.min()) >= 0 and int(rp.max()) < M and rp.unique().numel() > 200
        loss = e.step([0], 256)
        assert torch.isfinite(loss)
