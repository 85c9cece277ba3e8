"""GPU parity evidence added in round 2 (VERDICT r1 items 1, 2, 6, 8 and the ADVICE findings):
test-mode render + depth-L1 metric, full-size (BASELINE C2) and C5-shaped comparisons with the oracle,
sky rays, the one-kernel ray pick, filtered rays in the occupancy update, LR schedule / los_lambda decay,
width-64 networks, checkpoint surface, and the 2-rank NCCL gradient check."""
import importlib
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

from golden_util import Case, TestModeCase
from gpu_util import norm_relerr, relerr
from loner_b200 import engine as eng
from loner_b200 import ops, synth
from oracle import loner_oracle as orc
from oracle import tcnn_standin

pytestmark = pytest.mark.gpu
DEV = "cuda"
HERE = os.path.dirname(os.path.abspath(__file__))


def _engine(c, **over):
    kw = dict(scale=c.scale, shift=tuple(c.shift.tolist()), ray_range=c.ray_range, n_frequencies=10,
              n_neurons=c.W, n_hidden_layers=c.L, n_samples=c.S)
    kw.update(over)
    e = eng.MappingEngine(eng.EngineConfig(**kw), params=c.params)
    e.grid.copy_(c.grid[0, 0])
    return e


def _same_draws(z_mine, z_ref, tol=1e-5):
    """Rays whose sorted samples all agree (no importance draw crossed the discontinuity of sample_pdf's guard)."""
    return (z_mine - z_ref).abs().max(dim=1)[0] < tol


# ------------------------------------------------------------------------------------------ f3: test mode
def test_render_test_mode_and_depth_l1_match_reference_fixture():
    """MappingEngine.render = Model.forward(testing=True) (models/model_tcnn.py:73-75) at N_samples_test = 2048,
    perturb = 0, with the draws that stay active in the reference's test mode injected; then the depth-L1 metric
    of analysis/compute_l1_depth.py:42-64.  Compared with the fixture minted from the reference AND the oracle."""
    c = TestModeCase("testmode_2x128")
    e = _engine(c, n_samples=128, chunk_rays=1024)          # render chunks: 1024*128/2048 = 64 rays -> 4 chunks
    rays_o, depths, res_o, l1_o = c.run_oracle()
    out = e.render(rays_o.to(DEV), n_samples=c.S, injected=dict(u2=c.u2, noise=c.noise))
    torch.cuda.synchronize()
    g = c.g
    depth_m = out["depth_fine"].cpu() * c.scale
    l1 = orc.depth_l1_metric(out["depth_fine"].cpu(), depths, c.scale, c.ray_range)
    # sample_pdf's `denom < eps -> 1` guard makes the inverse CDF discontinuous: a draw within
    # ~1e-7 of a cdf edge of an (almost) empty bin lands one coarse bin further when the row sum differs in its last
    # bit (with 1024 bins x 1024 draws that is a few per cent of the rays).  Per-ray outputs are compared on the rays
    # where no draw moved; the depth-L1 METRIC - the thing BASELINE.json asks for - over ALL rays.
    same = _same_draws(out["samples_fine"].cpu(), res_o["samples_fine"])
    errs = dict(rays_with_moved_draw=float((~same).float().mean()),
                depth_vs_fixture=relerr(depth_m[same], torch.from_numpy(g["depth_m"])[same]),
                depth_vs_oracle=relerr(out["depth_fine"].cpu()[same], res_o["depth_fine"][same]),
                opacity=relerr(out["opacity_fine"].cpu()[same], torch.from_numpy(g["opacity"])[same]),
                variance=relerr(out["variance"].cpu()[same], torch.from_numpy(g["variance"])[same]),
                depth_all_rays=relerr(depth_m, g["depth_m"]),
                l1_vs_fixture=abs(float(l1) - float(g["l1"])) / float(g["l1"]),
                l1_vs_oracle=abs(float(l1) - float(l1_o)) / float(l1_o))
    print("test-mode render: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()) + f"  L1 = {float(l1):.5f} m")
    assert errs["rays_with_moved_draw"] < 0.15
    for k in ("depth_vs_fixture", "depth_vs_oracle", "opacity", "l1_vs_fixture", "l1_vs_oracle"):
        assert errs[k] < 1e-4, k              # north_star: depth L1 matching the reference within 1e-4
    assert errs["variance"] < 2e-4 and errs["depth_all_rays"] < 2e-3


class _Cfg(dict):
    def __getattr__(self, k):
        v = self[k]
        return _Cfg(v) if isinstance(v, dict) else v


def test_dropin_model_forward_testing_mode():
    """The drop-in `Model.forward(testing=True)` (N_samples_test, perturb = 0) through the reference-facing API, the
    reference's replayed draws attached to the sampler, against the reference-minted fixture."""
    from loner_b200 import dropin
    path = dropin.install()
    try:
        mt = importlib.import_module("models.model_tcnn")
        rs = importlib.import_module("models.ray_sampling")
        c = TestModeCase("testmode_2x128")
        cfg = _Cfg(model_type="nerf_decoupled", num_colors=3, ray_range=list(c.ray_range),
                   nerf_config=dict(enable_view_dependence=True, pos_encoding_sigma=dict(otype="Frequency", n_frequencies=10),
                                    sigma_network=dict(otype="FullyFusedMLP", n_neurons=c.W, n_hidden_layers=c.L)),
                   render=dict(N_samples_train=128, N_samples_test=c.S, retraw=True, perturb=1.0, white_bkgd=False,
                               raw_noise_std=1.0, chunk=16384, netchunk=0))
        model = mt.Model(cfg).cuda()
        with torch.no_grad():
            model.nerf_model._model_sigma.params.copy_(c.params)
        sampler = rs.OccGridRaySampler()
        sampler.update_occ_grid(c.grid.cuda())
        sampler.injected = dict(u2=c.u2, noise=c.noise)
        rays = torch.from_numpy(c.g["rays"]).cuda()
        with torch.no_grad():
            res = model(rays, sampler, c.scale, testing=True, return_variance=True, camera=False)
        assert res["samples_fine"].shape == (c.n, c.S)
        _, _, res_o, _ = c.run_oracle()
        same = _same_draws(res["samples_fine"].cpu(), res_o["samples_fine"])       # see the test above
        depth_m = res["depth_fine"].cpu() * c.scale
        l1 = orc.depth_l1_metric(res["depth_fine"].cpu(), c.distances / c.scale, c.scale, c.ray_range)
        e1 = relerr(depth_m[same], torch.from_numpy(c.g["depth_m"])[same])
        e2 = abs(float(l1) - float(c.g["l1"])) / float(c.g["l1"])
        print(f"drop-in Model.forward(testing=True): depth rel {e1:.2e} depth-L1 rel {e2:.2e} "
              f"(rays with a moved draw: {float((~same).float().mean()):.3f})")
        assert e1 < 1e-4 and e2 < 1e-4 and float((~same).float().mean()) < 0.15
    finally:
        sys.path.remove(path)
        for m in list(sys.modules):
            if m == "models" or m.startswith("models."):
                del sys.modules[m]


# ------------------------------------------------------------------------------------------ full sizes
def _oracle_forward_chunked(scans, poses6, idx, params, spec, grid, S, scale, shift, ray_range, u1, u2, noise, chunk=1024):
    """The oracle's forward + loss at sizes where one call would need tens of GB: rays are rendered in chunks
    (no autograd) and the loss is evaluated once on the concatenated results, as the reference's chunk loop does."""
    rays_l, dep_l = [], []
    for sc, p6, ix in zip(scans, poses6, idx):
        r, d, keep = orc.build_lidar_rays(sc.ray_directions, sc.distances, ix, orc.pose6_to_matrix(p6), ray_range, scale, shift)
        assert bool(keep.all())
        rays_l.append(r)
        dep_l.append(d)
    rays, depths = torch.cat(rays_l).float(), torch.cat(dep_l).float()
    parts = []
    with torch.no_grad():
        for c0 in range(0, rays.shape[0], chunk):
            sl = slice(c0, c0 + chunk)
            z = orc.ogm_samples(rays[sl], grid, S, 1.0, u1[sl], u2[sl])
            r = orc.render_rays(rays[sl], z, params, spec, noise[sl])
            parts.append({k: r[k] for k in ("depth_fine", "weights_fine", "opacity_fine", "variance", "samples_fine")})
        res = {k: torch.cat([p[k] for p in parts]) for k in parts[0]}
        out = orc.compute_loss(rays, depths, res, scale, orc.LossCfg())
    return rays, depths, res, out


def test_full_size_c2_step_vs_oracle():
    """BASELINE config 2 at FULL size (8192 rays x 512 samples, 4 x 256, canteen geometry, trained occupancy grid):
    depth / opacity / loss / mean margin of one fused iteration against the oracle on identical picks and draws."""
    wc = synth.world_cube("canteen")
    N, S = 8192, 512
    scans, poses = synth.make_window("canteen", 1, seed=0)
    p6 = synth.axis_angle_from_yaw_pose(poses[0])
    spec = orc.NetSpec(10, 256, 4, "fp16")
    params = tcnn_standin.xavier_uniform_flat(spec.shapes, 1337)
    grid = synth.trained_occupancy_grid("canteen")
    g = torch.Generator().manual_seed(5)
    M = scans[0].distances.shape[0]
    # picks restricted to returns whose ray survives the validity filter (all of them here) - checked by the assert below
    idx = torch.randint(0, M, (N,), generator=g)
    u1, u2 = torch.rand(N, S // 2, generator=g), torch.rand(N, S // 2, generator=g)
    noise = torch.randn(N, S, generator=g)
    cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_samples=S)
    e = eng.MappingEngine(cfg, params=params)
    e.add_keyframe(scans[0].ray_directions, scans[0].distances, p6)
    e.grid.copy_(grid[0, 0])
    e.new_phase(False)
    loss = e.step([0], N, injected=dict(ray_point=idx, u1=u1, u2=u2, noise=noise), want_outputs=True)
    torch.cuda.synchronize()
    o = e.last["outs"][0]
    rays, depths, res, out = _oracle_forward_chunked(scans, [p6], [idx], params, spec, grid, S, wc.scale_factor,
                                                     torch.tensor(wc.shift), (1.0, 50.0), u1, u2, noise)
    assert int(e.last["counters"][0]) == N
    # sample_pdf's `denom < eps -> 1` guard is discontinuous: a one-ulp cdf difference can move an
    # isolated importance draw by one coarse bin.  Rays where that happened are counted (and must be rare); per-ray
    # outputs are compared on the other rays, the loss and the mean margin over ALL rays.
    same = _same_draws(o["z_vals"].cpu(), res["samples_fine"])
    errs = dict(rays_with_moved_draw=float((~same).float().mean()),
                depth=relerr(o["depth"].cpu()[same], res["depth_fine"][same]),
                opacity=relerr(o["opacity"].cpu()[same], res["opacity_fine"][same]),
                depth_all_rays=relerr(o["depth"], res["depth_fine"]),
                loss=abs(float(loss) - float(out["loss"])) / float(out["loss"]),
                depth_eps=abs(float(e.last["depth_eps"]) - out["depth_eps_mean"]) / out["depth_eps_mean"])
    print("C2 full size: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert errs["rays_with_moved_draw"] < 5e-3
    assert errs["depth"] < 1e-4 and errs["opacity"] < 1e-4 and errs["loss"] < 1e-4 and errs["depth_eps"] < 1e-4


def test_c5_shaped_step_vs_oracle_autograd():
    """BASELINE config 5's shape at an oracle-friendly size: 16 keyframes, joint pose + map optimisation, the ray batch
    split into 4 chunks; loss, depth, MLP gradient and pose gradients against the oracle's autograd."""
    K, n, S, W, L = 16, 48, 128, 128, 2
    wc = synth.world_cube("canteen")
    scans, poses = synth.make_window("canteen", K, seed=3, n_beams=16, n_azimuth=256)
    poses6 = [synth.axis_angle_from_yaw_pose(poses[k]) for k in range(K)]
    spec = orc.NetSpec(10, W, L, "fp16")
    params = tcnn_standin.xavier_uniform_flat(spec.shapes, 7)
    grid = synth.trained_occupancy_grid("canteen")
    g = torch.Generator().manual_seed(9)
    M = scans[0].distances.shape[0]
    idx = [torch.randint(0, M, (n,), generator=g) for _ in range(K)]
    N = K * n
    u1, u2 = torch.rand(N, S // 2, generator=g), torch.rand(N, S // 2, generator=g)
    noise = torch.randn(N, S, generator=g)
    cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_neurons=W, n_hidden_layers=L,
                           n_samples=S, chunk_rays=N // 4)
    e = eng.MappingEngine(cfg, params=params)
    for k in range(K):
        e.add_keyframe(scans[k].ray_directions, scans[k].distances, poses6[k])
    e.grid.copy_(grid[0, 0])
    e.new_phase(optimize_poses=True)
    ray_point = torch.cat([idx[k] + e.kf_offsets[k] for k in range(K)])
    loss = e.step(list(range(K)), n, optimize_poses=True, injected=dict(ray_point=ray_point, u1=u1, u2=u2, noise=noise),
                  want_outputs=True)
    torch.cuda.synchronize()
    p = params.clone().requires_grad_(True)
    p6 = [q.clone().requires_grad_(k > 0) for k, q in enumerate(poses6)]
    rays, depths, res, out = orc.mapping_iteration(scans, p6, idx, p, spec, grid, S, wc.scale_factor, torch.tensor(wc.shift),
                                                   (1.0, 50.0), 1.0, u1, u2, noise, orc.LossCfg())
    assert rays.shape[0] == N
    out["loss"].backward()
    depth = torch.cat([o["depth"] for o in e.last["outs"]])
    mine = torch.stack([q.grad.cpu() if q.grad is not None else torch.zeros(6) for q in e.poses6])
    ref = torch.stack([q.grad if q.grad is not None else torch.zeros(6) for q in p6])
    errs = dict(depth=relerr(depth, res["depth_fine"]), loss=abs(float(loss) - float(out["loss"])) / float(out["loss"]),
                d_params=norm_relerr(e.d_params, p.grad), pose=norm_relerr(mine, ref))
    print("C5 shape (16 KF, 4 chunks, poses on): " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert len(e.last["outs"]) == 4
    assert errs["depth"] < 1e-4 and errs["loss"] < 1e-4
    assert errs["d_params"] < 2e-3          # fp16 dZ with a power-of-two loss scale vs the oracle's fp32 backward
    assert errs["pose"] < 5e-2              # fp16 input-gradient chain summed over N*S samples with heavy cancellation
    assert torch.equal(mine[0], torch.zeros(6))     # keyframe 0 is anchored


# ------------------------------------------------------------------------------------------ sky rays / ray pick
def test_sky_rays_step_matches_reference_fixture():
    """num_samples.sky picks per keyframe among LidarScan.sky_rays, at distance ray_range[1] + 1, from the DETACHED
    pose (optimizer.py:299-305, keyframe.py:87-99, sensors.py:162-167): fixture minted from the reference."""
    c = Case("kf2_2x128_sky")
    e = _engine(c, n_sky=c.n_sky)
    for k in range(c.K):
        e.add_keyframe(c.scans[k].ray_directions, c.scans[k].distances, c.poses6[k], sky_rays=c.sky_dirs[k])
    e.new_phase(optimize_poses=True)
    rk, rp = [], []
    for k in range(c.K):
        rk += [torch.full((c.n,), k, dtype=torch.int32), torch.full((c.n_sky,), k | ops.KF_DETACHED, dtype=torch.int32)]
        rp += [c.idx[k] + e.kf_offsets[k], c.sky_idx[k] + e.kf_sky_offsets[k]]
    inj = dict(ray_kf=torch.cat(rk), ray_point=torch.cat(rp), u1=c.u1, u2=c.u2, noise=c.noise)
    loss = e.step(list(range(c.K)), c.n, optimize_poses=True, injected=inj, want_outputs=True)
    torch.cuda.synchronize()
    g, o = c.g, e.last["outs"][0]
    assert relerr(e.last["rays"], g["rays"]) < 2e-6 and relerr(e.last["depths"], g["depths"]) < 1e-6
    sky_rows = torch.cat([torch.arange(c.n, c.n + c.n_sky) + k * (c.n + c.n_sky) for k in range(c.K)])
    assert bool(((e.last["flags"].cpu()[sky_rows] & ops.FLAG_OPAQUE) == 0).all())        # depth > far: transparent
    mine = torch.stack([p.grad.cpu() if p.grad is not None else torch.zeros(6) for p in e.poses6])
    errs = dict(depth=relerr(o["depth"], g["depth_fine"]), loss=abs(float(loss) - float(g["loss"])) / float(g["loss"]),
                depth_eps=abs(float(e.last["depth_eps"]) - float(g["depth_eps"])) / float(g["depth_eps"]),
                pose=norm_relerr(mine, g["grad_poses"]))
    print("sky rays: " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert errs["depth"] < 1e-4 and errs["loss"] < 1e-4 and errs["depth_eps"] < 1e-4 and errs["pose"] < 5e-2
    # the detached pose: the same step with the sky rows NOT flagged moves the pose gradient
    e2 = _engine(c, n_sky=c.n_sky)
    for k in range(c.K):
        e2.add_keyframe(c.scans[k].ray_directions, c.scans[k].distances, c.poses6[k], sky_rays=c.sky_dirs[k])
    e2.new_phase(optimize_poses=True)
    inj2 = dict(inj, ray_kf=inj["ray_kf"] & ops.KF_MASK)
    e2.step(list(range(c.K)), c.n, optimize_poses=True, injected=inj2)
    attached = torch.stack([p.grad.cpu() if p.grad is not None else torch.zeros(6) for p in e2.poses6])
    assert norm_relerr(attached, g["grad_poses"]) > 10 * errs["pose"] or norm_relerr(attached, mine) > 1e-3


def test_ray_pick_kernel_segments_and_range_checks():
    wc = synth.world_cube("canteen")
    scans, poses = synth.make_window("canteen", 2, seed=0, n_beams=16, n_azimuth=256)
    M = scans[0].distances.shape[0]
    sky = synth.sky_directions(50, 1)
    cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_neurons=128, n_hidden_layers=2,
                           n_samples=64, n_sky=16)
    e = eng.MappingEngine(cfg)
    e.add_keyframe(scans[0].ray_directions, scans[0].distances, synth.axis_angle_from_yaw_pose(poses[0]), sky_rays=sky)
    e.add_keyframe(scans[1].ray_directions, scans[1].distances, synth.axis_angle_from_yaw_pose(poses[1]))   # no sky set
    n = 4096
    rk, rp = e._pick_rays([0, 1], n)
    rk, rp = rk.cpu(), rp.cpu()
    assert rk.numel() == n + 16 + n                                    # keyframe 1 has no sky directions: no sky segment
    assert bool((rk[:n] == 0).all()) and bool((rk[n:n + 16] == (0 | ops.KF_DETACHED)).all()) and bool((rk[n + 16:] == 1).all())
    assert int(rp[:n].min()) >= 0 and int(rp[:n].max()) < M and rp[:n].unique().numel() > 0.55 * M * (1 - np.exp(-n / M))
    assert int(rp[n:n + 16].min()) >= M and int(rp[n:n + 16].max()) < M + 50
    assert int(rp[n + 16:].min()) >= M + 50 and int(rp[n + 16:].max()) < 2 * M + 50
    hist = torch.bincount(rp[:n] * 8 // M, minlength=8).float()        # uniform over the scan (chi-square, 7 dof)
    assert float(((hist - n / 8) ** 2 / (n / 8)).sum()) < 30.0
    rk2, rp2 = e._pick_rays([0, 1], n)
    e.global_step += 1
    rp_a = rp2.cpu().clone()
    _, rp3 = e._pick_rays([0, 1], n)
    assert torch.equal(rp_a, rp) and not torch.equal(rp3.cpu(), rp)   # a function of (seed, step), new draws every step
    loss = e.step([0, 1], 128)
    assert torch.isfinite(loss)
    # the reference raises IndexError for indices past the scan (torch indexing); so do the injected / FIXED paths
    with pytest.raises(IndexError):
        e.step([0, 1], 4, injected=dict(ray_point=torch.tensor([0, 1, 2, 3, 0, 1, 2, 3])))     # keyframe 1 rows point into keyframe 0
    e.cfg.rays_selection = "FIXED"
    e._wcache = {}
    with pytest.raises(IndexError):
        e.step([0, 1], M + 1)


# ------------------------------------------------------------------------------------------ occupancy update
def test_occupancy_update_skips_rays_dropped_by_build_lidar_rays():
    """rows without LONER_FLAG_VALID never reach `points_fine` in the reference (ray_utils.py:321-322), so
    they must not touch the grid.  Sensor near a cube face: part of the scan exits the cube within near + 1/scale."""
    wc = synth.world_cube("canteen")
    scale, shift = wc.scale_factor, torch.tensor(wc.shift)
    scans, _ = synth.make_window("canteen", 1, seed=0, n_beams=16, n_azimuth=256)
    p6 = torch.tensor([scale * 0.985 - float(shift[0]), -float(shift[1]), -float(shift[2]), 0.0, 0.0, 0.0])   # origin x = 0.985 (cube units)
    n, S = 512, 64
    g = torch.Generator().manual_seed(2)
    idx = torch.randint(0, scans[0].distances.shape[0], (n,), generator=g)
    rays_o, depths_o, keep = orc.build_lidar_rays(scans[0].ray_directions, scans[0].distances, idx, orc.pose6_to_matrix(p6),
                                                  (1.0, 50.0), scale, shift)
    assert 0 < int(keep.sum()) < n, "the geometry must drop some rays and keep others"
    points = ops.pack_points(scans[0].ray_directions, scans[0].distances).to(DEV)
    P = orc.pose6_to_matrix(p6)
    poses12 = torch.cat([P[:3, :3].reshape(-1), P[:3, 3]])[None].to(DEV)
    counters = torch.zeros(2, dtype=torch.int32, device=DEV)
    rays, depths, flags = ops.ray_build(points, torch.zeros(n, dtype=torch.int32, device=DEV), idx.to(DEV), poses12,
                                        shift.tolist(), scale, (1.0, 50.0), counters)
    assert torch.equal((flags.cpu() & 1).bool(), keep)
    z = ops.sample_uniform(rays, S, 0.0)
    grid0 = synth.trained_occupancy_grid("canteen")
    dg = ops.ogm_grad(rays, z, depths, scale, 100, flags=flags)
    after = grid0[0, 0].to(DEV) - 1e-4 * dg
    zk = z.cpu()[keep]
    pts = rays_o[:, None, 0:3] + rays_o[:, None, 3:6] * zk[:, :, None]
    ref = orc.occupancy_step(grid0, pts, zk * scale, depths_o.reshape(-1, 1) * scale, 1e-4)
    d_ref, d_mine = (ref - grid0).flatten(), (after.cpu() - grid0[0, 0]).flatten()
    err = float((d_ref - d_mine).abs().max() / d_ref.abs().max())
    unfiltered = grid0[0, 0].to(DEV) - 1e-4 * ops.ogm_grad(rays, z, depths, scale, 100)
    err_unf = float((d_ref - (unfiltered.cpu() - grid0[0, 0]).flatten()).abs().max() / d_ref.abs().max())
    print(f"occupancy update with {n - int(keep.sum())} dropped rays: rel {err:.2e} (without the flags: {err_unf:.2e})")
    assert err < 1e-4 and err_unf > 10 * err


# ------------------------------------------------------------------------------------------ schedule / decay
def test_lr_schedule_and_los_lambda_decay():
    """ExponentialLR(lrate_gamma) (optimizer.py:269,378) and decay_los_lambda (optimizer.py:448-452)."""
    c = Case("kf2_2x128_fp16")
    e = _engine(c, lrate_gamma=0.5, decay_los_lambda=True, los_lambda_decay_rate=0.5, los_lambda_decay_steps=1.0,
                min_los_lambda=100.0)
    for k in range(c.K):
        e.add_keyframe(c.scans[k].ray_directions, c.scans[k].distances, c.poses6[k])
    e.new_phase(optimize_poses=False)
    ray_point = torch.cat([c.idx[k] + e.kf_offsets[k] for k in range(c.K)])
    inj = dict(ray_point=ray_point, u1=c.u1, u2=c.u2, noise=c.noise)
    p = c.params.clone()
    m, v = torch.zeros_like(p), torch.zeros_like(p)
    for t in range(3):
        # the loss the reference would report with the decayed weight at this global step
        lam = max(1000.0 * 0.5 ** ((t + 1) / 1.0), 100.0)
        pr = e.params.detach().cpu().clone().requires_grad_(True)
        _, _, _, out = orc.mapping_iteration(c.scans, c.poses6, c.idx, pr, c.spec, e.grid.cpu()[None, None], c.S, c.scale,
                                             c.shift, c.ray_range, 1.0, c.u1, c.u2, c.noise, orc.LossCfg(los_lambda=lam))
        loss = e.step(list(range(c.K)), c.n, injected=inj)
        assert abs(float(loss) - float(out["loss"])) / float(out["loss"]) < 1e-4, t
        p, m, v = orc.adam_update(p, e.d_params.cpu(), m, v, t + 1, 0.01 * 0.5 ** t)
        assert relerr(e.params, p) < 1e-6, t


# ------------------------------------------------------------------------------------------ width 64 (BASELINE C1)
@pytest.mark.parametrize("name", ["c1_2x64_fp16", "kf3_2x64_fp16"])
def test_width_64_network_on_gpu(name):
    """BASELINE config 1 (2048 rays x 128 samples, 2 x 64) and a 3-keyframe 2 x 64 case: the 64-wide network runs
    zero-padded on the 128-wide tensor-core kernels (exact), flat parameters and gradients in tcnn's 64-wide layout."""
    c = Case(name)
    e = _engine(c)
    for k in range(c.K):
        e.add_keyframe(c.scans[k].ray_directions, c.scans[k].distances, c.poses6[k])
    assert e.params.numel() == 64 * 64 + 64 * 64 + 16 * 64
    e.new_phase(optimize_poses=c.pose_grads)
    ray_point = torch.cat([c.idx[k] + e.kf_offsets[k] for k in range(c.K)])
    loss = e.step(list(range(c.K)), c.n, optimize_poses=c.pose_grads,
                  injected=dict(ray_point=ray_point, u1=c.u1, u2=c.u2, noise=c.noise), want_outputs=True)
    torch.cuda.synchronize()
    g, o = c.g, e.last["outs"][0]
    r = c.run_oracle()
    errs = dict(depth=relerr(o["depth"], g["depth_fine"]), loss=abs(float(loss) - float(g["loss"])) / float(g["loss"]),
                d_params=norm_relerr(e.d_params, r["params"].grad),
                d_params_fixture=relerr(e.d_params.cpu(), g["grad_params"]))
    print(f"[{name}] " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert errs["depth"] < 1e-4 and errs["loss"] < 1e-4 and errs["d_params"] < 2e-3


# ------------------------------------------------------------------------------------------ checkpoint surface
def test_checkpoint_surface_loads_through_the_reference_shaped_modules():
    """what Mapper.build_ckpt saves (mapper.py:161-175) must load where the analysis scripts load it
    (renderer_lidar.py:172-180, compute_l1_depth.py:146-153): OccupancyGridModel [1,1,V,V,V], Model's sigma params,
    torch-shaped optimiser state dicts."""
    from test_gpu_fused_optimizer import _KF, _settings
    from loner_b200 import dropin
    from loner_b200.dropin.mapping_optimizer import FusedOptimizer
    wc = synth.world_cube("canteen")
    world_cube = types.SimpleNamespace(scale_factor=torch.tensor(wc.scale_factor), shift=torch.tensor(wc.shift))
    st = _settings(n_first=3, n_joint=2)
    opt = FusedOptimizer(st, None, world_cube, 0, False, True, False)
    scans, poses = synth.make_window("canteen", 2, seed=3, n_beams=16, n_azimuth=256)
    kfs = [_KF(scans[k], synth.axis_angle_from_yaw_pose(poses[k]), 3.0 * k) for k in range(2)]
    opt.iterate_optimizer([kfs[0]])
    opt.iterate_optimizer(kfs)
    ckpt = {"network_state_dict": opt._model.state_dict(), "optimizer_state_dict": opt._optimizer.state_dict(),
            "occ_model_state_dict": opt._occupancy_grid_model.state_dict(),
            "occ_optimizer_state_dict": opt._occupancy_grid_optimizer.state_dict()}
    assert tuple(ckpt["occ_model_state_dict"]["occupancy_grid"].shape) == (1, 1, 100, 100, 100)
    path = dropin.install()
    try:
        mt = importlib.import_module("models.model_tcnn")
        mc = st.model_config.model
        occ = mt.OccupancyGridModel(mc.occ_model)
        occ.load_state_dict(ckpt["occ_model_state_dict"])                                   # strict
        assert torch.equal(occ().cpu(), opt._engine.grid.cpu()[None, None])
        cfg = _Cfg(model_type="nerf_decoupled", num_colors=3, ray_range=[1, 50],
                   nerf_config=dict(mc.nerf_config, enable_view_dependence=True),
                   render=dict(mc.render))
        model = mt.Model(cfg)
        model.load_state_dict(ckpt["network_state_dict"])                                   # strict
        assert torch.equal(model.nerf_model._model_sigma.params.detach().cpu(), opt._engine.params.cpu())
    finally:
        sys.path.remove(path)
        for m in list(sys.modules):
            if m == "models" or m.startswith("models."):
                del sys.modules[m]
    # optimiser state dicts load into the torch optimisers the reference constructs (optimizer.py:108-109, :259-267)
    w = torch.nn.Parameter(torch.zeros_like(opt._engine.params))
    pose = torch.nn.Parameter(torch.zeros(6, device=DEV))
    adam = torch.optim.Adam([{"params": [w], "lr": 0.01}, {"params": [pose], "lr": 0.001}])
    adam.load_state_dict(ckpt["optimizer_state_dict"])
    assert int(adam.state[w]["step"]) == opt._engine.adam_t and torch.equal(adam.state[w]["exp_avg"], opt._engine.exp_avg)
    sgd = torch.optim.SGD([torch.nn.Parameter(torch.zeros(1, 1, 100, 100, 100))], lr=1e-4)
    sgd.load_state_dict(ckpt["occ_optimizer_state_dict"])
    # the flat exchange views survive: a state dict loaded back reproduces the parameters
    opt._model.load_state_dict(ckpt["network_state_dict"])
    opt._occupancy_grid_model.load_state_dict(ckpt["occ_model_state_dict"])
    with pytest.raises(RuntimeError):
        opt._occupancy_grid_model.load_state_dict({"occupancy_grid": torch.zeros(1, 1, 50, 50, 50)})


def test_fused_optimizer_sky_gt_poses_and_decay_settings():
    """enable_sky_segmentation + num_samples.sky are honoured (no silent drop), use_gt_poses builds the rays from the
    ground-truth pose (keyframe.py:82-85), lrate_gamma / decay_los_lambda reach the engine."""
    from test_gpu_fused_optimizer import _KF, _Pose, _settings
    from loner_b200.dropin.mapping_optimizer import FusedOptimizer
    wc = synth.world_cube("canteen")
    world_cube = types.SimpleNamespace(scale_factor=torch.tensor(wc.scale_factor), shift=torch.tensor(wc.shift))
    st = _settings(n_first=4, n_joint=2)
    st["model_config"]["train"]["lrate_gamma"] = 0.9
    st["model_config"]["loss"].update(decay_los_lambda=True, los_lambda_decay_rate=0.99, los_lambda_decay_steps=1,
                                      min_los_lambda=100.0)
    scans, poses = synth.make_window("canteen", 1, seed=3, n_beams=16, n_azimuth=256)
    scans[0].sky_rays = synth.sky_directions(64, 5)
    opt = FusedOptimizer(st, None, world_cube, 0, True, True, True)
    assert opt._engine.cfg.n_sky == 64 and opt._engine.cfg.lrate_gamma == 0.9 and opt._engine.cfg.decay_los_lambda
    kf = _KF(scans[0], synth.axis_angle_from_yaw_pose(poses[0]) + torch.tensor([0.5, 0, 0, 0, 0, 0.0]), 0.0)
    gt = synth.axis_angle_from_yaw_pose(poses[0])
    kf._frame = types.SimpleNamespace(_gt_lidar_pose=_Pose(gt))
    losses = opt.iterate_optimizer([kf])
    assert torch.isfinite(losses[0]).all()
    assert torch.equal(opt._engine.poses6[0].cpu(), gt)                     # rays were built from the ground-truth pose
    assert opt._engine.last["rays"].shape[0] == 256 + 64                    # num_samples.lidar + num_samples.sky
    assert opt._engine.kf_sky_sizes == [64]


# ------------------------------------------------------------------------------------------ multi-GPU
@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs (run under gpurun --gpus 2)")
def test_two_rank_nccl_step_equals_single_gpu_step():
    """2-GPU loss / MLP gradient / pose gradients == 1-GPU on one global ray set."""
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", "29631", os.path.join(HERE, "mgpu_grad_check.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    line = [l for l in out.stdout.splitlines() if l.startswith("{")][-1]
    r = json.loads(line)
    print("2-rank NCCL grad check:", r)
    assert r["world"] == 2
    assert r["loss_rel"] < 1e-5 and r["d_params_rel"] < 1e-5 and r["pose_grad_rel"] < 1e-4


def test_pose_kernels_match_autograd_and_torch_adam():
    """loner_pose_matrices / loner_pose_step (a5 + the pose group of optimizer.py:249-267) against the PyTorch
    restatement of tensor_to_transform with autograd and torch.optim.Adam: matrices 1e-6, gradients 1e-5 of the
    largest entry, three Adam steps on the free rows 1e-6; anchored rows do not move and get a zero gradient."""
    from gpu_util import poses6_to_poses12
    g = torch.Generator().manual_seed(3)
    n = 7
    store = torch.randn(n, 6, generator=g)
    store[1, 3:] = 0.0                         # identity rotation (norm gradient 0 at the origin)
    store[2, 3:] = torch.tensor([3e-7, -2e-7, 1e-7])      # Taylor branch
    store[3, 3:] *= 2.5                        # angle > pi
    rows = torch.tensor([5, 0, 3, 2, 1], dtype=torch.int32)
    free = torch.tensor([0, 1, 1, 1, 0, 1, 1], dtype=torch.uint8)
    d12 = torch.randn(3, rows.numel(), 12, generator=g)
    # reference: autograd + torch.optim.Adam over the free rows
    leaves = [store[k].clone().requires_grad_(bool(free[k])) for k in range(n)]
    opt = torch.optim.Adam([p for p in leaves if p.requires_grad], lr=1e-2)
    ref_mats, ref_grads = [], []
    for it in range(3):
        opt.zero_grad(set_to_none=True)
        p12 = poses6_to_poses12(torch.stack([leaves[int(k)] for k in rows]))
        ref_mats.append(p12.detach().clone())
        p12.backward(d12[it])
        ref_grads.append(torch.stack([p.grad.clone() if p.grad is not None else torch.zeros(6) for p in leaves]))
        opt.step()
    # kernels
    st = store.to(DEV).contiguous()
    rows_d, free_d = rows.to(DEV), free.to(DEV)
    grad6 = torch.zeros(n, 6, device=DEV)
    m, v = torch.zeros(n, 6, device=DEV), torch.zeros(n, 6, device=DEV)
    steps = torch.zeros(n, dtype=torch.int32, device=DEV)
    for it in range(3):
        mats = ops.pose_matrices(st, rows_d)
        assert (mats.cpu() - ref_mats[it]).abs().max() < 2e-6
        grad6.zero_()
        ops.pose_step(st, rows_d, free_d, d12[it].to(DEV).contiguous(), grad6, m, v, steps, 1e-2)
        err = (grad6.cpu() - ref_grads[it]).abs().max() / ref_grads[it].abs().max()
        assert err < 1e-5, err
    want = torch.stack([p.detach() for p in leaves])
    assert (st.cpu() - want).abs().max() < 2e-6
    assert torch.equal(st.cpu()[[0, 4]], store[[0, 4]])                 # anchored rows untouched
    assert steps.cpu().tolist() == [0, 3, 3, 3, 0, 3, 0]                # row 6 is free but outside the window


def test_pose_guards_set_status_and_raise_like_the_reference():
    """optimizer.py:368-374 and ray_utils.py:301-303 as a lazily checked device status word."""
    st = torch.zeros(3, 6, device=DEV)
    st[1, :3] = torch.tensor([1000.0, 0.0, 0.0], device=DEV)          # far outside any world cube
    rows = torch.tensor([0, 1, 2], dtype=torch.int32, device=DEV)
    status = torch.zeros(1, dtype=torch.int32, device=DEV)
    ops.pose_matrices(st, rows[:1], shift=(0.0, 0.0, 0.0), scale=50.0, status=status)
    assert int(status.item()) == 0
    ops.pose_matrices(st, rows, shift=(0.0, 0.0, 0.0), scale=50.0, status=status)
    assert int(status.item()) == ops.STATUS_ORIGIN_OUTSIDE
    # a non-finite gradient: the row keeps its pose, the finite row steps
    status.zero_()
    free = torch.ones(3, dtype=torch.uint8, device=DEV)
    d12 = torch.ones(3, 12, device=DEV)
    d12[2, 10] = float("nan")
    grad6, m, v = (torch.zeros(3, 6, device=DEV) for _ in range(3))
    steps = torch.zeros(3, dtype=torch.int32, device=DEV)
    before = st.clone()
    ops.pose_step(st, rows, free, d12, grad6, m, v, steps, 1e-2, status=status)
    assert int(status.item()) == ops.STATUS_BAD_POSE_GRAD
    assert torch.equal(st[2], before[2]) and not torch.equal(st[0], before[0])
    assert steps.cpu().tolist() == [1, 1, 0]
    # through the engine: a keyframe whose origin leaves the cube raises the reference's assertion at check_status()
    c = Case("kf2_2x128_fp16")
    e = _engine(c)
    for k in range(c.K):
        e.add_keyframe(c.scans[k].ray_directions, c.scans[k].distances, c.poses6[k])
    e.new_phase(False)
    e.step([0, 1], 64)
    e.check_status()
    e.poses6[1].data[:3] += 10.0 * c.scale
    e._pose_cache = None
    e.step([0, 1], 64)
    with pytest.raises(AssertionError, match="outside the world cube"):
        e.check_status()
    e.pose_store[1, 3] = float("inf")
    e._pose_cache = None
    e.step([0, 1], 64)
    with pytest.raises((AssertionError, RuntimeError)):
        e.check_status()
