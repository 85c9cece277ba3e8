"""TEST INFRASTRUCTURE — mints tests/golden/*.npz by executing the REFERENCE'S OWN Python on CPU.

Run in the build container only (needs /root/reference):   python -m oracle.make_golden
The reference code path executed, unmodified, per case:
  KeyFrame.build_lidar_rays (mapping/keyframe.py:71) -> LidarRayDirections.build_lidar_rays
  (common/ray_utils.py:269) -> Optimizer.compute_loss (mapping/optimizer.py:437) -> Model.forward
  (models/model_tcnn.py:70) -> render_rays (models/rendering_tcnn.py:192) ->
  OccGridRaySampler.get_samples (models/ray_sampling.py:53) -> DecoupledNeRF.forward
  (models/nerf_tcnn.py:59; tcnn replaced by oracle/tcnn_standin.py) -> raw2outputs ->
  JS margin + 3 losses -> loss.backward() -> Optimizer._step_occupancy_grid (optimizer.py:598).
Randomness (randint / rand / randn) is replayed from seeded CPU generators so that the same
numbers can be regenerated from the seeds stored in the fixture (see `case_randoms`).
"""
import os
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from loner_b200 import synth  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle import hashgrid_standin, tcnn_standin  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")

CASES = {
    # name: geometry, K keyframes, rays per KF, S, hidden layers, width, grid, precision, pose grads
    "c1_2x64_fp32":   dict(geom="canteen", K=1, n=2048, S=128, L=2, W=64, grid="zeros", prec="fp32", poses=False, rows=64),
    "c1_2x64_fp16":   dict(geom="canteen", K=1, n=2048, S=128, L=2, W=64, grid="zeros", prec="fp16", poses=False, rows=64),
    "kf3_2x64_fp16":  dict(geom="garden", K=3, n=96, S=128, L=2, W=64, grid="trained", prec="fp16", poses=True, rows=288),
    "kf2_4x256_fp16": dict(geom="canteen", K=2, n=128, S=256, L=4, W=256, grid="trained", prec="fp16", poses=True, rows=64),
    "kf2_4x256_fp32": dict(geom="canteen", K=2, n=128, S=256, L=4, W=256, grid="trained", prec="fp32", poses=True, rows=64),
    "kf2_2x128_fp16": dict(geom="garden", K=2, n=128, S=128, L=2, W=128, grid="trained", prec="fp16", poses=True, rows=64),
    "kf2_2x128_l2js": dict(geom="garden", K=2, n=128, S=128, L=2, W=128, grid="trained", prec="fp16", poses=True, rows=64, loss="L2_JS"),
    "kf2_2x128_l1los": dict(geom="garden", K=2, n=128, S=128, L=2, W=128, grid="trained", prec="fp16", poses=True, rows=64, loss="L1_LOS"),
    "quad_4x256_fp16": dict(geom="quad", K=2, n=128, S=512, L=4, W=256, grid="trained", prec="fp16", poses=True, rows=32),
    # the reference's SHIPPED sigma head (cfg/nerf_config/default_nerf_hash.yaml): HashGrid + 1 x 64, with a
    # smaller table (2^14 entries per level) to keep the fixture's regeneration and the oracle run cheap
    "hash_1x64_fp16": dict(geom="canteen", K=2, n=128, S=128, L=1, W=64, grid="trained", prec="fp16", poses=True, rows=64,
                           hash=dict(n_levels=16, n_features_per_level=2, log2_hashmap_size=14, base_resolution=16)),
}
CASES["kf2_2x128_uni_l2los"] = dict(geom="garden", K=2, n=128, S=128, L=2, W=128, grid="zeros", prec="fp16", poses=True,
                                     rows=64, loss="L2_LOS", sampler="UNIFORM")   # the two branches no other case takes
# sky rays: num_samples.sky picks per keyframe among LidarScan.sky_rays, built from the
# detached pose at distance ray_range[1] + 1 (optimizer.py:299-305, keyframe.py:87-99, sensors.py:162-167)
CASES["kf2_2x128_sky"] = dict(geom="garden", K=2, n=96, S=128, L=2, W=128, grid="trained", prec="fp16", poses=True,
                               rows=64, n_sky=32, sky_count=200)
TABLE_SEED, TABLE_SCALE = 4242, 0.5
# test-mode render + the depth-L1 metric: Model.forward(testing=True) at N_samples_test over a
# chunked scan, then the L1 of analysis/compute_l1_depth.py:42-64
TESTMODE = {
    "testmode_2x128": dict(geom="garden", n=256, chunk=128, S_test=2048, L=2, W=128, grid="trained", prec="fp16"),
}
N_BEAMS, N_AZ = 16, 256   # 4,096-point scans keep the fixtures' regeneration cheap


def sky_randoms(seed, K, n_sky, sky_count):
    """Sky directions per keyframe and the replayed `torch.randint(0, sky_dirs.shape[1], (n_sky,))` draws."""
    g = torch.Generator().manual_seed(seed + 2)
    dirs = [synth.sky_directions(sky_count, seed + 100 + k) for k in range(K)]
    idx = [torch.randint(0, sky_count, (n_sky,), generator=g) for _ in range(K)]
    return dirs, idx


def case_randoms(seed, n_per_kf, K, M, n_rays, S, sampler="OGM"):
    """The replayed random numbers, regenerable from `seed` alone (CPU torch generators).  The OGM sampler
    draws two [n, S/2] uniforms (ray_sampling.py:71-72, rendering_tcnn.py:48), the uniform sampler one [n, S]
    (ray_sampling.py:39)."""
    g = torch.Generator().manual_seed(seed)
    idx = [torch.randint(0, M, (n_per_kf,), generator=g) for _ in range(K)]
    g2 = torch.Generator().manual_seed(seed + 1)
    if sampler == "UNIFORM":
        u1 = torch.rand(n_rays, S, generator=g2)
        return idx, u1, None, torch.randn(n_rays, S, generator=g2)
    u1 = torch.rand(n_rays, S // 2, generator=g2)
    u2 = torch.rand(n_rays, S // 2, generator=g2)
    noise = torch.randn(n_rays, S, generator=g2)
    return idx, u1, u2, noise


def build_reference_optimizer(ns, geom, S, L, W, tmpdir, loss="L1_JS", hash_cfg=None, sampler="OGM"):
    s = rh.load_settings()
    s["mapper"]["optimizer"]["samples_selection"]["strategy"] = sampler
    g = synth.GEOMETRY[geom]
    opt_s = s["mapper"]["optimizer"]
    mc = opt_s["model_config"]
    mc["data"]["ray_range"] = list(g["ray_range"])
    mc["model"]["ray_range"] = list(g["ray_range"])
    mc["model"]["render"]["N_samples_train"] = S
    mc["loss"]["loss_selection"] = loss
    nc = mc["model"]["nerf_config"]
    nc["pos_encoding_sigma"] = {"otype": "Frequency", "n_frequencies": 10}
    if hash_cfg is not None:
        nc["pos_encoding_sigma"] = dict(otype="HashGrid", **hash_cfg)
    nc["sigma_network"] = {"otype": "CutlassMLP" if W > 128 else "FullyFusedMLP", "activation": "ReLU",
                           "output_activation": "None", "n_neurons": W, "n_hidden_layers": L}
    opt_s["debug"] = s["debug"]["flags"]
    opt_s["log_directory"] = tmpdir
    wc = ns.pose_utils.compute_world_cube(None, None, None, None, g["ray_range"], padding=0.3,
                                          traj_bounding_box={k: list(v) for k, v in g["bbox"].items()})
    opt = ns.optimizer.Optimizer(s.mapper.optimizer, s.calibration, wc, "cpu", False, True, False)
    return opt, wc, s


def run_case(name, c, seed=1234):
    ns = rh.import_reference()
    tcnn_standin.PRECISION["mode"] = c["prec"]
    torch.manual_seed(0)
    with tempfile.TemporaryDirectory() as tmp:
        opt, wc, settings = build_reference_optimizer(ns, c["geom"], c["S"], c["L"], c["W"], tmp, c.get("loss", "L1_JS"),
                                                      c.get("hash"), c.get("sampler", "OGM"))
        if c.get("hash"):
            sig = opt._model.nerf_model._model_sigma
            with torch.no_grad():
                sig.params[sig.n_network_params:] = hashgrid_standin.init_table(sig.encoding.hash_spec, TABLE_SEED, TABLE_SCALE)
        g = synth.GEOMETRY[c["geom"]]
        scans, poses = synth.make_window(c["geom"], c["K"], seed=7, n_beams=N_BEAMS, n_azimuth=N_AZ)
        M = scans[0].distances.shape[0]
        # occupancy grid
        if c["grid"] == "trained":
            grid0 = synth.trained_occupancy_grid(c["geom"])
            with torch.no_grad():
                opt._occupancy_grid_model.occupancy_grid.copy_(grid0)
            opt._occupancy_grid = opt._occupancy_grid_model()
            opt._ray_sampler.update_occ_grid(opt._occupancy_grid.detach())
        # the reference's own containers
        kfs = []
        n_sky = c.get("n_sky", 0)
        sky_dirs, sky_idx = sky_randoms(seed, c["K"], n_sky, c["sky_count"]) if n_sky else (None, [None] * c["K"])
        for k in range(c["K"]):
            sc = ns.sensors.LidarScan(scans[k].ray_directions.clone(), scans[k].distances.clone(),
                                      scans[k].timestamps.clone(), sky_rays=sky_dirs[k].clone() if n_sky else None)
            fr = ns.frame.Frame(None, sc, None)
            p6 = synth.axis_angle_from_yaw_pose(poses[k])
            fr._lidar_pose = ns.pose.Pose(pose_tensor=p6.clone(), fixed=not (c["poses"] and k > 0))
            fr._gt_lidar_pose = fr._lidar_pose
            kfs.append(ns.keyframe.KeyFrame(fr, "cpu"))
        ray_range = torch.Tensor(list(g["ray_range"]))
        # pass 1 (no randomness needed): ray build, to learn the post-filter ray count
        idx, _, _, _ = case_randoms(seed, c["n"], c["K"], M, 1, c["S"], c.get("sampler", "OGM"))
        rays_l, dep_l = [], []
        for kf, ix, six in zip(kfs, idx, sky_idx):
            r, d = kf.build_lidar_rays(ix, ray_range, wc, False, sky_indices=six)
            rays_l.append(r)
            dep_l.append(d)
        rays = torch.vstack(rays_l).float()
        depths = torch.cat(dep_l).float()
        n_rays = rays.shape[0]
        idx, u1, u2, noise = case_randoms(seed, c["n"], c["K"], M, n_rays, c["S"], c.get("sampler", "OGM"))
        replay = rh.Replay()
        replay.rand = [u1, u2] if u2 is not None else [u1]
        replay.randn = [noise]
        opt._optimization_settings.freeze_poses = not c["poses"]
        with rh.injected_randomness(replay):
            loss = opt.compute_loss(None, (rays, depths), 0)
        assert not replay.rand and not replay.randn, "reference drew fewer randoms than queued"
        params = opt._model.nerf_model._model_sigma.params
        loss.backward()
        res = opt._results_lidar
        if c.get("sampler", "OGM") == "OGM":
            grid_before = opt._occupancy_grid.detach().clone()
            opt._step_occupancy_grid()
            grid_after = opt._occupancy_grid.detach().clone()
            dgrid = grid_after - grid_before
        else:                       # no occupancy grid with the uniform sampler (optimizer.py:102-118, :382-384)
            dgrid = torch.zeros(1, 1, 100, 100, 100)
        nz = dgrid.flatten().nonzero()[:, 0]

        rows = min(c["rows"], n_rays)
        out = dict(
            meta=np.array([seed, c["K"], c["n"], c["S"], c["L"], c["W"], N_BEAMS, N_AZ, n_rays], dtype=np.int64),
            geom=np.array(c["geom"]), grid=np.array(c["grid"]), prec=np.array(c["prec"]),
            scale=np.float32(float(wc.scale_factor)), shift=wc.shift.numpy().astype(np.float32),
            params_seed=np.int64(1337), loss_selection=np.array(c.get("loss", "L1_JS")),
            hash_cfg=np.array([c["hash"][k] for k in ("n_levels", "n_features_per_level", "log2_hashmap_size",
                                                       "base_resolution")] if c.get("hash") else [], dtype=np.int64),
            table_seed=np.int64(TABLE_SEED), table_scale=np.float32(TABLE_SCALE),
            sampler=np.array(c.get("sampler", "OGM")),
            sky=np.array([n_sky, c.get("sky_count", 0)], dtype=np.int64),
            rays=rays.detach().numpy(), depths=depths.numpy(),
            z_vals=res["samples_fine"][:rows].detach().numpy(),
            weights=res["weights_fine"][:rows].detach().numpy(),
            depth_fine=res["depth_fine"].detach().numpy(),
            opacity_fine=res["opacity_fine"].detach().numpy(),
            variance=res["variance"].detach().numpy(),
            loss=np.float32(loss.item()),
            depth_eps=np.float32(opt._depth_eps),
            grad_params_norm=np.float32(params.grad.norm().item()),
            grad_params=params.grad.numpy() if params.numel() <= 20000 else params.grad.numpy()[::16],
            ogm_delta_idx=nz.numpy().astype(np.int64)[:4096],
            ogm_delta_val=dgrid.flatten()[nz].numpy()[:4096],
            ogm_delta_sum=np.float64(dgrid.double().sum().item()),
            ogm_delta_abs=np.float64(dgrid.double().abs().sum().item()),
        )
        if params.numel() <= 20000:
            out["params"] = params.detach().numpy()
        if c["poses"]:
            out["grad_poses"] = np.stack([kf.get_lidar_pose().get_pose_tensor().grad.numpy()
                                          if kf.get_lidar_pose().get_pose_tensor().grad is not None
                                          else np.zeros(6, np.float32) for kf in kfs])
        # loss terms, recomputed by the same reference formulas from the stored results
        return out


def run_testmode(name, c, seed=4321):
    """Model.forward(testing=True) over a chunked scan exactly as analysis/compute_l1_depth.py:42-64 drives it
    (LidarRayDirections.fetch_chunk_rays, ray_utils.py:262-267), with the two draws that stay active in test
    mode replayed (sample_pdf's u, rendering_tcnn.py:48; raw noise, rendering_tcnn.py:104)."""
    ns = rh.import_reference()
    tcnn_standin.PRECISION["mode"] = c["prec"]
    torch.manual_seed(0)
    with tempfile.TemporaryDirectory() as tmp:
        opt, wc, settings = build_reference_optimizer(ns, c["geom"], 128, c["L"], c["W"], tmp)
        g = synth.GEOMETRY[c["geom"]]
        scans, poses = synth.make_window(c["geom"], 1, seed=7, n_beams=N_BEAMS, n_azimuth=N_AZ)
        n, S = c["n"], c["S_test"]
        grid0 = synth.trained_occupancy_grid(c["geom"])
        with torch.no_grad():
            opt._occupancy_grid_model.occupancy_grid.copy_(grid0)
        opt._ray_sampler.update_occ_grid(opt._occupancy_grid_model().detach())
        model = opt._model
        model.cfg["render"]["N_samples_test"] = S
        assert model.cfg.render.N_samples_test == S
        # a scan of the first n returns, spread over the whole sweep
        sel = torch.arange(n) * (scans[0].distances.shape[0] // n)
        scan = ns.sensors.LidarScan(scans[0].ray_directions[:, sel].clone(), scans[0].distances[sel].clone(),
                                    scans[0].timestamps[sel].clone())
        pose = ns.pose.Pose(pose_tensor=synth.axis_angle_from_yaw_pose(poses[0]).clone(), fixed=True)
        dirs = ns.ray_utils.LidarRayDirections(scan, chunk_size=c["chunk"])
        ray_range = torch.Tensor(list(g["ray_range"]))
        g2 = torch.Generator().manual_seed(seed)
        u2 = torch.rand(n, S // 2, generator=g2)
        noise = torch.randn(n, S, generator=g2)
        depth_m = torch.zeros(n, 1)
        var = torch.zeros(n)
        opa = torch.zeros(n)
        rays_all = []
        with torch.no_grad():
            for ci in range(dirs.num_chunks):
                rays = dirs.fetch_chunk_rays(ci, pose, wc, ray_range)
                lo = ci * c["chunk"]
                hi = lo + rays.shape[0]
                assert hi - lo == min(c["chunk"], n - lo), "validity filter dropped a ray: pick another scan subset"
                replay = rh.Replay()
                replay.rand = [u2[lo:hi]]
                replay.randn = [noise[lo:hi]]
                with rh.injected_randomness(replay):
                    res = model(rays.float(), opt._ray_sampler, wc.scale_factor, testing=True, return_variance=True,
                                camera=False)
                assert not replay.rand and not replay.randn
                depth_m[lo:hi, :] = res["depth_fine"].unsqueeze(1) * wc.scale_factor
                var[lo:hi] = res["variance"]
                opa[lo:hi] = res["opacity_fine"]
                rays_all.append(rays.float())
            gt = scan.distances
            good = torch.logical_and(gt.flatten() > ray_range[0], gt.flatten() < ray_range[1] - 0.25)
            l1 = torch.nn.functional.l1_loss(depth_m[good].flatten(), gt[good.flatten()].flatten())
        return dict(meta=np.array([seed, n, c["chunk"], S, c["L"], c["W"], N_BEAMS, N_AZ], dtype=np.int64),
                    geom=np.array(c["geom"]), prec=np.array(c["prec"]), scale=np.float32(float(wc.scale_factor)),
                    shift=wc.shift.numpy().astype(np.float32), params_seed=np.int64(1337),
                    rays=torch.cat(rays_all).numpy(), sel=sel.numpy(), depth_m=depth_m[:, 0].numpy(),
                    variance=var.numpy(), opacity=opa.numpy(), l1=np.float64(l1.item()), n_good=np.int64(int(good.sum())))


def main():
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]
    for name, c in TESTMODE.items():
        if only and name not in only:
            continue
        out = run_testmode(name, c)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: l1={out['l1']:.6f} over {out['n_good']} rays -> {os.path.getsize(path)/1024:.0f} KiB")
    for name, c in CASES.items():
        if only and name not in only:
            continue
        out = run_case(name, c)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(f"{name}: n_rays={out['meta'][-1]} loss={out['loss']:.6f} depth_eps={out['depth_eps']:.4f} "
              f"-> {os.path.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main()
