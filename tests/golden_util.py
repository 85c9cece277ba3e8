"""Rebuilds the inputs of a golden case (tests/golden/*.npz, minted by oracle/make_golden.py from
the reference's own code) from the seeds stored in the fixture, and runs the oracle on them."""
import glob
import os
import sys

import numpy as np
import torch

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

from loner_b200 import synth  # noqa: E402
from oracle import hashgrid_standin  # noqa: E402
from oracle import loner_oracle as orc  # noqa: E402
from oracle import tcnn_standin  # noqa: E402

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")


def golden_names():
    """Training-iteration fixtures (the test-mode render fixtures `testmode_*` have their own loader)."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if not n.startswith("testmode_")]


def sky_randoms(seed, K, n_sky, sky_count):
    """Same protocol as oracle/make_golden.py::sky_randoms."""
    g = torch.Generator().manual_seed(seed + 2)
    dirs = [synth.sky_directions(sky_count, seed + 100 + k) for k in range(K)]
    idx = [torch.randint(0, sky_count, (n_sky,), generator=g) for _ in range(K)]
    return dirs, idx


def case_randoms(seed, n_per_kf, K, M, n_rays, S, sampler="OGM"):
    """Same generator protocol as oracle/make_golden.py::case_randoms."""
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


class Case:
    def __init__(self, name):
        self.name = name
        self.g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        seed, K, n, S, L, W, nb, naz, n_rays = [int(v) for v in self.g["meta"]]
        self.seed, self.K, self.n, self.S, self.L, self.W, self.n_rays = seed, K, n, S, L, W, n_rays
        self.geom = str(self.g["geom"])
        self.prec = str(self.g["prec"])
        self.scale = float(self.g["scale"])
        self.shift = torch.from_numpy(self.g["shift"])
        self.ray_range = synth.GEOMETRY[self.geom]["ray_range"]
        self.scans, poses = synth.make_window(self.geom, K, seed=7, n_beams=nb, n_azimuth=naz)
        self.poses6 = [synth.axis_angle_from_yaw_pose(poses[k]) for k in range(K)]
        self.M = self.scans[0].distances.shape[0]
        self.sampler = str(self.g["sampler"]) if "sampler" in self.g.files else "OGM"
        self.idx, self.u1, self.u2, self.noise = case_randoms(seed, n, K, self.M, n_rays, S, self.sampler)
        self.hash_spec = None
        if "hash_cfg" in self.g.files and self.g["hash_cfg"].size:      # the reference's shipped HashGrid sigma head
            nl, nf, lt, br = [int(v) for v in self.g["hash_cfg"]]
            self.hash_spec = hashgrid_standin.HashGridSpec(n_levels=nl, n_features_per_level=nf, log2_hashmap_size=lt,
                                                           base_resolution=br)
        self.spec = orc.NetSpec(n_frequencies=10, n_neurons=W, n_hidden_layers=L, precision=self.prec, hash=self.hash_spec)
        self.params = tcnn_standin.xavier_uniform_flat(self.spec.shapes, int(self.g["params_seed"]))
        if self.hash_spec is not None:
            self.params = torch.cat([self.params, hashgrid_standin.init_table(self.hash_spec, int(self.g["table_seed"]),
                                                                              float(self.g["table_scale"]))])
        if str(self.g["grid"]) == "trained":
            self.grid = synth.trained_occupancy_grid(self.geom)
        else:
            self.grid = torch.zeros(1, 1, 100, 100, 100)
        self.n_sky, self.sky_count = ([int(v) for v in self.g["sky"]] if "sky" in self.g.files else (0, 0))
        self.sky_dirs, self.sky_idx = (sky_randoms(seed, K, self.n_sky, self.sky_count) if self.n_sky
                                       else (None, None))
        self.pose_grads = "grad_poses" in self.g.files
        self.loss_selection = str(self.g["loss_selection"]) if "loss_selection" in self.g.files else "L1_JS"
        # iteration_idx = 0 in the fixtures: the *_LOS margin is depth_eps * 0.95**0 = 3.0 (default_model_config.yaml:51-55)
        self.loss_cfg = orc.LossCfg(loss_selection=self.loss_selection, fixed_eps=3.0)

    def run_oracle(self):
        params = self.params.clone().requires_grad_(True)
        poses6 = [p.clone().requires_grad_(self.pose_grads and k > 0) for k, p in enumerate(self.poses6)]
        rays, depths, res, out = orc.mapping_iteration(
            self.scans, poses6, self.idx, params, self.spec, self.grid, self.S, self.scale, self.shift,
            self.ray_range, 1.0, self.u1, self.u2, self.noise, self.loss_cfg, sampler=self.sampler,
            sky=(self.sky_dirs, self.sky_idx) if self.n_sky else None)
        out["loss"].backward()
        s = res["samples_fine"].detach() * self.scale
        G = depths.reshape(-1, 1) * self.scale
        # the occupancy grid only exists (and is only stepped) with the OGM sampler   optimizer.py:102-118,382-384
        grid_after = orc.occupancy_step(self.grid, res["points_fine"], s, G, 1e-4) if self.sampler == "OGM" else self.grid
        return dict(rays=rays, depths=depths, res=res, out=out, params=params, poses6=poses6,
                    grid_after=grid_after)


class TestModeCase:
    """tests/golden/testmode_*.npz: Model.forward(testing=True) over a chunked scan + the depth-L1 metric, minted
    from the reference (oracle/make_golden.py::run_testmode)."""
    __test__ = False

    def __init__(self, name):
        self.g = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        seed, n, chunk, S, L, W, nb, naz = [int(v) for v in self.g["meta"]]
        self.seed, self.n, self.chunk, self.S, self.L, self.W = seed, n, chunk, S, L, W
        self.geom = str(self.g["geom"])
        self.prec = str(self.g["prec"])
        self.scale = float(self.g["scale"])
        self.shift = torch.from_numpy(self.g["shift"])
        self.ray_range = synth.GEOMETRY[self.geom]["ray_range"]
        scans, poses = synth.make_window(self.geom, 1, seed=7, n_beams=nb, n_azimuth=naz)
        sel = torch.from_numpy(self.g["sel"])
        self.directions = scans[0].ray_directions[:, sel]
        self.distances = scans[0].distances[sel]
        self.pose6 = synth.axis_angle_from_yaw_pose(poses[0])
        g2 = torch.Generator().manual_seed(seed)
        self.u2 = torch.rand(n, S // 2, generator=g2)
        self.noise = torch.randn(n, S, generator=g2)
        self.spec = orc.NetSpec(n_frequencies=10, n_neurons=W, n_hidden_layers=L, precision=self.prec)
        self.params = tcnn_standin.xavier_uniform_flat(self.spec.shapes, int(self.g["params_seed"]))
        self.grid = synth.trained_occupancy_grid(self.geom)

    def run_oracle(self):
        """fetch_chunk_rays -> render (perturb = 0) -> depth-L1, through the oracle."""
        rays, depths, keep = orc.build_lidar_rays(self.directions, self.distances, torch.arange(self.n),
                                                  orc.pose6_to_matrix(self.pose6), self.ray_range, self.scale, self.shift)
        assert bool(keep.all())
        with torch.no_grad():
            z = orc.ogm_samples(rays.float(), self.grid, self.S, 0.0, None, self.u2)
            res = orc.render_rays(rays.float(), z, self.params, self.spec, self.noise)
        l1 = orc.depth_l1_metric(res["depth_fine"], depths, self.scale, self.ray_range)
        return rays.float(), depths, res, l1
