"""Seeded synthetic LiDAR scans and sequence geometry.

There is no dataset and no network: every benchmark and test runs on analytic scans of a
closed scene (axis-aligned room + ground + three spheres) cast from keyframe poses on an arc
inside the sequence's `trajectory_bounding_box`.  Geometry numbers are those of the reference's
sequence files: /root/reference/cfg/fusion_portable/canteen.yaml:12-19, garden.yaml:12-19,
/root/reference/cfg/newer_college/quad.yaml:12-19; the world cube follows the
bounding-box branch of `compute_world_cube` (/root/reference/src/common/pose_utils.py:159-248)
with `padding=0.3` (/root/reference/src/loner.py:104-105).

Pure torch/CPU, no dependency on the CUDA extension or on oracle/.
"""
import math
from dataclasses import dataclass

import torch

GEOMETRY = {
    "canteen": dict(bbox=dict(x=(-25.0, 10.0), y=(-25.0, 15.0), z=(-10.0, 10.0)),
                    ray_range=(1.0, 50.0), fov=(-22.5, 22.5)),
    "garden": dict(bbox=dict(x=(-5.0, 25.0), y=(-15.0, 20.0), z=(-10.0, 10.0)),
                   ray_range=(1.0, 50.0), fov=(-22.5, 22.5)),
    "quad": dict(bbox=dict(x=(-5.0, 50.0), y=(-25.0, 15.0), z=(-3.0, 10.0)),
                 ray_range=(1.0, 75.0), fov=(-45.0, 45.0)),
    # the reference's defaults.yaml geometry (ray_range [1,10] in default_model_config.yaml:2)
    "default": dict(bbox=dict(x=(-10.0, 10.0), y=(-10.0, 10.0), z=(-10.0, 10.0)),
                    ray_range=(1.0, 10.0), fov=(-22.5, 22.5)),
}


@dataclass
class WorldCubeSpec:
    scale_factor: float
    shift: tuple  # 3 floats


def world_cube(geom: str, padding: float = 0.3) -> WorldCubeSpec:
    """Bounding-box branch of compute_world_cube (pose_utils.py:165-178, :217-248), lidar-only."""
    g = GEOMETRY[geom]
    r1 = g["ray_range"][1]
    lo = torch.tensor([g["bbox"][a][0] - r1 for a in "xyz"], dtype=torch.float32)
    hi = torch.tensor([g["bbox"][a][1] + r1 for a in "xyz"], dtype=torch.float32)
    origin = lo + (hi - lo) / 2
    scale = (torch.linalg.norm(hi - lo) / (2 * torch.sqrt(torch.tensor([3.0])))) * (1 + padding)
    return WorldCubeSpec(float(scale), tuple(float(-o) for o in origin))


def beam_directions(fov, n_beams=64, n_azimuth=1024) -> torch.Tensor:
    """[3, M] unit vectors in the sensor frame, beam-major."""
    el = torch.deg2rad(torch.linspace(fov[0], fov[1], n_beams, dtype=torch.float64))
    az = torch.arange(n_azimuth, dtype=torch.float64) * (2 * math.pi / n_azimuth)
    el, az = torch.meshgrid(el, az, indexing="ij")
    d = torch.stack([torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)], 0)
    return d.reshape(3, -1).float()


def _scene(geom):
    g = GEOMETRY[geom]
    cx = 0.5 * (g["bbox"]["x"][0] + g["bbox"]["x"][1])
    cy = 0.5 * (g["bbox"]["y"][0] + g["bbox"]["y"][1])
    half = 0.45 * g["ray_range"][1] + 6.0
    box_lo = torch.tensor([cx - half, cy - half, -2.0], dtype=torch.float64)
    box_hi = torch.tensor([cx + half, cy + half, 9.0], dtype=torch.float64)
    spheres = [((cx + 6.0, cy + 2.0, 0.0), 2.0), ((cx - 5.0, cy - 7.0, 1.0), 3.0), ((cx + 1.0, cy + 9.0, -0.5), 1.5)]
    return box_lo, box_hi, spheres


def raycast(origin: torch.Tensor, dirs_world: torch.Tensor, geom: str) -> torch.Tensor:
    """Distance along each world-frame direction [3,M] from `origin` [3] to the closed scene."""
    o = origin.double()
    d = dirs_world.double()
    box_lo, box_hi, spheres = _scene(geom)
    dd = torch.where(d.abs() < 1e-12, torch.full_like(d, 1e-12), d)
    t_lo = (box_lo[:, None] - o[:, None]) / dd
    t_hi = (box_hi[:, None] - o[:, None]) / dd
    t_exit = torch.maximum(t_lo, t_hi).min(dim=0)[0]          # inside the box: first wall hit
    t = t_exit
    for c, r in spheres:
        c = torch.tensor(c, dtype=torch.float64)
        oc = o - c
        b = (d * oc[:, None]).sum(0)
        cc = (oc * oc).sum() - r * r
        disc = b * b - cc
        ts = -b - torch.sqrt(disc.clamp(min=0))
        hit = (disc > 0) & (ts > 0)
        t = torch.where(hit & (ts < t), ts, t)
    return t.float()


def rotz(yaw: float) -> torch.Tensor:
    c, s = math.cos(yaw), math.sin(yaw)
    return torch.tensor([[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]])


def keyframe_poses(geom: str, K: int) -> torch.Tensor:
    """[K,4,4] sensor-to-world poses on a smooth arc inside the trajectory bounding box,
    yaw sweeping 0 -> 90 degrees; keyframe 0 is the identity (anchored)."""
    g = GEOMETRY[geom]
    poses = torch.eye(4).repeat(K, 1, 1)
    rx = 0.25 * min(abs(g["bbox"]["x"][0]), abs(g["bbox"]["x"][1]), 8.0)
    ry = 0.25 * min(abs(g["bbox"]["y"][0]), abs(g["bbox"]["y"][1]), 8.0)
    for k in range(K):
        a = 0.0 if K == 1 else (k / (K - 1)) * (math.pi / 2)
        poses[k, :3, :3] = rotz(a)
        poses[k, :3, 3] = torch.tensor([rx * math.sin(a), ry * (1 - math.cos(a)), 0.1 * k / max(K, 1)])
    return poses


def axis_angle_from_yaw_pose(pose: torch.Tensor) -> torch.Tensor:
    """6-vector [t, axis-angle] of a yaw-only pose (the synthetic arc)."""
    yaw = math.atan2(float(pose[1, 0]), float(pose[0, 0]))
    return torch.tensor([float(pose[0, 3]), float(pose[1, 3]), float(pose[2, 3]), 0.0, 0.0, yaw])


@dataclass
class Scan:
    ray_directions: torch.Tensor  # [3, M] sensor frame, unit
    distances: torch.Tensor       # [M] metres
    timestamps: torch.Tensor      # [M] seconds, sorted


def make_scan(geom: str, pose: torch.Tensor, seed: int, t0: float = 0.0,
              n_beams: int = 64, n_azimuth: int = 1024) -> Scan:
    g = GEOMETRY[geom]
    dirs = beam_directions(g["fov"], n_beams, n_azimuth)
    M = dirs.shape[1]
    dist = raycast(pose[:3, 3], pose[:3, :3] @ dirs, geom)
    gen = torch.Generator().manual_seed(seed)
    dist = dist + 0.02 * torch.randn(M, generator=gen)
    dist = dist.clamp(min=0.3)
    sel = torch.rand(M, generator=gen)
    beyond = sel < 0.05           # ~5 % of returns past ray_range[1]  -> "transparent" rays
    zero = (sel >= 0.05) & (sel < 0.06)  # ~1 % dropped returns     -> not "opaque"
    dist = torch.where(beyond, g["ray_range"][1] + 1.0 + 9.0 * torch.rand(M, generator=gen), dist)
    dist = torch.where(zero, torch.zeros_like(dist), dist)
    ts = torch.linspace(t0, t0 + 0.1, M)
    return Scan(dirs, dist.float(), ts)


def sky_directions(n: int, seed: int) -> torch.Tensor:
    """[3, n] unit directions above the horizon (elevation 25..85 degrees), the shape LidarScan.sky_rays has
    (common/sensors.py:57-82; the tracker's compute_sky_rays produces them from gaps in the scan)."""
    gen = torch.Generator().manual_seed(seed)
    el = torch.deg2rad(25.0 + 60.0 * torch.rand(n, generator=gen, dtype=torch.float64))
    az = 2 * math.pi * torch.rand(n, generator=gen, dtype=torch.float64)
    return torch.stack([torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)], 0).float()


def make_window(geom: str, K: int, seed: int = 0, n_beams: int = 64, n_azimuth: int = 1024):
    """K scans + their poses."""
    poses = keyframe_poses(geom, K)
    scans = [make_scan(geom, poses[k], seed + 17 * k, t0=3.0 * k, n_beams=n_beams, n_azimuth=n_azimuth)
             for k in range(K)]
    return scans, poses


def trained_occupancy_grid(geom: str, V: int = 100) -> torch.Tensor:
    """[1,1,V,V,V] logit grid with +4 in a thin shell around the scene surfaces (grid[z,y,x])."""
    wc = world_cube(geom)
    box_lo, box_hi, spheres = _scene(geom)
    ax = (torch.arange(V, dtype=torch.float32) + 0.5) / V * 2 - 1     # voxel centres in cube units
    zz, yy, xx = torch.meshgrid(ax, ax, ax, indexing="ij")
    shift = torch.tensor(wc.shift)
    wx = xx * wc.scale_factor - shift[0]
    wy = yy * wc.scale_factor - shift[1]
    wz = zz * wc.scale_factor - shift[2]
    vox = 2 * wc.scale_factor / V
    p = torch.stack([wx, wy, wz], -1).double()
    inside = ((p > box_lo) & (p < box_hi)).all(-1)
    dwall = torch.minimum((p - box_lo).abs().min(-1)[0], (p - box_hi).abs().min(-1)[0])
    shell = inside & (dwall < 2 * vox)
    for c, r in spheres:
        dc = (p - torch.tensor(c, dtype=torch.float64)).norm(dim=-1)
        shell |= (dc - r).abs() < 2 * vox
    grid = torch.where(shell, torch.tensor(4.0), torch.tensor(-2.0))
    grid = torch.where(inside | shell, grid, torch.zeros_like(grid))
    return grid.float().reshape(1, 1, V, V, V)
