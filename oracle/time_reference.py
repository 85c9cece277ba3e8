"""TEST INFRASTRUCTURE — times the REFERENCE'S OWN `Optimizer.iterate_optimizer` (CPU, /root/reference, stub-imported by
oracle/ref_harness.py) next to the oracle port that bench.py uses as its CPU arm, on the same machine and workload:

    python -m oracle.time_reference [rays] [iterations]

This validates `cpu_baseline.kind = "port"` (missing item 6): the port must not be slower than the code it
stands for.  Build-container only (needs /root/reference); the result is recorded in BASELINE.md."""
import json
import os
import sys
import tempfile
import time

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, REPO)

from loner_b200 import synth  # noqa: E402
from oracle import make_golden as mg  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from oracle import tcnn_standin  # noqa: E402


def time_reference(n_rays=512, iters=3, S=512, W=256, L=4, geom="canteen"):
    ns = rh.import_reference()
    tcnn_standin.PRECISION["mode"] = "fp32"
    torch.manual_seed(0)
    with tempfile.TemporaryDirectory() as tmp:
        opt, wc, settings = mg.build_reference_optimizer(ns, geom, S, L, W, tmp)
        opt._settings["num_samples"]["lidar"] = n_rays
        opt._num_lidar_samples = n_rays
        scans, poses = synth.make_window(geom, 1, seed=0)
        grid0 = synth.trained_occupancy_grid(geom)
        with torch.no_grad():
            opt._occupancy_grid_model.occupancy_grid.copy_(grid0)
        opt._occupancy_grid = opt._occupancy_grid_model()
        opt._ray_sampler.update_occ_grid(opt._occupancy_grid.detach())
        sc = ns.sensors.LidarScan(scans[0].ray_directions.clone(), scans[0].distances.clone(), scans[0].timestamps.clone())
        fr = ns.frame.Frame(None, sc, None)
        fr._lidar_pose = ns.pose.Pose(pose_tensor=synth.axis_angle_from_yaw_pose(poses[0]).clone(), fixed=True)
        fr._gt_lidar_pose = fr._lidar_pose
        kf = ns.keyframe.KeyFrame(fr, "cpu")
        os_ = ns.optimizer.OptimizationSettings(num_iterations=1, freeze_poses=True)
        opt.iterate_optimizer([kf], os_)                      # warm-up (1 iteration)
        os_ = ns.optimizer.OptimizationSettings(num_iterations=iters, freeze_poses=True)
        t0 = time.perf_counter()
        opt.iterate_optimizer([kf], os_)
        dt = (time.perf_counter() - t0) / iters
    return n_rays / dt, dt


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    sys.path.insert(0, REPO)
    import bench
    out = {"rays": n, "threads": {}}
    for th in (8, os.cpu_count() or 8):
        torch.set_num_threads(th)
        ref_rps, ref_dt = time_reference(n, iters)
        port_rps, port_dt = bench.cpu_port_rays_per_sec(bench.WORKLOADS["c2"], n, iters, 1, tune=False)
        out["threads"][th] = {"reference_rays_per_s": round(ref_rps, 1), "reference_s_per_iteration": round(ref_dt, 3),
                              "port_rays_per_s": round(port_rps, 1), "port_s_per_iteration": round(port_dt, 3),
                              "port_over_reference": round(port_rps / ref_rps, 3)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
