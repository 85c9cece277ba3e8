"""Multi-GPU gradient check (not collected by pytest; launched by tests/test_gpu_round2.py and usable by hand):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29631 \
        tests/mgpu_grad_check.py

Every rank builds the same keyframes; the k-rank sharded step (NCCL all-reduce of the loss normalisers and of the
flat [MLP grads | pose grads | loss sums] buffer) must reproduce the single-GPU step on the same global ray set
(SURVEY.md section 4 item 5).  Rank 0 prints one JSON line."""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loner_b200 import engine as eng, parallel, synth  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    K, S, W, L = 3, 128, 128, 2
    wc = synth.world_cube("garden")
    scans, poses = synth.make_window("garden", K, seed=3, n_beams=16, n_azimuth=256)
    grid = synth.trained_occupancy_grid("garden")[0, 0]

    def make(distributed):
        cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_neurons=W, n_hidden_layers=L,
                               n_samples=S, chunk_rays=256)
        e = eng.MappingEngine(cfg, device=torch.device("cuda", local), distributed=distributed)
        for k in range(K):
            e.add_keyframe(scans[k].ray_directions, scans[k].distances, synth.axis_angle_from_yaw_pose(poses[k]))
        e.grid.copy_(grid)
        return e

    r = parallel.grad_check(make, list(range(K)), 160, S, optimize_poses=True)
    if dist.get_rank() == 0:
        print(json.dumps(r))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
