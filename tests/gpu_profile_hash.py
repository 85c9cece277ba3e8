"""ncu target (not a test): the hash-grid sigma head at MB_N rays x 512 samples on the synthetic canteen scan.
    ncu --set full --import-source on -k regex:hash_ -s 3 -c 3 python tests/gpu_profile_hash.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import ops, synth, engine as eng

N, S = int(os.environ.get("MB_N", 8192)), 512
dev = "cuda"
wc = synth.world_cube("canteen")
cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_neurons=64, n_hidden_layers=1,
                       n_samples=S, encoding="HashGrid")
e = eng.MappingEngine(cfg)
scans, poses = synth.make_window("canteen", 1, seed=0)
e.add_keyframe(scans[0].ray_directions, scans[0].distances, synth.axis_angle_from_yaw_pose(poses[0]))
e.grid.copy_(synth.trained_occupancy_grid("canteen")[0, 0])
e.new_phase(False)
for _ in range(2):
    e.step([0], N)
torch.cuda.synchronize()
print("done")
