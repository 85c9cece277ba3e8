"""ncu target (not a test): every kernel of the library once, at the C2 size (MB_N rays x 512 samples), after warm-up.
One joint pose+map iteration of the Frequency 4x256 engine (occupancy update included), one test-mode render, and one
iteration of the shipped HashGrid + 1x64 engine run between cudaProfilerStart/Stop:

    ncu --set full --import-source on --profile-from-start off -k regex:'^(?!.*(at::|elementwise|vectorized|reduce_kernel<|cub::))' \
        -o gpurun_out/prof python tests/gpu_profile_target.py
MB_FLAGS selects the MLP kernel variants (loner_net_t.flags)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import ops, synth, engine as eng

N, S = int(os.environ.get("MB_N", 8192)), 512
flags = os.environ.get("MB_FLAGS")
wc = synth.world_cube("canteen")
scans, poses = synth.make_window("canteen", 2, seed=0)


def engine(**kw):
    cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=(1.0, 50.0), n_samples=S,
                           net_flags=None if flags is None else int(flags), **kw)
    e = eng.MappingEngine(cfg)
    for k in range(2):
        e.add_keyframe(scans[k].ray_directions, scans[k].distances, synth.axis_angle_from_yaw_pose(poses[k]))
    e.grid.copy_(synth.trained_occupancy_grid("canteen")[0, 0])
    e.new_phase(optimize_poses=True)
    return e


ef = engine()
eh = engine(encoding="HashGrid", n_neurons=64, n_hidden_layers=1)
for e in (ef, eh):
    for _ in range(10):                       # global_step 10 -> the profiled step runs the occupancy update too
        e.step([0, 1], N // 2, optimize_poses=True)
rays = ef.last["rays"]
ef.render(rays, seed=1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ef.step([0, 1], N // 2, optimize_poses=True)
ef.step([0, 1], N // 2, optimize_poses=False)      # the map-only dgrad variant (no d_pos)
ef.render(rays, seed=2)
eh.step([0, 1], N // 2, optimize_poses=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
