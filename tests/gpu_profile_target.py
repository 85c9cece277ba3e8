"""ncu target (not a test): every kernel of the step once after one warm-up pass, at MB_N rays x 512 samples, 4x256.
    ncu --set full --import-source on -k regex:mlp_ --launch-skip <n> -c <m> python tests/gpu_profile_target.py
MB_FLAGS selects the kernel variants (loner_net_t.flags), MB_ONLY=wgrad runs only forward + backward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import ops, synth, engine as eng

N, S, W, L = int(os.environ.get("MB_N", 2048)), 512, 256, 4
FLAGS = int(os.environ.get("MB_FLAGS", ops.DEFAULT_NET_FLAGS))
dev = "cuda"
net = ops.Net(10, W, L, flags=FLAGS)
params = eng.xavier_uniform_flat(net.layer_shapes(), 1337).to(dev)
packed = ops.mlp_pack(net, params)
wc = synth.world_cube("canteen")
scans, poses = synth.make_window("canteen", 1, seed=0)
points = ops.pack_points(scans[0].ray_directions, scans[0].distances).to(dev)
g = torch.Generator().manual_seed(0)
ray_point = torch.randint(0, points.shape[0], (N,), generator=g).to(dev)
ray_kf = torch.zeros(N, dtype=torch.int32, device=dev)
P6 = synth.axis_angle_from_yaw_pose(poses[0])
poses12 = eng.poses6_to_poses12(P6[None]).to(dev)
grid = synth.trained_occupancy_grid("canteen")[0, 0].to(dev)
P = N * S
acts = torch.empty(net.act_bytes(P), device=dev, dtype=torch.uint8)
sigma = torch.empty(P, device=dev)
cfg7 = [wc.scale_factor, 0.5, 1.0, 10.0, 1.0, 1000.0, 0.005]
scratch = torch.empty(net.bwd_scratch_bytes(P), device=dev, dtype=torch.uint8)
gs = ops.default_grad_scale(N, S)
dp = torch.zeros(net.param_count, device=dev)
m, v = torch.zeros_like(dp), torch.zeros_like(dp)
params2 = params.clone()
only = os.environ.get("MB_ONLY", "")
for rep in range(2):      # rep 0 = warm-up
    counters = torch.zeros(2, dtype=torch.int32, device=dev)
    rays, depths, flags = ops.ray_build(points, ray_kf, ray_point, poses12, wc.shift, wc.scale_factor, (1.0, 50.0), counters)
    z = ops.sample_ogm(rays, grid, S, 1.0, None, None, seed=1)
    if not only:
        ops.mlp_fwd(net, packed, P, rays=rays, z=z, stash=False, sigma=sigma)
    ops.mlp_fwd(net, packed, P, rays=rays, z=z, stash=True, sigma=sigma, acts=acts)
    rl = ops.render_loss(sigma, z, rays, depths, flags, counters, cfg7, want_outputs=False)
    if not only:
        ops.mlp_dgrad(net, packed, P, rl["d_sigma"], acts, gs, scratch, rays=rays, z=z, want_dpos=True)
    ops.mlp_dgrad(net, packed, P, rl["d_sigma"], acts, gs, scratch, rays=rays, z=z)
    ops.mlp_wgrad(net, packed, P, rl["d_sigma"], acts, gs, dp, scratch)
    if not only:
        ops.adam_step(params2, dp, m, v, 1, 0.01)
        ops.mlp_pack(net, params)
        ops.ogm_grad(rays, z, depths, wc.scale_factor, 100, flags=flags)
        ops.render_fwd(sigma.view(N, S), z, rays, raw_noise_std=1.0, seed=3)
    torch.cuda.synchronize()
print("done")
