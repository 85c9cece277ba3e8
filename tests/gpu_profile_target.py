"""ncu target (not a test): one launch each of fwd-infer, fwd-stash, dgrad, wgrad after one warm-up.
    ncu --set full --import-source on -k regex:mlp_ --launch-skip 5 -c 5 python tests/gpu_profile_target.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import ops, synth, engine as eng

N, S, W, L = int(os.environ.get("MB_N", 2048)), 512, 256, 4
dev = "cuda"
net = ops.Net(10, W, L)
params = eng.xavier_uniform_flat(net.layer_shapes(), 1337).to(dev)
packed = ops.mlp_pack(net, params)
wc = synth.world_cube("canteen")
g = torch.Generator().manual_seed(0)
rays = torch.zeros(N, 13)
d = torch.randn(N, 3, generator=g); d = d / d.norm(dim=1, keepdim=True)
rays[:, 3:6] = d; rays[:, 6:9] = -d; rays[:, 11] = 1 / wc.scale_factor; rays[:, 12] = 50 / wc.scale_factor
rays = rays.to(dev)
grid = synth.trained_occupancy_grid("canteen")[0, 0].to(dev)
P = N * S
z = ops.sample_ogm(rays, grid, S, 1.0, None, None, seed=1)
acts = torch.empty(net.act_bytes(P), device=dev, dtype=torch.uint8)
sigma = torch.empty(P, device=dev)
depths = torch.full((N,), 0.3, device=dev)
flags = torch.full((N,), 3, dtype=torch.uint8, device=dev)
counts = torch.tensor([N, N], dtype=torch.int32, device=dev)
cfg7 = [wc.scale_factor, 0.5, 1.0, 10.0, 1.0, 1000.0, 0.005]
scratch = torch.empty(net.bwd_scratch_bytes(P), device=dev, dtype=torch.uint8)
gs = ops.default_grad_scale(N, S)
dp = torch.zeros(net.param_count, device=dev)
for rep in range(2):      # rep 0 = warm-up (5 mlp_* launches: infer, stash, dgrad, wgrad, wgrad_reduce)
    ops.mlp_fwd(net, packed, P, rays=rays, z=z, stash=False, sigma=sigma)
    ops.mlp_fwd(net, packed, P, rays=rays, z=z, stash=True, sigma=sigma, acts=acts)
    rl = ops.render_loss(sigma, z, rays, depths, flags, counts, cfg7, want_outputs=False)
    ops.mlp_dgrad(net, packed, P, rl["d_sigma"], acts, gs, scratch, rays=rays, z=z)
    ops.mlp_wgrad(net, packed, P, rl["d_sigma"], acts, gs, dp, scratch)
    torch.cuda.synchronize()
print("done")
