"""Per-kernel timings at the C2 size (not a test): CUDA events, 5 reps after 2 warm-ups."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import ops, synth, engine as eng

N, S, W, L = int(os.environ.get("MB_N", 8192)), 512, 256, 4
dev = "cuda"
FLAGS = int(os.environ.get("MB_FLAGS", ops.DEFAULT_NET_FLAGS))      # LONER_NET_* kernel variants
net = ops.Net(10, W, L, flags=FLAGS)
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

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = {}
z = ops.sample_ogm(rays, grid, S, 1.0, None, None, seed=1)
res["sample_ogm"] = timeit(lambda: ops.sample_ogm(rays, grid, S, 1.0, None, None, seed=1))
acts = torch.empty(net.act_bytes(P), device=dev, dtype=torch.uint8)
sigma = torch.empty(P, device=dev)
res["mlp_fwd_stash"] = timeit(lambda: ops.mlp_fwd(net, packed, P, rays=rays, z=z, stash=True, sigma=sigma, acts=acts))
if os.environ.get("MB_ONLY") == "fwd":
    print(json.dumps({k: round(v, 4) for k, v in res.items()})); sys.exit(0)
res["mlp_fwd_infer"] = timeit(lambda: ops.mlp_fwd(net, packed, P, rays=rays, z=z, stash=False, sigma=sigma))
depths = torch.full((N,), 0.3, device=dev)
flags = torch.full((N,), 3, dtype=torch.uint8, device=dev)
counts = torch.tensor([N, N], dtype=torch.int32, device=dev)
cfg7 = [wc.scale_factor, 0.5, 1.0, 10.0, 1.0, 1000.0, 0.005]
rl = ops.render_loss(sigma, z, rays, depths, flags, counts, cfg7, want_outputs=False)
res["render_loss"] = timeit(lambda: ops.render_loss(sigma, z, rays, depths, flags, counts, cfg7, want_outputs=False))
scratch = torch.empty(net.bwd_scratch_bytes(P), device=dev, dtype=torch.uint8)
gs = ops.default_grad_scale(N, S)
res["mlp_dgrad"] = timeit(lambda: ops.mlp_dgrad(net, packed, P, rl["d_sigma"], acts, gs, scratch, rays=rays, z=z))
res["mlp_dgrad_dpos"] = timeit(lambda: ops.mlp_dgrad(net, packed, P, rl["d_sigma"], acts, gs, scratch, rays=rays, z=z, want_dpos=True))
dp = torch.zeros(net.param_count, device=dev)
res["mlp_wgrad_all"] = timeit(lambda: ops.mlp_wgrad(net, packed, P, rl["d_sigma"], acts, gs, dp, scratch))
f_fwd = 2 * (64 * W + (L - 1) * W * W + W)
print(json.dumps(dict(flags=FLAGS, **{k: round(v, 4) for k, v in res.items()})))
if os.environ.get("MB_SHORT"): sys.exit(0)
print("fwd TF/s stash %.0f infer %.0f | stash GB %.2f dz GB %.2f" % (
    f_fwd * P / res["mlp_fwd_stash"] / 1e9, f_fwd * P / res["mlp_fwd_infer"] / 1e9,
    net.act_bytes(P) / 1e9, (net.bwd_scratch_bytes(P)) / 1e9))
buf = torch.empty(10 * 1024**3 // 4, device=dev, dtype=torch.float32)
t = timeit(lambda: buf.fill_(1.0))
print("pure write (fill_) 10 GiB: %.3f ms -> %.2f TB/s" % (t, 10 * 1024**3 / t / 1e9))
src = torch.empty(5 * 1024**3 // 4, device=dev, dtype=torch.float32)
t = timeit(lambda: buf[: src.numel()].copy_(src))
print("copy 5 GiB -> 5 GiB: %.3f ms -> %.2f TB/s (read+write)" % (t, 10 * 1024**3 / t / 1e9))
