"""Timing of the hash-grid sigma head at the C2 size (not a test): CUDA events, 5 reps after 2 warm-ups."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import ops, synth

N, S = int(os.environ.get("MB_N", 8192)), 512
dev = "cuda"
net = ops.HashNet(flags=int(os.environ.get("MB_HASH_FLAGS", 0)))
g = torch.Generator().manual_seed(0)
params = torch.cat([(torch.rand(net.n_network_params, generator=g) - 0.5) * 0.4, (torch.rand(2 * net.table_entries, generator=g) - 0.5)]).to(dev)
packed = ops.hash_pack(net, params)
wc = synth.world_cube("canteen")
rays = torch.zeros(N, 13)
d = torch.randn(N, 3, generator=g); d = d / d.norm(dim=1, keepdim=True)
rays[:, 3:6] = d; rays[:, 11] = 1 / wc.scale_factor; rays[:, 12] = 50 / wc.scale_factor
rays = rays.to(dev)
z = (torch.rand(N, S, generator=g).sort(dim=1).values * 0.55 + 0.01).to(dev)
P = N * S
d_sigma = (torch.randn(P, generator=g) * 1e-4).to(dev)
d_params = torch.zeros(net.param_count, device=dev)
scratch = torch.empty(net.bwd_scratch_bytes(P), device=dev, dtype=torch.uint8)
sigma = torch.empty(P, device=dev)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = {"hash_fwd_ms": timeit(lambda: ops.hash_fwd(net, packed, P, rays=rays, z=z, sigma=sigma)),
       "hash_bwd_ms": timeit(lambda: ops.hash_bwd(net, packed, P, d_sigma, 1024.0, d_params, rays=rays, z=z, scratch=scratch)),
       "hash_bwd_dpos_ms": timeit(lambda: ops.hash_bwd(net, packed, P, d_sigma, 1024.0, d_params, rays=rays, z=z, want_dpos=True, scratch=scratch)),
       "hash_pack_ms": timeit(lambda: ops.hash_pack(net, params, packed)),
       "samples": P, "gathers_per_sample": 128}
res["fwd_gather_GBps"] = round(P * 128 * 4 / res["hash_fwd_ms"] / 1e6, 1)
print(json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in res.items()}))
