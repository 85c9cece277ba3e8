"""Hash backward on (a) ray-ordered samples of one keyframe (all rays share the sensor origin) and (b) the same number of
independent uniform positions: the difference is the cost of same-address atomics on the coarse levels (not a test)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import ops

N, S = 8192, 512
dev = "cuda"
net = ops.HashNet()
g = torch.Generator().manual_seed(0)
params = torch.cat([(torch.rand(net.n_network_params, generator=g) - 0.5) * 0.4, (torch.rand(2 * net.table_entries, generator=g) - 0.5)]).to(dev)
packed = ops.hash_pack(net, params)
P = N * S
d = torch.randn(N, 3, generator=g); d = d / d.norm(dim=1, keepdim=True)
z = (torch.rand(N, S, generator=g).sort(dim=1).values * 0.55 + 0.01)
pos_ray = (d[:, None, :] * z[:, :, None]).reshape(-1, 3).contiguous().to(dev)
pos_uni = (torch.rand(P, 3, generator=g) * 1.8 - 0.9).to(dev)
d_sigma = (torch.randn(P, generator=g) * 1e-4).to(dev)
d_params = torch.zeros(net.param_count, device=dev)
scratch = torch.empty(net.bwd_scratch_bytes(P), device=dev, dtype=torch.uint8)

def timeit(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

res = {}
for name, pos in (("ray_ordered_shared_origin", pos_ray), ("independent_uniform", pos_uni), ("ray_shuffled", pos_ray[torch.randperm(P, device=dev)].contiguous())):
    res[name] = {"fwd_ms": round(timeit(lambda: ops.hash_fwd(net, packed, P, pos=pos)), 3),
                 "bwd_ms": round(timeit(lambda: ops.hash_bwd(net, packed, P, d_sigma, 1024.0, d_params, pos=pos, scratch=scratch)), 3)}
print(json.dumps(res))
