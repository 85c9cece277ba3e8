"""Diagnostic (not a test): where does d_sigma of the fused render+loss kernel differ from autograd?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from golden_util import Case
from loner_b200 import ops
from oracle import loner_oracle as orc

for name in sys.argv[1:] or ["c1_2x64_fp16"]:
    c = Case(name)
    r = c.run_oracle()
    rays, depths, res = r["rays"].detach(), r["depths"], r["res"]
    sigma = res["sigma"].detach(); z = res["samples_fine"]; n = rays.shape[0]
    sg = sigma.clone().requires_grad_(True)
    d_, w_, o_, v_ = orc.raw2outputs(sg, z, rays[:, 3:6], c.noise, rays[:, -1:])
    out2 = orc.compute_loss(rays, depths, dict(depth_fine=d_, weights_fine=w_, opacity_fine=o_, variance=v_, samples_fine=z), c.scale, orc.LossCfg())
    out2["loss"].backward()
    far = rays[:, 12]; opaque = (depths > 0) & ~(depths > far)
    flags = (1 + 2 * opaque.to(torch.uint8)).to(torch.uint8).cuda()
    counts = torch.tensor([n, int(opaque.sum())], dtype=torch.int32, device="cuda")
    k = ops.render_loss(sigma.cuda(), z.cuda().contiguous(), rays.cuda().contiguous(), depths.cuda(), flags, counts,
                        [c.scale, 0.5, 1.0, 10.0, 1.0, 1000.0, 0.005], noise=c.noise.cuda().contiguous(), raw_noise_std=1.0)
    got = k["d_sigma"].cpu(); ref = sg.grad
    err = (got - ref).abs()
    print(name, "norm-rel", float((got - ref).norm() / ref.norm()), "ref norm", float(ref.norm()))
    per_ray = (got - ref).pow(2).sum(1).sqrt()
    top = per_ray.topk(5)
    for v, i in zip(top.values, top.indices):
        i = int(i)
        j = int(err[i].argmax())
        print(f"  ray {i} err {float(v):.3e} |ref row| {float(ref[i].norm()):.3e} opaque {bool(opaque[i])} depth {float(depths[i]):.4f} far {float(far[i]):.4f} "
              f"worst s={j} ref {float(ref[i,j]):.4e} got {float(got[i,j]):.4e} w {float(w_[i,j]):.3e} A {float(o_[i]):.4f} eps {float(out2['eps_dynamic'][i]):.3f}")
    print("  err by opaque:", float(per_ray[opaque].pow(2).sum().sqrt()), "non-opaque:", float(per_ray[~opaque].pow(2).sum().sqrt()))
    last = (got[:, -1] - ref[:, -1]).norm(); print("  last-sample err", float(last), " all-but-last", float((got[:, :-1]-ref[:, :-1]).norm()))
