"""GPU parity of the WHOLE mapping iteration (ray build -> sample -> MLP -> render -> JS loss ->
backward) against (a) fixtures minted from the reference's own Python and (b) the oracle's
autograd, on identical keyframes, ray indices and injected random numbers."""
import pytest
import torch

from golden_util import Case
from gpu_util import norm_relerr, relerr
from loner_b200 import engine as eng
from oracle import loner_oracle as orc

pytestmark = pytest.mark.gpu


def _engine_for(c: Case, sampler="OGM"):
    cfg = eng.EngineConfig(scale=c.scale, shift=tuple(c.shift.tolist()), ray_range=c.ray_range, n_frequencies=10,
                           n_neurons=c.W, n_hidden_layers=c.L, n_samples=c.S, sampler=sampler,
                           loss_selection=c.loss_selection)
    e = eng.MappingEngine(cfg, params=c.params)
    e.grid.copy_(c.grid[0, 0])
    for k in range(c.K):
        e.add_keyframe(c.scans[k].ray_directions, c.scans[k].distances, c.poses6[k])
    return e


def _run(c: Case):
    e = _engine_for(c)
    e.new_phase(optimize_poses=c.pose_grads)
    params0 = e.params.clone()
    ray_point = torch.cat([c.idx[k] + e.kf_offsets[k] for k in range(c.K)])
    inj = dict(ray_point=ray_point, u1=c.u1, u2=c.u2, noise=c.noise)
    loss = e.step(list(range(c.K)), c.n, optimize_poses=c.pose_grads, injected=inj, want_outputs=True)
    torch.cuda.synchronize()
    return e, loss, params0


@pytest.mark.parametrize("name", ["kf2_4x256_fp16", "quad_4x256_fp16", "kf2_2x128_fp16", "kf2_2x128_l2js",
                                  "kf2_2x128_l1los"])
def test_step_matches_reference_fixture(name):
    c = Case(name)
    e, loss, params0 = _run(c)
    g = c.g
    o = e.last["outs"][0]
    rows = g["z_vals"].shape[0]
    errs = dict(
        z=float((o["z_vals"][:rows].cpu() - torch.from_numpy(g["z_vals"])).abs().max()),
        depth=relerr(o["depth"], g["depth_fine"]), opacity=relerr(o["opacity"], g["opacity_fine"]),
        variance=relerr(o["variance"], g["variance"]), weights=relerr(o["weights"][:rows], g["weights"]),
        loss=abs(float(loss) - float(g["loss"])) / float(g["loss"]),
        depth_eps=abs(float(e.last["depth_eps"]) - float(g["depth_eps"])) / float(g["depth_eps"]))
    print(f"[{name}] " + " ".join(f"{k} {v:.2e}" for k, v in errs.items()))
    assert errs["z"] < 1e-5
    for k in ("depth", "opacity", "variance", "loss", "depth_eps"):
        assert errs[k] < 1e-4, k                  # north_star: depths and losses within 1e-4 rel
    assert errs["weights"] < 2e-4
    # gradients: oracle autograd (fp32 backward through the fp16-rounded forward)
    r = c.run_oracle()
    gp = e.d_params.cpu()
    en = norm_relerr(gp, r["params"].grad)
    cos = float(torch.dot(gp, r["params"].grad) / (gp.norm() * r["params"].grad.norm()))
    print(f"[{name}] d_params norm-rel {en:.2e} cosine {cos:.6f} |g| {float(gp.norm()):.3e} vs fixture "
          f"{float(g['grad_params_norm']):.3e}")
    assert en < 2e-3 and cos > 0.99999          # measured (B200): 0.9e-4 .. 4.1e-4
    if c.pose_grads:
        mine = torch.stack([p.grad.cpu() if p.grad is not None else torch.zeros(6) for p in e.poses6])
        ep = norm_relerr(mine, g["grad_poses"])
        print(f"[{name}] pose grads norm-rel vs reference fixture {ep:.2e}")
        # fp16 input-gradient chain (like tcnn's) summed over N*S samples with heavy cancellation: measured (B200)
        # 0.5e-2 .. 2.7e-2 on these fixtures (the oracle's own fp16-vs-fp32 difference is of the same size)
        assert ep < 5e-2
    # Adam moved the parameters exactly as torch.optim.Adam would with these gradients
    p_ref, _, _ = orc.adam_update(params0.cpu(), gp, torch.zeros_like(gp), torch.zeros_like(gp), 1, 0.01)
    assert relerr(e.params, p_ref) < 1e-6


def test_fp16_tensor_core_path_vs_pure_fp32_reference():
    """Documented deviation: the reference's tcnn runs fp16; against a PURE fp32 network the fp16
    tensor-core path is expected to sit at the 1e-4..1e-3 level on depth (reported, loosely bounded)."""
    c16, c32 = Case("kf2_4x256_fp16"), Case("kf2_4x256_fp32")
    e, loss, _ = _run(c16)
    o = e.last["outs"][0]
    ed = relerr(o["depth"], c32.g["depth_fine"])
    el = abs(float(loss) - float(c32.g["loss"])) / float(c32.g["loss"])
    print(f"fp16 path vs pure-fp32 reference fixture: depth rel {ed:.2e} loss rel {el:.2e}")
    assert ed < 5e-3 and el < 5e-3


def test_uniform_sampler_step_and_multi_chunk_equivalence():
    c = Case("kf2_4x256_fp16")
    e1 = _engine_for(c)
    e2 = _engine_for(c)
    e2.cfg.chunk_rays = 96        # forces 3 ragged chunks over the 256 rays
    ray_point = torch.cat([c.idx[k] + e1.kf_offsets[k] for k in range(c.K)])
    inj = dict(ray_point=ray_point, u1=c.u1, u2=c.u2, noise=c.noise)
    l1 = e1.step(list(range(c.K)), c.n, injected=inj)
    l2 = e2.step(list(range(c.K)), c.n, injected=inj)
    assert abs(float(l1) - float(l2)) / float(l1) < 1e-5
    assert norm_relerr(e2.d_params, e1.d_params) < 1e-3
