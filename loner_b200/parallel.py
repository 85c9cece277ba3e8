"""Ray-sharded data parallelism: host-side logic of the two exchanges per step.

Every rank holds all keyframes and the full (replicated) parameters / optimiser state; rank r draws
its own rays.  Exactness w.r.t. the single-GPU loss needs GLOBAL normalisers, so the step does
  (1) all-reduce of the two loss normalisers (#valid, #opaque rays)   -- before the loss kernel
  (2) in-place all-reduce of ONE flat buffer  [MLP grads | pose grads | 4 loss sums]  -- after backward
The same code runs over NCCL (GPU box) and gloo (CPU tests).
"""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank():
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_slice(n_total: int, r: int, world: int):
    """Contiguous slice [lo, hi) of n_total rays owned by rank r (remainder spread over the first ranks)."""
    base, rem = divmod(n_total, world)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


def allreduce_counts(counters: torch.Tensor):
    """int32[2] (#valid, #opaque) -> global sums, in place."""
    if world_size() > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM)
    return counters


class FlatGrads:
    """The step's exchange buffer  [MLP grads | pose grads [K,12] | 4 loss sums]  as ONE flat fp32 tensor with three
    views.  The kernels accumulate straight into the views, so the multi-GPU exchange is a single in-place
    all-reduce with no staging copies (round 1 packed, reduced and unpacked: three extra copies per step)."""

    def __init__(self, n_params: int, K: int, device):
        self.n_params, self.K = int(n_params), int(K)
        self.flat = torch.zeros(self.n_params + 12 * self.K + 4, device=device, dtype=torch.float32)
        self.d_params = self.flat[:self.n_params]
        self.d_poses12 = self.flat[self.n_params:self.n_params + 12 * self.K].view(self.K, 12)
        self.loss_acc = self.flat[self.n_params + 12 * self.K:]

    def zero_(self):
        self.flat.zero_()

    def allreduce(self):
        """Sums the whole buffer over the ranks, in place."""
        if world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        return self


def grad_check(make_engine, window, n_per_kf, S, optimize_poses, seed=11):
    """the k-GPU step equals the 1-GPU step on ONE global ray set.
    make_engine(distributed) -> a MappingEngine with its keyframes added (identical on every rank).
    Every rank runs (a) the sharded step - rank r renders slice r of the global rays, normalisers and the flat
    gradient buffer are all-reduced - and (b) the whole global set on a non-distributed engine; returns the
    relative differences of loss, MLP gradient and pose gradient (max over ranks)."""
    world, r = world_size(), rank()
    e_ref, e_dp = make_engine(False), make_engine(True)
    K = len(window)
    n_global = n_per_kf * world
    g = torch.Generator().manual_seed(seed)
    # global ray set: per keyframe n_global picks; rank r owns columns [r*n_per_kf, (r+1)*n_per_kf) of every keyframe
    idx = torch.stack([torch.randint(0, e_ref.kf_sizes[k], (n_global,), generator=g) + e_ref.kf_offsets[k] for k in window])
    H = S // 2
    u1 = torch.rand(K, n_global, H, generator=g)
    u2 = torch.rand(K, n_global, H, generator=g)
    noise = torch.randn(K, n_global, S, generator=g)
    sl = slice(r * n_per_kf, (r + 1) * n_per_kf)

    def inj(cols):
        return dict(ray_point=idx[:, cols].reshape(-1), u1=u1[:, cols].reshape(-1, H), u2=u2[:, cols].reshape(-1, H),
                    noise=noise[:, cols].reshape(-1, S))

    out = {}
    for name, e, cols, n in (("ref", e_ref, slice(0, n_global), n_global), ("dp", e_dp, sl, n_per_kf)):
        e.new_phase(optimize_poses=optimize_poses)
        loss = e.step(window, n, optimize_poses=optimize_poses, injected=inj(cols))
        poses = torch.stack([p.grad if p.grad is not None else torch.zeros_like(p) for p in e.poses6])
        out[name] = (loss.detach().double(), e.d_params.detach().double().clone(), poses.double())
    (l0, g0, p0), (l1, g1, p1) = out["ref"], out["dp"]
    errs = torch.stack([(l0 - l1).abs() / l0.abs(), (g0 - g1).norm() / g0.norm(),
                        (p0 - p1).norm() / (p0.norm() + 1e-30)])
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
    return dict(loss_rel=float(errs[0]), d_params_rel=float(errs[1]), pose_grad_rel=float(errs[2]), world=world)
