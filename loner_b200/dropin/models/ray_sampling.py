"""models.ray_sampling of the reference (/root/reference/src/models/ray_sampling.py) on the fused
sampler kernels (loner_sample_uniform / loner_sample_ogm)."""
import torch

from loner_b200 import ops


def _inj(sampler, key, dev):
    t = (getattr(sampler, "injected", None) or {}).get(key)
    return None if t is None else t.to(dev).contiguous()


class UniformRaySampler():
    def __init__(self):
        self._calls = 0
        self.seed = 0
        self.injected = None        # parity hook: dict(u1=[N,S]) replays the reference's torch.rand draws

    def get_samples(self, rays, N_samples, perturb):
        self._calls += 1
        with torch.no_grad():
            return ops.sample_uniform(rays.detach().contiguous().float(), N_samples, perturb, _inj(self, "u1", rays.device),
                                      seed=self.seed * 1000003 + self._calls)


class OccGridRaySampler():
    def __init__(self):
        self._occ_gamma = None
        self._calls = 0
        self.seed = 0
        self.injected = None        # parity hook: dict(u1=[N,S/2], u2=[N,S/2], noise=[N,S])

    def update_occ_grid(self, occ_gamma):
        self._occ_gamma = occ_gamma

    def get_samples(self, rays, N_samples, perturb):
        self._calls += 1
        g = self._occ_gamma
        grid = g.detach().reshape(g.shape[-3:]).contiguous().float()
        with torch.no_grad():
            return ops.sample_ogm(rays.detach().contiguous().float(), grid, N_samples, perturb,
                                  _inj(self, "u1", rays.device) if perturb > 0 else None, _inj(self, "u2", rays.device),
                                  seed=self.seed * 1000003 + self._calls)
