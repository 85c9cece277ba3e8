"""models.model_tcnn of the reference (/root/reference/src/models/model_tcnn.py): Model and
OccupancyGridModel with the same constructor arguments, methods and state_dict surface."""
from collections import defaultdict

import torch
import torch.nn as nn

from models.nerf_tcnn import DecoupledNeRF
from models.rendering_tcnn import inference, render_rays


class Model(nn.Module):
    def __init__(self, cfg):
        super(Model, self).__init__()
        self.cfg = cfg
        if cfg.model_type == 'nerf_decoupled':
            self.nerf_model = DecoupledNeRF(cfg.nerf_config, cfg.num_colors)
        else:
            raise NotImplementedError()

    def get_rgb_parameters(self, ignore_requires_grad=False):
        mods = [self.nerf_model._model_intensity, self.nerf_model._pos_encoding, self.nerf_model._dir_encoding]
        ps = [p for m in mods if m is not None for p in m.parameters()]
        return ps if ignore_requires_grad else [p for p in ps if p.requires_grad]

    def get_rgb_mlp_parameters(self):
        return list(self.nerf_model._model_intensity.parameters())

    def get_rgb_feature_parameters(self):
        mods = [self.nerf_model._pos_encoding, self.nerf_model._dir_encoding]
        return [p for m in mods if m is not None for p in m.parameters() if p.requires_grad]

    def get_sigma_parameters(self, ignore_requires_grad=False):
        ps = list(self.nerf_model._model_sigma.parameters())
        return ps if ignore_requires_grad else [p for p in ps if p.requires_grad]

    def freeze_sigma_head(self, should_freeze=True):
        for p in self.get_sigma_parameters(True):
            p.requires_grad = not should_freeze

    def freeze_rgb_head(self, should_freeze=True):
        for p in self.get_rgb_parameters(True):
            p.requires_grad = not should_freeze

    def inference_points(self, xyz_, dir_, sigma_only):
        return inference(self.nerf_model, xyz_, dir_, netchunk=0, sigma_only=sigma_only, meshing=True)

    def forward(self, rays, ray_sampler, scale_factor, testing=False, camera=True, detach_sigma=True,
                return_variance=False):
        """model_tcnn.py:70-105: chunked render_rays, results concatenated."""
        r = self.cfg.render
        n_samples, perturb = (r.N_samples_test, 0.) if testing else (r.N_samples_train, r.perturb)
        results = defaultdict(list)
        for i in range(0, rays.shape[0], r.chunk):
            out = render_rays(rays[i:i + r.chunk, :], ray_sampler, self.nerf_model, self.cfg.ray_range, scale_factor,
                              N_samples=n_samples, retraw=r.retraw, perturb=perturb, white_bkgd=r.white_bkgd,
                              raw_noise_std=r.raw_noise_std, netchunk=r.netchunk, num_colors=self.cfg.num_colors,
                              sigma_only=(not camera), detach_sigma=detach_sigma, return_variance=return_variance)
            for k, v in out.items():
                results[k] += [v]
        return {k: torch.cat(v, 0) for k, v in results.items()}


class OccupancyGridModel(nn.Module):
    def __init__(self, cfg):
        super(OccupancyGridModel, self).__init__()
        self.cfg = cfg
        v = cfg.voxel_size
        self.occupancy_grid = nn.Parameter(torch.zeros(1, 1, v, v, v))

    def forward(self):
        return self.occupancy_grid

    @staticmethod
    def interpolate(occupancy_grid, ray_bin_centers, mode='bilinear'):
        """Trilinear lookup with autograd to the grid (model_tcnn.py:124-131).  Inside the sampler the
        fused kernel does this lookup itself; this entry point serves Optimizer._step_occupancy_grid."""
        n_rays, n_bins, _ = ray_bin_centers.shape
        g = ray_bin_centers.reshape(1, 1, n_rays, n_bins, 3)
        return nn.functional.grid_sample(occupancy_grid, g, mode=mode, align_corners=False).reshape(n_rays, n_bins)
