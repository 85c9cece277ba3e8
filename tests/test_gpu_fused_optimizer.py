"""GPU: FusedOptimizer follows the reference Optimizer's contract (keyframe schedule, phases, pose
hand-back, attributes the mapper reads) on duck-typed keyframes."""
import types

import pytest
import torch

from loner_b200 import synth
from loner_b200.dropin.mapping_optimizer import FusedOptimizer, OptimizationSettings

pytestmark = pytest.mark.gpu


class Cfg(dict):
    def __getattr__(self, k):
        v = self[k]
        return Cfg(v) if isinstance(v, dict) else v


class _Pose:
    def __init__(self, p6):
        self._t = p6.clone()

    def get_pose_tensor(self):
        return self._t


class _KF:
    def __init__(self, scan, p6, t):
        self._scan, self._pose, self._time, self.is_anchored = scan, _Pose(p6), t, False

    def get_lidar_scan(self):
        return self._scan

    def get_lidar_pose(self):
        return self._pose

    def get_time(self):
        return torch.tensor(self._time)


def _settings(n_first=30, n_joint=20):
    """The keys Optimizer reads from cfg/defaults.yaml + default_model_config.yaml, with a Frequency sigma head."""
    return Cfg(
        freeze_poses=False, skip_pose_refinement=True, num_samples=dict(lidar=256, sky=64),
        rays_selection=dict(strategy="RANDOM"), samples_selection=dict(strategy="OGM"),
        keyframe_schedule=(
            dict(num_keyframes=1, iteration_schedule=(dict(num_iterations=n_first, freeze_poses=True,
                                                           freeze_sigma_mlp=False, freeze_rgb_mlp=True),)),
            dict(num_keyframes=-1, iteration_schedule=(
                dict(num_iterations=50, freeze_poses=False, latest_kf_only=True, freeze_sigma_mlp=True, freeze_rgb_mlp=True),
                dict(num_iterations=n_joint, freeze_poses=False, freeze_sigma_mlp=False, freeze_rgb_mlp=True)))),
        model_config=dict(
            model=dict(ray_range=[1, 50], num_colors=3, model_type="nerf_decoupled",
                       nerf_config=dict(pos_encoding_sigma=dict(otype="Frequency", n_frequencies=10),
                                        sigma_network=dict(otype="CutlassMLP", n_neurons=128, n_hidden_layers=2)),
                       render=dict(N_samples_train=128, N_samples_test=256, perturb=1.0, raw_noise_std=1.0, chunk=16384,
                                   netchunk=0, retraw=True, white_bkgd=False),
                       occ_model=dict(voxel_size=100, lr=1e-4, N_iters_acc=10)),
            train=dict(lrate_sigma_mlp=0.01, lrate_pose=0.001, lrate_gamma=1.0),
            loss=dict(loss_selection="L1_JS", JS_loss=dict(min_js_score=1.0, max_js_score=10.0, alpha=1.0),
                      decay_los_lambda=False, los_lambda=1000.0, min_depth_eps=0.5, depthloss_lambda=0.005)))


def test_fused_optimizer_runs_the_reference_schedule():
    wc = synth.world_cube("canteen")
    world_cube = types.SimpleNamespace(scale_factor=torch.tensor(wc.scale_factor), shift=torch.tensor(wc.shift))
    opt = FusedOptimizer(_settings(), None, world_cube, 0, False, True, False)
    scans, poses = synth.make_window("canteen", 2, seed=3, n_beams=32, n_azimuth=512)
    kfs = [_KF(scans[k], synth.axis_angle_from_yaw_pose(poses[k]), 3.0 * k) for k in range(2)]
    # keyframe 0: map-only phase (optimizer.py keyframe_schedule[0])
    l0 = opt.iterate_optimizer([kfs[0]])
    assert kfs[0].is_anchored and opt._keyframe_count == 1 and opt._global_step == 30
    assert len(l0) == 1 and l0[0].shape == (30,) and float(l0[0][-1]) < float(l0[0][0])
    # keyframe 1: tracking refinement is skipped (skip_pose_refinement), joint phase moves pose 1 only
    true_pose = kfs[1].get_lidar_pose().get_pose_tensor().clone()
    kfs[1].get_lidar_pose().get_pose_tensor()[:3] += torch.tensor([0.05, -0.04, 0.02])     # a tracking error
    start = kfs[1].get_lidar_pose().get_pose_tensor().clone()
    p0 = kfs[0].get_lidar_pose().get_pose_tensor().clone()
    l1 = opt.iterate_optimizer(kfs)
    assert opt._keyframe_count == 2 and opt._global_step == 50 and len(l1) == 1
    assert torch.equal(kfs[0].get_lidar_pose().get_pose_tensor(), p0)                       # anchored
    moved = kfs[1].get_lidar_pose().get_pose_tensor()
    assert not torch.equal(moved, start) and torch.isfinite(moved).all()
    assert float((moved - start).abs().max()) < 0.05                                        # 20 Adam steps at lr 1e-3
    assert opt._depth_eps is not None and 0.5 <= opt._depth_eps <= 5.5
    # explicit OptimizationSettings override (analysis scripts use it): poses frozen, map only
    opt.iterate_optimizer(kfs, OptimizationSettings(num_iterations=3, freeze_poses=True))
    assert torch.equal(kfs[1].get_lidar_pose().get_pose_tensor(), moved)
    # checkpoint surface (mapper.py:161-175)
    sd = opt._model.state_dict()
    assert list(sd) == ["nerf_model._model_sigma.params"] and sd["nerf_model._model_sigma.params"].numel() == opt._engine.net.param_count
    assert list(opt._occupancy_grid_model.state_dict()) == ["occupancy_grid"]
