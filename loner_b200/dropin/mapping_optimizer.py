"""FusedOptimizer: the reference's `Optimizer` contract (/root/reference/src/mapping/optimizer.py:74-192)
on top of the fused device-resident step (loner_b200.engine.MappingEngine).

`Mapper` constructs `Optimizer(settings.optimizer, calibration, world_cube, 0, use_gt_poses, lidar_only,
enable_sky_segmentation)` (mapping/mapper.py:63) and calls `iterate_optimizer(active_window)` per keyframe
(mapper.py:104); this class takes the same arguments, reads the same settings keys, follows the same
keyframe / iteration schedule (optimizer.py:144-269) and exposes the attributes the mapper reads
(`_model`, `_optimizer`, `_occupancy_grid_model`, `_occupancy_grid_optimizer`, `_global_step`,
`_keyframe_count`, mapper.py:108-175).  To use it:

    import mapping.optimizer as mo
    from loner_b200.dropin.mapping_optimizer import FusedOptimizer
    mo.Optimizer = FusedOptimizer            # before Mapper is constructed

Keyframes are duck-typed: `.get_lidar_scan()` (with `ray_directions [3,M]`, `distances [M]`),
`.get_lidar_pose().get_pose_tensor()` ([6] = t, axis-angle), `.is_anchored`, `.get_time()`.
"""
from dataclasses import dataclass

import torch
import torch.nn as nn

from loner_b200 import engine as eng


@dataclass
class OptimizationSettings:
    """Same fields as the reference's container (optimizer.py:42-61)."""
    num_iterations: int = 1
    freeze_poses: bool = False
    latest_kf_only: bool = False
    freeze_sigma_mlp: bool = False
    freeze_rgb_mlp: bool = False

    @staticmethod
    def from_dict(d):
        return OptimizationSettings(d.get("num_iterations", 1), d.get("freeze_poses", False),
                                    d.get("latest_kf_only", False), d.get("freeze_sigma_mlp", False),
                                    d.get("freeze_rgb_mlp", False))


class _ParamView(nn.Module):
    """Checkpoint surface: `state_dict()` with the reference's key names AND shapes, so that
    `Model(...).load_state_dict(ckpt['network_state_dict'])` / `OccupancyGridModel(...).load_state_dict(
    ckpt['occ_model_state_dict'])` of the analysis scripts (analysis/renderer_lidar.py:172-180,
    compute_l1_depth.py:146-153) accept what Mapper.build_ckpt (mapping/mapper.py:161-175) saved."""

    def __init__(self, key, tensor, shape=None):
        super().__init__()
        self._key, self._t = key, tensor
        self._shape = tuple(shape) if shape is not None else tuple(tensor.shape)

    def state_dict(self, *a, **kw):
        return {self._key: self._t.detach().clone().reshape(self._shape)}

    def load_state_dict(self, sd, strict=True):
        if strict and set(sd) != {self._key}:
            raise RuntimeError(f"unexpected / missing keys in state_dict: {sorted(set(sd) ^ {self._key})}")
        src = sd[self._key]
        if src.numel() != self._t.numel():
            raise RuntimeError(f"size mismatch for {self._key}: {tuple(src.shape)} vs {self._shape}")
        self._t.copy_(src.reshape(self._t.shape))


class _AdamState:
    """`Optimizer._optimizer.state_dict()` in torch.optim.Adam's format (mapper.py:167): one param group for the
    sigma network (state index 0), and the pose group of the current phase when poses are optimised."""

    def __init__(self, engine):
        self._e = engine

    def state_dict(self):
        e = self._e
        base = dict(betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False, maximize=False, foreach=None,
                    capturable=False, differentiable=False, fused=None)
        state = {0: {"step": torch.tensor(float(e.adam_t)), "exp_avg": e.exp_avg.detach().clone(),
                     "exp_avg_sq": e.exp_avg_sq.detach().clone()}}
        groups = [dict(base, lr=e.cfg.lrate_sigma_mlp, params=[0])]
        if e.pose_phase:
            sd = e.pose_state_dict()
            ids = sd["param_groups"][0]["params"]
            for i in ids:
                if i in sd["state"]:
                    state[1 + i] = sd["state"][i]
            groups.append(dict(base, lr=e.cfg.lrate_pose, params=[1 + i for i in ids]))
        return {"state": state, "param_groups": groups}


class _SgdState:
    """`_occupancy_grid_optimizer.state_dict()`: torch.optim.SGD without momentum is stateless (optimizer.py:108-109)."""

    def __init__(self, lr):
        self._lr = lr

    def state_dict(self):
        return {"state": {}, "param_groups": [dict(lr=self._lr, momentum=0, dampening=0, weight_decay=0, nesterov=False,
                                                   maximize=False, foreach=None, differentiable=False, fused=None,
                                                   params=[0])]}


class FusedOptimizer:
    def __init__(self, settings, calibration, world_cube, device, use_gt_poses=False, lidar_only=True,
                 enable_sky_segmentation=True):
        if not lidar_only:
            raise NotImplementedError("camera supervision is disabled in the reference (optimizer.py:433-434)")
        self._settings = settings
        self._calibration = calibration
        self._device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self._use_gt_poses = use_gt_poses
        self._lidar_only = lidar_only
        self._optimization_settings = OptimizationSettings()
        mc = settings.model_config
        self._model_config = mc
        self._scale_factor = world_cube.scale_factor
        m = mc.model
        nc = m.nerf_config
        enc, net = nc["pos_encoding_sigma"], nc["sigma_network"]
        if enc["otype"] not in ("Frequency", "HashGrid"):
            raise NotImplementedError("sigma-head encoding %r: Frequency and HashGrid are implemented" % enc["otype"])
        if mc.loss.loss_selection not in ("L1_JS", "L2_JS", "L1_LOS", "L2_LOS"):
            raise ValueError(f"Can't use unknown Loss {mc.loss.loss_selection}")
        sampler = settings.samples_selection.strategy
        if sampler not in ("OGM", "UNIFORM"):
            raise RuntimeError(f"Can't find samples_selection strategy: {sampler}")
        if settings.rays_selection.strategy not in ("RANDOM", "FIXED", "MASK"):
            raise RuntimeError(f"Can't find rays_selection strategy: {settings.rays_selection.strategy}")
        cfg = eng.EngineConfig(
            scale=float(world_cube.scale_factor), shift=tuple(float(x) for x in world_cube.shift),
            ray_range=tuple(float(x) for x in m.ray_range),
            encoding=enc["otype"], n_levels=int(enc.get("n_levels", 16)),
            log2_hashmap_size=int(enc.get("log2_hashmap_size", 18)), base_resolution=int(enc.get("base_resolution", 16)),
            per_level_scale=float(enc.get("per_level_scale", 2.0)),
            n_frequencies=int(enc.get("n_frequencies", 10)), n_neurons=int(net["n_neurons"]),
            n_hidden_layers=int(net["n_hidden_layers"]), n_samples=int(m.render.N_samples_train),
            perturb=float(m.render.perturb), raw_noise_std=float(m.render.raw_noise_std), sampler=sampler,
            voxel_size=int(m.occ_model.voxel_size), occ_lr=float(m.occ_model.lr), occ_every=int(m.occ_model.N_iters_acc),
            min_depth_eps=float(mc.loss.min_depth_eps), min_js=float(mc.loss.JS_loss.min_js_score),
            max_js=float(mc.loss.JS_loss.max_js_score), js_alpha=float(mc.loss.JS_loss.alpha),
            los_lambda=float(mc.loss.los_lambda), depthloss_lambda=float(mc.loss.depthloss_lambda),
            decay_los_lambda=bool(mc.loss.get("decay_los_lambda", False)),
            los_lambda_decay_rate=float(mc.loss.get("los_lambda_decay_rate", 0.999)),
            los_lambda_decay_steps=float(mc.loss.get("los_lambda_decay_steps", 1)),
            min_los_lambda=float(mc.loss.get("min_los_lambda", 100.0)),
            loss_selection=mc.loss.loss_selection, depth_eps=float(mc.loss.get("depth_eps", 3.0)),
            decay_depth_eps=bool(mc.loss.get("decay_depth_eps", True)),
            depth_eps_decay_rate=float(mc.loss.get("depth_eps_decay_rate", 0.95)),
            depth_eps_decay_steps=float(mc.loss.get("depth_eps_decay_steps", 1)),
            rays_selection=settings.rays_selection.strategy,
            lrate_sigma_mlp=float(mc.train.lrate_sigma_mlp), lrate_pose=float(mc.train.lrate_pose),
            lrate_gamma=float(mc.train.get("lrate_gamma", 1.0)),
            n_sky=int(settings.num_samples.get("sky", 0)) if enable_sky_segmentation else 0,
            chunk_rays=int(m.render.chunk))
        self._enable_sky_segmentation = enable_sky_segmentation
        self._engine = eng.MappingEngine(cfg, device=self._device)
        V = cfg.voxel_size
        self._model = _ParamView("nerf_model._model_sigma.params", self._engine.params)
        self._occupancy_grid_model = _ParamView("occupancy_grid", self._engine.grid, shape=(1, 1, V, V, V))
        self._optimizer = _AdamState(self._engine)
        self._occupancy_grid_optimizer = _SgdState(cfg.occ_lr)
        self._keyframe_count = 0
        self._global_step = 0
        self._keyframe_schedule = settings["keyframe_schedule"]
        self._num_lidar_samples = settings.num_samples.lidar
        self._kf_ids = {}
        self._depth_eps = None
        self._progress_bar = None

    # ------------------------------------------------------------------ schedule (optimizer.py:144-156)
    def iterate_optimizer(self, keyframe_window, optimizer_settings=None):
        cumulative = 0
        for item in self._keyframe_schedule:
            kf_count, iteration_schedule = item["num_keyframes"], item["iteration_schedule"]
            cumulative += kf_count
            if cumulative >= self._keyframe_count + 1 or kf_count == -1:
                break
        result = self._do_iterate_optimizer(keyframe_window, iteration_schedule, optimizer_settings=optimizer_settings)
        self._keyframe_count += 1
        return result

    def _pose_of(self, kf):
        """KeyFrame.build_lidar_rays builds from the ground-truth pose when use_gt_poses is set (keyframe.py:82-85)."""
        if self._use_gt_poses:
            return kf._frame._gt_lidar_pose.get_pose_tensor()
        return kf.get_lidar_pose().get_pose_tensor()

    def _register(self, kf):
        # keyed by id(kf) with a strong reference held next to the slot, so an id cannot be recycled for another
        # keyframe; like the reference's KeyFrameManager (keyframe_manager.py) keyframes are never dropped
        entry = self._kf_ids.get(id(kf))
        if entry is None:
            scan = kf.get_lidar_scan()
            k = self._engine.add_keyframe(scan.ray_directions, scan.distances, self._pose_of(kf),
                                          mask=getattr(scan, "mask", None),
                                          sky_rays=getattr(scan, "sky_rays", None) if self._enable_sky_segmentation else None)
            self._kf_ids[id(kf)] = (k, kf)
        else:   # the tracker / previous phases may have moved the pose
            k = entry[0]
            self._engine.poses6[k].data.copy_(self._pose_of(kf).detach().to(self._device))
            self._engine._pose_cache = None
        return k

    def _do_iterate_optimizer(self, keyframe_window, iteration_schedule, profiler=None, optimizer_settings=None):
        if len(keyframe_window) == 1:
            keyframe_window[0].is_anchored = True
        if len(iteration_schedule) > 1 and self._settings.skip_pose_refinement:
            iteration_schedule = iteration_schedule[1:]
        if optimizer_settings is not None:
            iteration_schedule = [None]
        losses = []
        for it_cfg in iteration_schedule:
            o = self._optimization_settings
            if optimizer_settings is None:
                o.freeze_poses = it_cfg["freeze_poses"] or self._settings.freeze_poses or self._use_gt_poses
                o.latest_kf_only = it_cfg["latest_kf_only"] if "latest_kf_only" in it_cfg else False
                o.freeze_rgb_mlp = it_cfg["freeze_rgb_mlp"]
                o.freeze_sigma_mlp = it_cfg["freeze_sigma_mlp"]
                o.num_iterations = it_cfg["num_iterations"]
            else:
                self._optimization_settings = o = optimizer_settings
            o.freeze_poses = o.freeze_poses or self._settings.freeze_poses or self._use_gt_poses
            optimize_poses = not o.freeze_poses
            active = keyframe_window
            if o.latest_kf_only:
                active = [max(keyframe_window, key=lambda kf: float(kf.get_time()))]
            ids = [self._register(kf) for kf in active]
            free = {k for kf, k in zip(active, ids) if not kf.is_anchored}
            if not (optimize_poses or not o.freeze_sigma_mlp):
                continue                                                 # nothing to optimise (should_enable_lidar)
            self._engine.new_phase(optimize_poses and len(free) > 0, train_map=not o.freeze_sigma_mlp, pose_ids=free)
            phase_losses = []
            for _ in range(o.num_iterations):
                loss = self._engine.step(ids, self._num_lidar_samples, optimize_poses=optimize_poses and len(free) > 0)
                phase_losses.append(loss)
                self._global_step += 1
            self._depth_eps = float(self._engine.last["depth_eps"])
            self._engine.check_status()        # the pose / world-cube guards of optimizer.py:368-374, ray_utils.py:301-303
            # hand the optimised poses back to the keyframes (their tensors are what the mapper emits)
            for kf, k in zip(active, ids):
                if k in free and optimize_poses:
                    kf.get_lidar_pose().get_pose_tensor().data.copy_(self._engine.poses6[k].detach().to(
                        kf.get_lidar_pose().get_pose_tensor().device))
            losses.append(torch.stack(phase_losses).detach())
        if losses:
            last = losses[-1]
            if not torch.isfinite(last).all():
                raise AssertionError("NaN Loss Encountered")                # optimizer.py:590
        return losses

    def should_enable_lidar(self):
        o = self._optimization_settings
        return not o.freeze_sigma_mlp or not o.freeze_poses

    def should_enable_camera(self):
        return False
