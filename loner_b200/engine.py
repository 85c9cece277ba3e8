"""Device-resident mapping step: the hot loop of Optimizer._do_iterate_optimizer
(/root/reference/src/mapping/optimizer.py:276-384) as a fixed sequence of hand-written kernels.

Per iteration (all on the current CUDA stream, no host synchronisation):
  ray pick (torch.randint on device, optimizer.py:288)  ->  loner_ray_build (keyframe.py:71,
  ray_utils.py:269)  ->  loner_sample_ogm / loner_sample_uniform (ray_sampling.py)  ->
  loner_mlp_fwd (nerf_tcnn.py:59)  ->  loner_render_loss (rendering_tcnn.py:71 + optimizer.py:437,
  forward AND backward)  ->  loner_mlp_bwd  ->  [pose gradients: loner_points_bwd,
  loner_ray_build_bwd, Rodrigues by autograd]  ->  [NCCL all-reduce of the flat gradient]  ->
  loner_adam_step + loner_mlp_pack  ->  every N_iters_acc steps loner_ogm_grad + loner_sgd_step
  (optimizer.py:382-384, :598-609).

Multi-GPU: one process per GPU, every rank holds all keyframes, draws its own
rays, and the ranks exchange (a) the two loss normalisers before the loss kernel and (b) one flat
gradient buffer after backward; parameters and optimiser state stay replicated.
"""
import contextlib
import math
from dataclasses import dataclass, field

import torch
import torch.distributed as dist

from . import ops, parallel


@dataclass
class EngineConfig:
    # geometry (WorldCube + ray_range of the sequence file)
    scale: float
    shift: tuple
    ray_range: tuple
    # network (nerf_config: pos_encoding_sigma / sigma_network keys)
    encoding: str = "Frequency"         # pos_encoding_sigma.otype: Frequency | HashGrid (the shipped default)
    n_frequencies: int = 10
    n_neurons: int = 256
    n_hidden_layers: int = 4
    # HashGrid keys of cfg/nerf_config/default_nerf_hash.yaml (with encoding="HashGrid": n_neurons 64, 1 hidden layer)
    n_levels: int = 16
    log2_hashmap_size: int = 18
    base_resolution: int = 16
    per_level_scale: float = 2.0
    # render (default_model_config.yaml:11-20)
    n_samples: int = 512
    perturb: float = 1.0
    raw_noise_std: float = 1.0
    sampler: str = "OGM"
    # occupancy grid (default_model_config.yaml:22-25)
    voxel_size: int = 100
    occ_lr: float = 1e-4
    occ_every: int = 10
    # loss (default_model_config.yaml:40-58)
    min_depth_eps: float = 0.5
    min_js: float = 1.0
    max_js: float = 10.0
    js_alpha: float = 1.0
    los_lambda: float = 1000.0
    depthloss_lambda: float = 0.005
    decay_los_lambda: bool = False      # optimizer.py:448-452 (default_model_config.yaml:44-48)
    los_lambda_decay_rate: float = 0.999
    los_lambda_decay_steps: float = 1.0
    min_los_lambda: float = 100.0
    loss_selection: str = "L1_JS"       # L1_JS | L2_JS | L1_LOS | L2_LOS   (optimizer.py:497-532,568-574)
    depth_eps: float = 3.0              # *_LOS: fixed / decaying margin (optimizer.py:516-521)
    decay_depth_eps: bool = True
    depth_eps_decay_rate: float = 0.95
    depth_eps_decay_steps: float = 1.0
    # train (default_model_config.yaml:27-31)
    lrate_sigma_mlp: float = 0.01
    lrate_pose: float = 0.001
    lrate_gamma: float = 1.0            # ExponentialLR stepped every iteration (optimizer.py:269,378)
    # engine
    rays_selection: str = "RANDOM"      # RANDOM | FIXED | MASK   (optimizer.py:286-296)
    n_sky: int = 0                      # num_samples.sky picks per keyframe among its sky directions (optimizer.py:299-303)
    chunk_rays: int = 16384             # render.chunk (default_model_config.yaml:19); the stash is ~4.3 KB per sample
    seed: int = 0
    net_flags: int = None               # LONER_NET_* kernel A/B variants (None: ops.DEFAULT_NET_FLAGS)

    def los_lambda_at(self, global_step):
        """optimizer.py:448-452: the LOS weight, optionally decayed with the optimiser's global step."""
        if not self.decay_los_lambda:
            return self.los_lambda
        return max(self.los_lambda * self.los_lambda_decay_rate ** ((global_step + 1) / self.los_lambda_decay_steps),
                   self.min_los_lambda)

    def loss_cfg(self, iteration_idx=0, global_step=0):
        """The 9 floats loner_render_loss takes; *_LOS margins follow optimizer.py:516-521."""
        if self.loss_selection not in ("L1_JS", "L2_JS", "L1_LOS", "L2_LOS"):
            raise ValueError(f"Can't use unknown Loss {self.loss_selection}")
        fixed = 0.0
        if self.loss_selection.endswith("LOS"):
            fixed = self.depth_eps
            if self.decay_depth_eps:
                fixed = max(self.depth_eps * self.depth_eps_decay_rate ** (iteration_idx / self.depth_eps_decay_steps),
                            self.min_depth_eps)
        return [self.scale, self.min_depth_eps, self.min_js, self.max_js, self.js_alpha, self.los_lambda_at(global_step),
                self.depthloss_lambda, 1.0 if self.loss_selection.startswith("L2") else 0.0, fixed]


class MappingEngine:
    def __init__(self, cfg: EngineConfig, device="cuda", params: torch.Tensor = None, distributed: bool = True):
        """distributed=False keeps this engine single-GPU inside a torch.distributed process (the reference run of
        the multi-GPU gradient check)."""
        self.cfg = cfg
        self.dev = torch.device(device)
        if cfg.encoding not in ("Frequency", "HashGrid"):
            raise ValueError(f"unknown pos_encoding_sigma.otype {cfg.encoding}")
        self.hash = cfg.encoding == "HashGrid"
        if self.hash:
            self.net = ops.HashNet(cfg.n_levels, 2, cfg.log2_hashmap_size, cfg.base_resolution, cfg.per_level_scale,
                                   cfg.n_neurons, cfg.n_hidden_layers, flags=cfg.net_flags or 0)
        else:
            self.net = ops.Net(cfg.n_frequencies, cfg.n_neurons, cfg.n_hidden_layers, flags=cfg.net_flags)
        if params is None:
            params = xavier_uniform_flat(self.net.layer_shapes(), 1337)
            if self.hash:           # tcnn: table ~ U(-1e-4, 1e-4), after the network matrices
                g = torch.Generator().manual_seed(1338)
                params = torch.cat([params, (torch.rand(2 * self.net.table_entries, generator=g) * 2 - 1) * 1e-4])
        assert params.numel() == self.net.param_count
        self.params = params.detach().to(self.dev, torch.float32).contiguous().clone()
        self.packed = torch.empty(self.net.packed_bytes, device=self.dev, dtype=torch.uint8)
        self._pack()
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.adam_t = 0
        V = cfg.voxel_size
        self.grid = torch.zeros(V, V, V, device=self.dev, dtype=torch.float32)
        self.d_grid = torch.zeros_like(self.grid)
        self.global_step = 0
        # keyframe store: one float4 per return, lidar returns of a keyframe followed by its sky directions
        self.points = torch.empty(0, 4, device=self.dev, dtype=torch.float32)
        self.n_points = 0
        self.kf_offsets = []
        self.kf_sizes = []
        self.kf_sky_offsets = []
        self.kf_sky_sizes = []
        self.kf_masks = []
        # pose store [capacity,6] = [t | axis-angle] per keyframe, with its gradient, Adam moments and step counts;
        # poses6[k] is a VIEW of row k (whose .grad is set after a step that optimised poses)
        self._pose_cap = 0
        self.poses6 = []
        self._grow_pose_store(256)
        self.status = torch.zeros(1, device=self.dev, dtype=torch.int32)      # LONER_STATUS_* bits, see check_status
        self._pose_free_host = []   # per keyframe: its pose is optimised in the current phase
        self.pose_phase = False     # a pose-optimising phase is open (new_phase(optimize_poses=True))
        self._pose_rows = None      # (window, int32 device rows)
        self._grad_rows = []        # keyframes whose poses6[k].grad is currently set
        self.world = dist.get_world_size() if distributed and dist.is_available() and dist.is_initialized() else 1
        self.rank = dist.get_rank() if self.world > 1 else 0
        # every rank draws its own rays and noise: the streams are keyed by (cfg.seed, rank)
        self._seed_base = (int(cfg.seed) * 1000003 + self.rank * 7368787) & 0xFFFFFFFFFFFF
        self._bufs = {}
        self.launches = 0           # kernels of OURS launched (bench reports it)
        self.last = {}
        self.timers = None          # name -> [(start_event, end_event)] when bench.py profiles sections
        self._wcache = {}
        self._pose_cache = None
        self.phase_iteration = 0
        self.train_map = True       # False = "tracking" phase: poses only, MLP frozen (freeze_sigma_mlp)
        self._flat = None           # [MLP grads | pose grads | 4 loss sums]: ONE buffer, all-reduced in place
        self._flat_K = -1
        self._views(1)
        self._z_all = None
        self._last_d_poses12 = None
        self._counters = torch.zeros(2, device=self.dev, dtype=torch.int32)
        self._side = None           # side stream of the loss-normaliser all-reduce (multi-GPU)

    def _grow_pose_store(self, cap):
        n = len(self.poses6)
        new = lambda *shape, dt=torch.float32: torch.zeros(*shape, device=self.dev, dtype=dt)
        store, grad, m, v = new(cap, 6), new(cap, 6), new(cap, 6), new(cap, 6)
        steps, free = new(cap, dt=torch.int32), new(cap, dt=torch.uint8)
        if n:
            for dst, src in ((store, self.pose_store), (grad, self.pose_grad), (m, self.pose_m), (v, self.pose_v),
                             (steps, self.pose_steps), (free, self.pose_free)):
                dst[:n] = src[:n]
        self.pose_store, self.pose_grad, self.pose_m, self.pose_v = store, grad, m, v
        self.pose_steps, self.pose_free = steps, free
        self._pose_cap = cap
        self.poses6 = [store[k] for k in range(n)]
        self._pose_rows = None
        self._grad_rows = []

    def check_status(self):
        """The reference's runtime guards, checked lazily: the kernels OR LONER_STATUS_* bits into a device word and
        this call (one 4-byte read, a host sync) raises what the reference would have raised inside the iteration."""
        bits = int(self.status.item())
        if bits:
            self.status.zero_()
        if bits & ops.STATUS_ORIGIN_OUTSIDE:          # common/ray_utils.py:301-303
            raise AssertionError("ray origins are outside the world cube")
        if bits & ops.STATUS_BAD_POSE_GRAD:           # mapping/optimizer.py:368-370
            raise RuntimeError("Fatal: Encountered invalid gradient in pose.")
        if bits & ops.STATUS_BAD_POSE:                # mapping/optimizer.py:372-374
            raise RuntimeError("Fatal: Encountered invalid pose tensor.")

    def pose_state_dict(self):
        """The pose group of the current phase in torch.optim.Adam's state_dict format (checkpoint surface)."""
        ids = [k for k, f in enumerate(self._pose_free_host) if f]
        steps = self.pose_steps[:len(self.poses6)].cpu()
        state = {i: {"step": torch.tensor(float(steps[k])), "exp_avg": self.pose_m[k].detach().clone(),
                     "exp_avg_sq": self.pose_v[k].detach().clone()}
                 for i, k in enumerate(ids) if int(steps[k]) > 0}
        return {"state": state, "param_groups": [dict(lr=self.cfg.lrate_pose, params=list(range(len(ids))))]}

    def _views(self, K):
        """The flat exchange buffer for a K-keyframe window and its three views.  The kernels write the
        gradients and loss sums straight into it, so the multi-GPU all-reduce needs no staging copies."""
        if self._flat_K != K:
            self._flat = parallel.FlatGrads(self.net.param_count, K, self.dev)
            self.d_params, self._d_poses12, self._loss_acc = self._flat.d_params, self._flat.d_poses12, self._flat.loss_acc
            self._flat_K = K
        return self._flat

    def _pack(self):
        """fp32 master parameters -> the fp16 image the kernels read (after construction and every Adam step)."""
        if self.hash:
            ops.hash_pack(self.net, self.params, self.packed)
        else:
            ops.mlp_pack(self.net, self.params, self.packed)

    def _sigma(self, P, rays, z):
        """Inference forward (no stash)."""
        if self.hash:
            return ops.hash_fwd(self.net, self.packed, P, rays=rays, z=z)
        return ops.mlp_fwd(self.net, self.packed, P, rays=rays, z=z, stash=False)[0]

    # ---------------------------------------------------------------- keyframes
    def add_keyframe(self, ray_directions, distances, pose6, mask=None, sky_rays=None):
        """LidarScan buffers of one keyframe (common/sensors.py:57-82) -> the device store.
        mask: optional bool [M] (LidarScan.mask) used by rays_selection = MASK;
        sky_rays: optional [3,Ks] unit directions (LidarScan.sky_rays); stored as returns at distance
        ray_range[1] + 1 like LidarScan.get_sky_scan (sensors.py:162-167, keyframe.py:92)."""
        self.kf_masks.append(None if mask is None else mask.nonzero(as_tuple=True)[0].to(self.dev))
        pts = ops.pack_points(ray_directions, distances)
        n_l = pts.shape[0]
        n_s = 0
        if sky_rays is not None and sky_rays.numel() > 0:
            sky = ops.pack_points(sky_rays, torch.full_like(sky_rays[0], float(self.cfg.ray_range[1]) + 1.0))
            n_s = sky.shape[0]
            pts = torch.cat([pts, sky.to(pts.device)])
        need = self.n_points + pts.shape[0]
        if need > self.points.shape[0]:            # capacity doubling: adding a keyframe does not re-copy the whole store
            cap = max(need, 2 * self.points.shape[0])
            grown = torch.empty(cap, 4, device=self.dev, dtype=torch.float32)
            grown[:self.n_points] = self.points[:self.n_points]
            self.points = grown
        self.points[self.n_points:need] = pts.to(self.dev)
        self.kf_offsets.append(self.n_points)
        self.kf_sizes.append(n_l)
        self.kf_sky_offsets.append(self.n_points + n_l)
        self.kf_sky_sizes.append(n_s)
        self.n_points = need
        k = len(self.poses6)
        if k >= self._pose_cap:
            self._grow_pose_store(2 * self._pose_cap)
        self.pose_store[k] = pose6.detach().to(self.dev, torch.float32)
        self.poses6.append(self.pose_store[k])
        self._pose_free_host.append(False)
        self._pose_cache = None
        self._wcache = {}
        return len(self.poses6) - 1

    def new_phase(self, optimize_poses: bool, train_map: bool = True, pose_ids=None):
        """A new Adam per optimisation phase, as the reference does (optimizer.py:257-267).
        pose_ids: keyframes whose pose is optimised (default: all but the anchored keyframe 0)."""
        self.train_map = bool(train_map)
        self.phase_iteration = 0            # `iteration_idx` of optimizer.py:276 (drives the *_LOS decay and the LR schedule)
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        self.adam_t = 0
        self._pose_cache = None
        n = len(self.poses6)
        # keyframe 0 is anchored by default
        self._pose_free_host = [bool(optimize_poses and ((k > 0) if pose_ids is None else (k in pose_ids))) for k in range(n)]
        self.pose_phase = bool(optimize_poses) and any(self._pose_free_host)
        self.pose_m.zero_(); self.pose_v.zero_(); self.pose_steps.zero_()
        if n:
            self.pose_free[:n] = torch.tensor(self._pose_free_host, dtype=torch.uint8).to(self.dev)

    # ---------------------------------------------------------------- helpers
    @contextlib.contextmanager
    def _sec(self, name):
        """CUDA-event bracket around one kernel family on the launching stream (bench.py roofline)."""
        if self.timers is None:
            yield
            return
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        yield
        b.record()
        self.timers.setdefault(name, []).append((a, b))

    def _buf(self, name, nbytes):
        b = self._bufs.get(name)
        if b is None or b.numel() < nbytes:
            self._bufs[name] = None
            b = torch.empty(nbytes, device=self.dev, dtype=torch.uint8)
            self._bufs[name] = b
        return b

    def rays_per_keyframe(self, k, n_per_kf):
        """Rays a keyframe contributes per iteration: n lidar picks + num_samples.sky picks if it has sky directions."""
        return n_per_kf + (self.cfg.n_sky if self.cfg.n_sky > 0 and self.kf_sky_sizes[k] > 0 else 0)

    def _window_consts(self, window, n_per_kf):
        """Per-window pick segments (device array of loner_pick_seg_t), built once per window."""
        key = (tuple(window), n_per_kf)
        c = self._wcache.get(key)
        if c is None:
            strategy = self.cfg.rays_selection
            if strategy not in ("RANDOM", "FIXED", "MASK"):
                raise RuntimeError(f"Can't find rays_selection strategy: {strategy}")
            segs, maps, out, map_off = [], [], 0, 0
            for row, k in enumerate(window):
                if strategy == "RANDOM":                    # torch.randint(len(scan), (n,))       optimizer.py:288
                    segs.append((row, ops.PICK_RANDOM, self.kf_offsets[k], self.kf_sizes[k], 0, out))
                elif strategy == "FIXED":                   # torch.arange(n)                       optimizer.py:293-294
                    if n_per_kf > self.kf_sizes[k]:
                        raise IndexError(f"rays_selection FIXED asks for {n_per_kf} rays but keyframe {k} has "
                                         f"{self.kf_sizes[k]} returns")
                    segs.append((row, ops.PICK_FIXED, self.kf_offsets[k], self.kf_sizes[k], 0, out))
                else:                                       # picks among scan.mask.nonzero()       optimizer.py:289-292
                    m = self.kf_masks[k]
                    if m is None or m.numel() == 0:
                        raise RuntimeError("rays_selection MASK needs add_keyframe(..., mask=...) with a non-empty mask")
                    segs.append((row, ops.PICK_MASK, self.kf_offsets[k], m.numel(), map_off, out))
                    maps.append(m)
                    map_off += m.numel()
                out += n_per_kf
                if self.cfg.n_sky > 0 and self.kf_sky_sizes[k] > 0:     # optimizer.py:299-303, pose detached keyframe.py:93-95
                    segs.append((row | ops.KF_DETACHED, ops.PICK_RANDOM, self.kf_sky_offsets[k], self.kf_sky_sizes[k], 0, out))
                    out += self.cfg.n_sky
            c = dict(segs=ops.pick_segments(segs, self.dev), n_segs=len(segs), n_rays=out,
                     index_map=torch.cat(maps).contiguous() if maps else None,
                     ray_kf=torch.empty(out, device=self.dev, dtype=torch.int32),
                     ray_point=torch.empty(out, device=self.dev, dtype=torch.int64))
            self._wcache = {key: c}
        return c

    def _pick_rays(self, window, n_per_kf):
        """ray pick of optimizer.py:286-305 as ONE kernel (loner_ray_pick)."""
        c = self._window_consts(window, n_per_kf)
        seed = (self._seed_base + self.global_step * 104729 + 5) & 0x7FFFFFFFFFFF
        ops.ray_pick(c["segs"], c["n_segs"], c["index_map"], seed, c["n_rays"], c["ray_kf"], c["ray_point"])
        self.launches += 1
        return c["ray_kf"], c["ray_point"]

    def _injected_rays(self, window, n_per_kf, ray_point, ray_kf=None):
        """Parity tests inject the picked indices: validated like the reference's indexing would (IndexError).
        ray_kf (optional, int32, LONER_KF_DETACHED allowed) is needed when sky rays are injected too."""
        ray_point = ray_point.to(self.dev)
        if ray_point.dtype != torch.int64:
            raise TypeError("injected ray_point must be int64")
        K = len(window)
        if ray_kf is not None:
            ray_kf = ray_kf.to(self.dev, torch.int32).contiguous()
            rows = (ray_kf & ops.KF_MASK).long()
            if ray_kf.numel() != ray_point.numel() or bool((rows >= K).any()):
                raise IndexError("injected ray_kf does not match the window")
            lo = torch.tensor([self.kf_offsets[k] for k in window], device=self.dev)[rows]
            hi = torch.tensor([self.kf_sky_offsets[k] + self.kf_sky_sizes[k] for k in window], device=self.dev)[rows]
            if bool(((ray_point < lo) | (ray_point >= hi)).any()):
                raise IndexError("injected ray_point outside its keyframe's scan")
            return ray_kf, ray_point.contiguous()
        if ray_point.numel() != K * n_per_kf:
            raise IndexError(f"injected ray_point has {ray_point.numel()} entries, expected {K} x {n_per_kf}")
        lo = torch.tensor([self.kf_offsets[k] for k in window], device=self.dev).repeat_interleave(n_per_kf)
        hi = torch.tensor([self.kf_sky_offsets[k] + self.kf_sky_sizes[k] for k in window], device=self.dev).repeat_interleave(n_per_kf)
        if bool(((ray_point < lo) | (ray_point >= hi)).any()):
            raise IndexError("injected ray_point outside its keyframe's scan")
        ray_kf = torch.arange(K, device=self.dev, dtype=torch.int32).repeat_interleave(n_per_kf).contiguous()
        return ray_kf, ray_point.contiguous()

    def _poses12(self, window, optimize_poses):
        """[K,12] pose matrices of the window (loner_pose_matrices); recomputed every step only while poses move."""
        key = tuple(window)
        if not optimize_poses and self._pose_cache is not None and self._pose_cache[0] == key:
            return self._pose_cache[1]
        if self._pose_rows is None or self._pose_rows[0] != key:
            self._pose_rows = (key, torch.tensor(list(window), dtype=torch.int32).to(self.dev))
        p12 = ops.pose_matrices(self.pose_store, self._pose_rows[1], shift=self.cfg.shift, scale=self.cfg.scale,
                                status=self.status)
        self.launches += 1
        self._pose_cache = None if optimize_poses else (key, p12)
        return p12

    def _lr(self, base):
        """ExponentialLR(gamma) stepped after every iteration of the phase (optimizer.py:269,378)."""
        g = self.cfg.lrate_gamma
        return base if g == 1.0 else base * g ** self.phase_iteration

    def _reduce_counts_async(self, counters):
        """Global loss normalisers (#valid, #opaque): all-reduced on a side stream right after ray_build,
        hidden behind the sampler and the forward kernel; returns the event the loss kernel waits for."""
        if self.world == 1:
            return None
        if self._side is None:
            self._side = torch.cuda.Stream(device=self.dev)
        cur = torch.cuda.current_stream(self.dev)
        self._side.wait_stream(cur)
        with torch.cuda.stream(self._side):
            parallel.allreduce_counts(counters)
            ev = torch.cuda.Event()
            ev.record(self._side)
        return ev

    # ---------------------------------------------------------------- the step
    def step(self, window, n_per_kf, optimize_poses=False, injected=None, want_outputs=False):
        """One mapping iteration over `window` (keyframe ids), n_per_kf lidar rays (+ cfg.n_sky sky rays) per
        keyframe on THIS rank.  injected: optional dict(ray_point, u1, u2, noise) for parity tests.
        Returns the loss (0-dim device tensor, no host sync)."""
        cfg = self.cfg
        K = len(window)
        self._views(K)
        if injected is not None and "ray_point" in injected:
            ray_kf, ray_point = self._injected_rays(window, n_per_kf, injected["ray_point"], injected.get("ray_kf"))
        else:
            ray_kf, ray_point = self._pick_rays(window, n_per_kf)
        p12 = self._poses12(window, optimize_poses)
        self._flat.zero_()                 # d_params, d_poses12, loss sums: one memset
        counters = self._counters.zero_()
        rays, depths, flags = ops.ray_build(self.points, ray_kf, ray_point, p12, cfg.shift, cfg.scale,
                                            cfg.ray_range, counters)
        self.launches += 1
        counts_ready = self._reduce_counts_async(counters)
        loss_acc, d_rays = self._forward_backward(rays, depths, flags, counters, optimize_poses, injected,
                                                  want_outputs, counts_ready)
        d_poses12 = None
        if optimize_poses:
            d_poses12 = ops.ray_build_bwd(self.points, ray_kf, ray_point, p12, cfg.shift, cfg.scale, cfg.ray_range,
                                          d_rays, out=self._d_poses12)
            self.launches += 1
        loss = self._exchange_and_update(loss_acc, counters, d_poses12)
        if d_poses12 is not None:
            # chain rule to the 6-vectors + the pose group's Adam step (optimizer.py:249-267,:376): one kernel
            ops.pose_step(self.pose_store, self._pose_rows[1], self.pose_free, self._last_d_poses12, self.pose_grad,
                          self.pose_m, self.pose_v, self.pose_steps, self._lr(cfg.lrate_pose), apply=self.pose_phase,
                          status=self.status)
            self.launches += 1
            for k in self._grad_rows:
                self.poses6[k].grad = None
            self._grad_rows = [k for k in window if self._pose_free_host[k]]
            for k in self._grad_rows:
                self.poses6[k].grad = self.pose_grad[k]
        self._occupancy_update(rays, depths, flags)
        return loss

    def step_from_host(self, rays_host, depths_host):
        """The reference-facing call (Optimizer.compute_loss + backward + Adam, optimizer.py:340-380)
        fed with HOST rays [N,13] and depths [N] (pinned), as the reference's data_prep_on_cpu path
        does: H2D copy -> step -> loss read back."""
        cfg = self.cfg
        self._views(1)
        self._flat.zero_()
        rays = rays_host.to(self.dev, non_blocking=True)
        depths = depths_host.to(self.dev, non_blocking=True)
        far, near = rays[:, 12], rays[:, 11]
        valid = far > near + 1.0 / cfg.scale
        opaque = valid & (depths > 0) & ~(depths > far)
        flags = (valid.to(torch.uint8) + 2 * opaque.to(torch.uint8)).contiguous()
        counters = torch.stack([valid.sum(), opaque.sum()]).to(torch.int32)
        counts_ready = self._reduce_counts_async(counters)
        loss_acc, _ = self._forward_backward(rays, depths, flags, counters, False, None, False, counts_ready)
        loss = self._exchange_and_update(loss_acc, counters, None)
        self._occupancy_update(rays, depths, flags)
        return float(loss.item())

    def _forward_backward(self, rays, depths, flags, counters, optimize_poses, injected, want_outputs, counts_ready=None):
        cfg = self.cfg
        N, S = rays.shape[0], cfg.n_samples
        los_lambda = cfg.los_lambda_at(self.global_step)
        gscale = ops.default_grad_scale(N * self.world, S, los_lambda)
        loss_acc = self._loss_acc
        d_rays = torch.zeros(N, ops.RAY_COLS, device=self.dev, dtype=torch.float32) if optimize_poses else None
        outs = []
        seed = (self._seed_base + self.global_step * 7919 + 13) & 0x7FFFFFFFFFFF
        keep_z = self.global_step % cfg.occ_every == 0 and cfg.sampler == "OGM"
        self._z_all = [] if keep_z else None
        loss_cfg = cfg.loss_cfg(self.phase_iteration, self.global_step)
        for c0 in range(0, N, cfg.chunk_rays):
            c1 = min(N, c0 + cfg.chunk_rays)
            r, dpt, fl = rays[c0:c1], depths[c0:c1], flags[c0:c1]
            P = (c1 - c0) * S
            inj = injected or {}
            u1 = inj["u1"][c0:c1].contiguous().to(self.dev) if "u1" in inj else None
            u2 = inj["u2"][c0:c1].contiguous().to(self.dev) if inj.get("u2") is not None else None
            noise = inj["noise"][c0:c1].contiguous().to(self.dev) if "noise" in inj else None
            with self._sec("sample"):
                if cfg.sampler == "OGM":
                    z = ops.sample_ogm(r, self.grid, S, cfg.perturb, u1, u2, seed=seed + c0)
                else:
                    z = ops.sample_uniform(r, S, cfg.perturb, u1, seed=seed + c0)
            acts = None if self.hash else self._buf("acts", self.net.act_bytes(P))
            with self._sec("mlp_fwd"):
                if self.hash:       # nothing is stashed: the hash-grid backward recomputes the forward
                    sigma = ops.hash_fwd(self.net, self.packed, P, rays=r, z=z)
                else:
                    sigma, _ = ops.mlp_fwd(self.net, self.packed, P, rays=r, z=z, stash=True, acts=acts)
            if counts_ready is not None:
                torch.cuda.current_stream(self.dev).wait_event(counts_ready)
                counts_ready = None
            with self._sec("render_loss"):
                res = ops.render_loss(sigma, z, r, dpt, fl, counters, loss_cfg, noise=noise,
                                      raw_noise_std=cfg.raw_noise_std, seed=seed + c0 + 1, loss_acc=loss_acc,
                                      want_outputs=want_outputs, d_rays=d_rays[c0:c1] if optimize_poses else None)
            scratch = self._buf("scratch", max(self.net.bwd_scratch_bytes(P), 16))
            if self.hash:
                with self._sec("mlp_dgrad"):
                    d_pos = ops.hash_bwd(self.net, self.packed, P, res["d_sigma"], gscale, self.d_params, rays=r, z=z,
                                         want_dpos=optimize_poses, scratch=scratch)
                self.launches += 3 + 2      # sampler, forward, render/loss; backward, partial reduce
            else:
                with self._sec("mlp_dgrad"):
                    d_pos = ops.mlp_dgrad(self.net, self.packed, P, res["d_sigma"], acts, gscale, scratch, rays=r, z=z,
                                          want_dpos=optimize_poses)
                with self._sec("mlp_wgrad"):
                    ops.mlp_wgrad(self.net, self.packed, P, res["d_sigma"], acts, gscale, self.d_params, scratch)
                self.launches += 3 + ops.MLP_BWD_LAUNCHES
            if optimize_poses:
                ops.points_bwd(d_pos, z, d_rays[c0:c1])
                self.launches += 1
            if keep_z:
                self._z_all.append(z)
            if want_outputs:
                res["z_vals"] = z
                res["sigma"] = sigma
                outs.append(res)
        self.last = dict(outs=outs, rays=rays, depths=depths, flags=flags)
        return loss_acc, d_rays

    def _exchange_and_update(self, loss_acc, counters, d_poses12):
        """ONE in-place all-reduce of [MLP grads | pose grads | 4 loss sums], then Adam + fp16 repack."""
        cfg = self.cfg
        if self.world > 1:
            self._flat.allreduce()
        self._last_d_poses12 = d_poses12
        if self.train_map:
            self.adam_t += 1
            with self._sec("adam_pack"):
                ops.adam_step(self.params, self.d_params, self.exp_avg, self.exp_avg_sq, self.adam_t,
                              self._lr(cfg.lrate_sigma_mlp))
                self._pack()
            self.launches += 2
        fin = ops.loss_finalize(loss_acc, counters, cfg.depthloss_lambda, cfg.los_lambda_at(self.global_step), cfg.n_samples)
        self.launches += 1
        loss = fin[0]
        self.last.update(loss_terms=fin, counters=counters, depth_eps=fin[1])
        return loss

    def _occupancy_update(self, rays, depths, flags=None):
        """Every occ_every global steps (optimizer.py:382-384): scatter the pseudo-gradient, SGD step."""
        cfg = self.cfg
        if self._z_all is not None:
            self.d_grid.zero_()
            with self._sec("ogm_update"):
                for ci, c0 in enumerate(range(0, rays.shape[0], cfg.chunk_rays)):
                    c1 = min(rays.shape[0], c0 + cfg.chunk_rays)
                    ops.ogm_grad(rays[c0:c1], self._z_all[ci], depths[c0:c1], cfg.scale, cfg.voxel_size, self.d_grid,
                                 flags=None if flags is None else flags[c0:c1])
                    self.launches += 1
                if self.world > 1:
                    dist.all_reduce(self.d_grid, op=dist.ReduceOp.SUM)
                ops.sgd_step(self.grid, self.d_grid, cfg.occ_lr)
            self.launches += 1
            self._z_all = None
        self.global_step += 1
        self.phase_iteration += 1

    # ---------------------------------------------------------------- inference (test mode)
    @torch.no_grad()
    def render(self, rays, n_samples=None, seed=0, injected=None, want_weights=False):
        """Model.forward(testing=True) (models/model_tcnn.py:70-105): N_samples_test samples, perturb = 0; like
        the reference the importance draws and raw_noise_std stay active, so parity
        tests inject u2 / noise.  Processed in chunks of cfg.chunk_rays rays like the reference's chunk loop
        (model_tcnn.py:82)."""
        cfg = self.cfg
        S = n_samples or cfg.n_samples
        inj = injected or {}
        N = rays.shape[0]
        step = max(1, min(cfg.chunk_rays, (cfg.chunk_rays * cfg.n_samples) // S))     # same samples per chunk as training
        parts = []
        for c0 in range(0, N, step):
            c1 = min(N, c0 + step)
            r = rays[c0:c1].contiguous()
            u2 = inj["u2"][c0:c1].contiguous().to(self.dev) if inj.get("u2") is not None else None
            noise = inj["noise"][c0:c1].contiguous().to(self.dev) if inj.get("noise") is not None else None
            if cfg.sampler == "OGM":
                z = ops.sample_ogm(r, self.grid, S, 0.0, None, u2, seed=seed + c0)
            else:
                z = ops.sample_uniform(r, S, 0.0, None, seed=seed + c0)
            sigma = self._sigma(r.shape[0] * S, r, z)
            w, d, o, v = ops.render_fwd(sigma, z, r, noise=noise, raw_noise_std=cfg.raw_noise_std, seed=seed + c0 + 1,
                                        want_weights=want_weights)
            self.launches += 3
            parts.append((d, o, v, z, w))
        cat = (lambda i: parts[0][i] if len(parts) == 1 else torch.cat([p[i] for p in parts]))
        out = dict(depth_fine=cat(0), opacity_fine=cat(1), variance=cat(2), samples_fine=cat(3))
        if want_weights:
            out["weights_fine"] = cat(4)
        return out


def xavier_uniform_flat(layer_shapes, seed):
    """Flat fp32 params ([out,in] row-major per layer), xavier-uniform like tcnn's default init."""
    g = torch.Generator().manual_seed(seed)
    chunks = []
    for (n_out, n_in) in layer_shapes:
        bound = math.sqrt(6.0 / (n_in + n_out))
        chunks.append(((torch.rand(n_out, n_in, generator=g) * 2 - 1) * bound).reshape(-1))
    return torch.cat(chunks)
