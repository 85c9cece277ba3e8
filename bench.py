#!/usr/bin/env python
"""bench.py — rays/sec of the LONER mapping step (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps 20 --warmup 5                 # our arm (hand-written sm_100a kernels)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --steps 3 --warmup 1          # reference arm: CPU port on host cores

A step = one full mapping iteration over one batch of synthetic rays: ray pick + ray build +
occupancy-guided sampling + encoding/MLP forward + volume render + JS-margin loss + backward +
Adam (+ occupancy-grid update every 10th step), exactly the loop body of
/root/reference/src/mapping/optimizer.py:276-384.  Prints ONE JSON line (rank 0).

The line's headline (`value`, `roofline`, `e2e`) is BASELINE.json configs[1] (C2) per GPU.  It also carries:
`hash` (the sigma head the reference SHIPS, at the C2 size and at the reference's default operating point),
`api` (the reference-facing FusedOptimizer.iterate_optimizer), and at N > 1 `grad_check` (k-GPU == 1-GPU
gradients on one global ray set), `c5` (BASELINE configs[4], weak) and at N = 4 `c4` (configs[3], strong).
"""
import argparse
import json
import os
import re
import subprocess
import sys
import threading
import time
import types

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on, per GPU
    "c2": dict(geom="canteen", K=1, rays_per_gpu=8192, S=512, W=256, L=4, poses=False,
               label="C2 canteen geometry, synthetic 64x1024 scan, 8192 rays/GPU x 512 samples, 4x256 MLP (Frequency-10)"),
    "c3": dict(geom="garden", K=8, rays_per_gpu=16384, S=512, W=256, L=4, poses=False,
               label="C3 garden geometry, 8-keyframe window, 16384 rays/GPU x 512 samples, 4x256 MLP"),
    # configs[3]: 65536 rays x 256 samples in total, ray-sharded (strong scaling); rays_per_gpu = 65536 / N
    "c4": dict(geom="quad", K=8, rays_total=65536, S=256, W=256, L=4, poses=False,
               label="C4 Newer College quad geometry, 65536 rays x 256 samples in total, ray-sharded, 4x256 MLP"),
    "c5": dict(geom="canteen", K=16, rays_per_gpu=32768, S=512, W=256, L=4, poses=True,
               label="C5 weak scaling, 16-keyframe window, 32768 rays/GPU x 512 samples, joint pose+map"),
    "c1": dict(geom="canteen", K=1, rays_per_gpu=2048, S=128, W=64, L=2, poses=False,
               label="C1 single scan, 2048 rays x 128 samples, 2x64 MLP"),
    "smoke": dict(geom="canteen", K=2, rays_per_gpu=512, S=128, W=128, L=2, poses=True, label="smoke"),
    # the sigma head the reference SHIPS (cfg/nerf_config/default_nerf_hash.yaml) at the C2 size
    "c2hash": dict(geom="canteen", K=1, rays_per_gpu=8192, S=512, W=64, L=1, poses=False, encoding="HashGrid",
                   label="C2 geometry and size with the reference's shipped sigma head: HashGrid (16 levels x 2, 2^18) + 1x64 MLP"),
    # the reference's own operating point: window of 8 keyframes x 512 rays x 512 samples, joint optimisation
    # (cfg/defaults.yaml:60,70; default_model_config.yaml:12), shipped HashGrid head
    "refdefault": dict(geom="canteen", K=8, rays_per_gpu=4096, S=512, W=64, L=1, poses=True, encoding="HashGrid",
                       label="reference default operating point: 8 keyframes x 512 rays x 512 samples, HashGrid + 1x64, joint pose+map"),
}


def rays_per_gpu(wl, world):
    return wl["rays_per_gpu"] if "rays_per_gpu" in wl else wl["rays_total"] // world


def flops_per_sample(wl):
    """Algorithmic FLOPs per sample: F_fwd = 2 (E_pad W + (L-1) W^2 + W); a training step is
    fwd + dgrad + wgrad, the first-layer dgrad only when input gradients (poses) are needed."""
    W, L = wl["W"], wl["L"]
    e_pad = 32 if wl.get("encoding") == "HashGrid" else 64
    f_fwd = 2 * (e_pad * W + (L - 1) * W * W + W)
    f_train = 3 * f_fwd - (0 if wl["poses"] else 2 * e_pad * W)
    return e_pad, f_fwd, f_train


def describe(wl, world):
    hashed = wl.get("encoding") == "HashGrid"
    return {"workload": wl["label"], "geometry": wl["geom"], "keyframes": wl["K"], "rays_per_gpu": rays_per_gpu(wl, world),
            "samples_per_ray": wl["S"], "mlp": f"{wl['L']}x{wl['W']}",
            "encoding": "HashGrid(16 levels x 2 features, 2^18 entries) -> 32" if hashed else "Frequency(10) -> 64",
            "pose_optimisation": wl["poses"], "sampler": "OGM", "parallelism": f"ray-sharded dp{world}",
            "l2_policy": ("no stash (the backward recomputes); the 14.8 MB table and its 29.7 MB gradient are meant to stay in L2; "
                          "z/sigma/d_sigma streams (50 MB/step) exceed nothing: steps are back to back, no flush"
                          if hashed else
                          "activation + gradient stash (>8 GB per step) streams through HBM: inputs larger than L2, no flush needed")}


def load_peaks():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], tflops_burst=d["bf16_tflops"], tflops_sustained=d["bf16_tflops_sustained"],
                    source="MEASURED_PEAKS.json (measured)")
    return dict(hbm_gbs=6650.0, tflops_burst=1590.0, tflops_sustained=1400.0, source="B200_PROFILING.md fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons, sampled every 20 ms from before the warm-up until after the
    timed region; only samples whose timestamp falls inside the timed region are reported (nvidia-smi
    needs ~0.3 s to start, longer than a short timed region)."""

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

        def collect(lines):
            sm, mx, pw, reasons = [], [], [], set()
            for _, ln in lines:
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 8:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
                except ValueError:
                    continue
                for nme, val in zip(names, f[4:8]):
                    if val.lower().startswith("active"):
                        reasons.add(nme)
            sm.sort()
            return sm, mx, pw, reasons

        inside = [x for x in self.lines if t0 is not None and t0 <= x[0] <= t1 + 0.03]
        window = "timed region"
        if not inside:
            inside, window = self.lines, "warm-up + timed region (timed region shorter than the sampling period)"
        sm, mx, pw, reasons = collect(inside)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "window": window,
                "reasons": sorted(reasons)}


NCU_SUMMARY = ("profiles/r2g_ncu_summary.csv", "profiles/r2f_ncu_summary.csv", "profiles/r2_ncu_summary.csv", "profiles/r1_ncu_summary.csv")


def ncu_traffic_bytes(kernel_substr, stash=True):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full
    capture (C2-sized launch); (None, None) if absent."""
    import csv
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    for rel in NCU_SUMMARY:
        path = os.path.join(REPO, rel)
        if not os.path.exists(path):
            continue
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        for r in rows[2:]:
            if kernel_substr not in r[ki]:
                continue
            if kernel_substr == "mlp_fwd":      # the training step runs the stash variant: mlp_fwd_kernel<W, true, ...>
                m = re.search(r"<\s*\d+\s*,\s*(\w+)", r[ki])
                if m is None or stash != (m.group(1) in ("1", "true")):
                    continue
            return float(r[ri]) * mult.get(units[ri], 1.0) + float(r[wi]) * mult.get(units[wi], 1.0), rel
    return None, None


def l2_probe_peaks():
    p = os.path.join(REPO, "profiles", "r2_probe_l2.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def build_engine(wl, device, seed, world=1):
    from loner_b200 import engine as eng
    from loner_b200 import synth
    wc = synth.world_cube(wl["geom"])
    cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=synth.GEOMETRY[wl["geom"]]["ray_range"],
                           n_frequencies=10, n_neurons=wl["W"], n_hidden_layers=wl["L"], n_samples=wl["S"],
                           sampler="OGM", seed=seed, encoding=wl.get("encoding", "Frequency"),
                           net_flags=wl.get("net_flags"))      # kernel A/B variants (tests/gpu_hash_agg.py); None = production
    e = eng.MappingEngine(cfg, device=device)
    scans, poses = synth.make_window(wl["geom"], wl["K"], seed=0)
    for k in range(wl["K"]):
        e.add_keyframe(scans[k].ray_directions, scans[k].distances, synth.axis_angle_from_yaw_pose(poses[k]))
    e.grid.copy_(synth.trained_occupancy_grid(wl["geom"])[0, 0])
    e.new_phase(optimize_poses=wl["poses"])
    return e


def best_cpu_threads(wl):
    """torch's CPU kernels stop scaling (and regress) long before 128 threads on these op sizes:
    time one small iteration at a few thread counts and keep the fastest, so that the CPU arm is
    the best the host can do, with the thread count reported as `cores`."""
    ncpu = os.cpu_count() or 1
    cands = sorted({c for c in (8, 16, 32, 64, ncpu) if c <= ncpu})
    best, best_t = cands[0], float("inf")
    for c in cands:
        torch.set_num_threads(c)
        _, med = cpu_port_rays_per_sec(wl, 64, 1, 1, tune=False)
        if med < best_t:
            best, best_t = c, med
    torch.set_num_threads(best)
    return best


def cpu_port_rays_per_sec(wl, n_rays, iters, warmup, tune=True):
    """The oracle (CPU restatement of the reference's path) timed on the host cores: full iteration
    incl. backward and Adam on `n_rays` rays of the same workload."""
    from loner_b200 import synth
    from oracle import loner_oracle as orc
    from oracle import tcnn_standin
    if tune:
        best_cpu_threads(wl)
    wc = synth.world_cube(wl["geom"])
    K = wl["K"]
    scans, poses = synth.make_window(wl["geom"], K, seed=0)
    poses6 = [synth.axis_angle_from_yaw_pose(poses[k]) for k in range(K)]
    hash_spec = None
    if wl.get("encoding") == "HashGrid":
        from oracle import hashgrid_standin
        hash_spec = hashgrid_standin.HashGridSpec()
    spec = orc.NetSpec(10, wl["W"], wl["L"], "fp32", hash=hash_spec)
    params = tcnn_standin.xavier_uniform_flat(spec.shapes, 1337)
    if hash_spec is not None:
        params = torch.cat([params, hashgrid_standin.init_table(hash_spec, 1338)])
    params.requires_grad_(True)
    m, v = torch.zeros_like(params), torch.zeros_like(params)
    grid = synth.trained_occupancy_grid(wl["geom"])
    shift = torch.tensor(wc.shift)
    S = wl["S"]
    n_per = max(n_rays // K, 1)
    g = torch.Generator().manual_seed(0)
    times = []
    for it in range(warmup + iters):
        t0 = time.perf_counter()
        idx = [torch.randint(0, scans[k].distances.shape[0], (n_per,), generator=g) for k in range(K)]
        n = n_per * K
        u1, u2 = torch.rand(n, S // 2, generator=g), torch.rand(n, S // 2, generator=g)
        noise = torch.randn(n, S, generator=g)
        params.grad = None
        rays, depths, res, out = orc.mapping_iteration(scans, poses6, idx, params, spec, grid, S, wc.scale_factor,
                                                       shift, synth.GEOMETRY[wl["geom"]]["ray_range"], 1.0, u1, u2,
                                                       noise, orc.LossCfg())
        out["loss"].backward()
        with torch.no_grad():
            p, m, v = orc.adam_update(params.detach(), params.grad, m, v, it + 1, 0.01)
            params.data.copy_(p)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    times.sort()
    med = times[len(times) // 2]
    return n_per * K / med, med


# ---------------------------------------------------------------------------------------------- timing helpers
class Timing:
    def __init__(self, dev, world):
        self.dev, self.world = dev, world

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed calls, then EXACTLY `steps` calls between two CUDA events on the launching stream, bracketed by
        barrier + synchronize on both sides; returns (max-over-ranks ms for all steps, wall clock window)."""
        for _ in range(warmup):
            fn()
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        w0 = time.time()
        a.record()
        for _ in range(steps):
            fn()
        b.record()
        self.barrier()
        w1 = time.time()
        ms = torch.tensor([a.elapsed_time(b)], device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), (w0, w1)


def run_workload(wl, tm, rank, steps, warmup, sections=False):
    """Times the engine step of one workload; returns dict(value rays/s over all ranks, ms_per_step, ...)."""
    e = build_engine(wl, tm.dev, seed=1000 + rank, world=tm.world)
    window = list(range(wl["K"]))
    n_per_kf = rays_per_gpu(wl, tm.world) // wl["K"]
    N = n_per_kf * wl["K"]
    for _ in range(max(warmup, 3)):
        e.step(window, n_per_kf, optimize_poses=wl["poses"])
    if sections:
        tm.barrier()
        e.timers = {}
    l0 = e.launches
    ms, wall = tm.timed(lambda: e.step(window, n_per_kf, optimize_poses=wl["poses"]), steps, 0)
    launches = e.launches - l0
    secs = {}
    if sections:
        secs = {k: (sum(a.elapsed_time(b) for a, b in v) / len(v), len(v)) for k, v in e.timers.items()}
        e.timers = None
    return dict(engine=e, N=N, n_per_kf=n_per_kf, window=window, ms_per_step=ms / steps, wall=wall, launches=launches,
                value=N * tm.world / (ms / steps * 1e-3), sections=secs)


def kernel_table(wl, secs, per_chunk, peaks, net_flags):
    """Per-section roofline entries: what bounds the kernel family, its algorithmic work per launch, the achieved rate
    and the fraction of the measured peak."""
    W, L, S = wl["W"], wl["L"], wl["S"]
    e_pad, f_fwd, _ = flops_per_sample(wl)
    hashed = wl.get("encoding") == "HashGrid"
    nb = max(W, 128) // 64
    gen = not (net_flags & 2)
    spec = {
        "sample": ("hbm", None, 4.0, "writes z [N,S]; the 4 MB occupancy grid is gathered from L2"),
        "render_loss": ("hbm", None, 12.0, "reads sigma and z, writes d_sigma (noise from Philox)"),
        "ogm_update": ("hbm", None, 4.0, "reads z; trilinear scatter into the L2-resident grid (atomics), every 10th step"),
        "adam_pack": ("hbm", None, None, "217 K parameters: launch-latency bound"),
    }
    if hashed:
        spec["mlp_fwd"] = ("l2-gather", 2 * (32 * 64 + 64), 512.0, "hash_fwd: 128 gathers x 4 B per sample from the L2-resident table")
        spec["mlp_dgrad"] = ("l2-atomic", 6 * (32 * 64 + 64), 1536.0, "hash_bwd: re-gathers 128 x 4 B and issues 128 x 8 B vector atomics per sample")
    else:
        gen = gen and L >= 2                  # mlp.cu wgrad_rebuilds_last
        fold = gen and not (net_flags & 128)  # mlp.cu wgrad_folds_out: A_L is neither stashed nor read (dW_out from dW_{L-1}'s partials)
        stash = 16384 * (1 + (L - (1 if fold else 0)) * nb) / 128 + L * (max(W, 128) // 32) * 4
        dz = 16384 * nb * (L - (1 if gen else 0)) / 128
        rd = (16384 * ((1 + nb + (0 if fold else nb)) + (L - 2) * 2 * nb + (nb if gen else 2 * nb)) / 128 if L >= 2
              else 16384 * (1 + nb + nb) / 128)
        spec["mlp_fwd"] = ("tensor", f_fwd, stash, "tcgen05 forward; also writes the activation stash (design traffic, bytes_per_sample)")
        spec["mlp_dgrad"] = ("tensor", 2 * ((L - 1) * W * W + (e_pad * W if wl["poses"] else 0)), dz,
                             "tcgen05 dgrad; also writes the dZ stash (design traffic)")
        spec["mlp_wgrad"] = ("hbm", 2 * (e_pad * W + (L - 1) * W * W + W), rd,
                             "tcgen05 wgrad incl. dW_out (folded into the last hidden layer's partials): streams the stash once (design traffic), accumulators in TMEM")
    out = {}
    for k, (t, cnt) in secs.items():
        bound, flops, byts, what = spec.get(k, ("hbm", None, None, ""))
        ent = {"ms": round(t, 4), "launch_groups_timed": cnt, "bound": bound, "what": what}
        if flops:
            ent["tflops"] = round(flops * per_chunk / (t * 1e-3) / 1e12, 1)
            ent["frac_of_sustained_tensor_peak"] = round(ent["tflops"] / peaks["tflops_sustained"], 4)
        if byts:
            ent["bytes_per_sample"] = round(byts, 1)
            ent["gbs"] = round(byts * per_chunk / (t * 1e-3) / 1e9, 1)
            ent["frac_of_hbm_peak"] = round(ent["gbs"] / peaks["hbm_gbs"], 4)
        out[k] = ent
    return out


class _Cfg(dict):
    def __getattr__(self, k):
        v = self[k]
        return _Cfg(v) if isinstance(v, dict) else v


def api_optimizer_rate(wl, tm, steps):
    """The reference-facing call: FusedOptimizer(settings, ...).iterate_optimizer(window, OptimizationSettings) on
    reference-shaped keyframe objects (LidarScan buffers on the HOST, poses as 6-vectors), as Mapper.update drives it
    (mapping/mapper.py:104).  Keyframe scans cross to the device once, when first seen; every call ends with the
    reference's own host reads (loss finiteness, depth_eps)."""
    from loner_b200 import synth
    from loner_b200.dropin.mapping_optimizer import FusedOptimizer, OptimizationSettings
    wc = synth.world_cube(wl["geom"])
    n_per_kf = rays_per_gpu(wl, tm.world) // wl["K"]
    hashed = wl.get("encoding") == "HashGrid"
    enc = (dict(otype="HashGrid", n_levels=16, n_features_per_level=2, log2_hashmap_size=18, base_resolution=16,
                per_level_scale=2.0) if hashed else dict(otype="Frequency", n_frequencies=10))
    st = _Cfg(freeze_poses=False, skip_pose_refinement=True, num_samples=dict(lidar=n_per_kf, sky=0),
              rays_selection=dict(strategy="RANDOM"), samples_selection=dict(strategy="OGM"),
              keyframe_schedule=(dict(num_keyframes=-1, iteration_schedule=(
                  dict(num_iterations=steps, freeze_poses=not wl["poses"], freeze_sigma_mlp=False, freeze_rgb_mlp=True),)),),
              model_config=dict(
                  model=dict(ray_range=list(synth.GEOMETRY[wl["geom"]]["ray_range"]), num_colors=3, model_type="nerf_decoupled",
                             nerf_config=dict(pos_encoding_sigma=enc, sigma_network=dict(otype="CutlassMLP", n_neurons=wl["W"],
                                                                                         n_hidden_layers=wl["L"])),
                             render=dict(N_samples_train=wl["S"], N_samples_test=2048, perturb=1.0, raw_noise_std=1.0,
                                         chunk=16384, netchunk=0, retraw=True, white_bkgd=False),
                             occ_model=dict(voxel_size=100, lr=1e-4, N_iters_acc=10)),
                  train=dict(lrate_sigma_mlp=0.01, lrate_pose=0.001, lrate_gamma=1.0),
                  loss=dict(loss_selection="L1_JS", JS_loss=dict(min_js_score=1.0, max_js_score=10.0, alpha=1.0),
                            decay_los_lambda=False, los_lambda=1000.0, min_depth_eps=0.5, depthloss_lambda=0.005)))
    world_cube = types.SimpleNamespace(scale_factor=torch.tensor(wc.scale_factor), shift=torch.tensor(wc.shift))
    opt = FusedOptimizer(st, None, world_cube, tm.dev.index, False, True, False)
    scans, poses = synth.make_window(wl["geom"], wl["K"], seed=0)

    class Pose:
        def __init__(self, p6):
            self._t = p6.clone()

        def get_pose_tensor(self):
            return self._t

    class KF:
        def __init__(self, scan, p6, t):
            self._scan, self._pose, self._time, self.is_anchored = scan, Pose(p6), t, (t == 0.0)

        def get_lidar_scan(self):
            return self._scan

        def get_lidar_pose(self):
            return self._pose

        def get_time(self):
            return torch.tensor(self._time)

    kfs = [KF(scans[k], synth.axis_angle_from_yaw_pose(poses[k]), 3.0 * k) for k in range(wl["K"])]
    opt._engine.grid.copy_(synth.trained_occupancy_grid(wl["geom"])[0, 0])
    opt.iterate_optimizer(kfs, OptimizationSettings(num_iterations=3, freeze_poses=not wl["poses"]))   # registers the scans (H2D)
    ms, _ = tm.timed(lambda: opt.iterate_optimizer(kfs, OptimizationSettings(num_iterations=steps, freeze_poses=not wl["poses"])),
                     1, 0)
    N = n_per_kf * wl["K"]
    return {"value": N * tm.world * steps / (ms * 1e-3), "unit": "rays/s", "iterations_per_call": steps,
            "api": "FusedOptimizer.iterate_optimizer(keyframe_window, OptimizationSettings) - the reference Optimizer's entry point "
                   "(mapping/optimizer.py:144); on-device ray pick/build, loss finiteness + depth_eps read back per call",
            "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 8.0 / steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-rays", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the hash / api / c5 / c4 sub-objects")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    e_pad, f_fwd, f_train = flops_per_sample(wl)
    config = describe(wl, max(world, args.gpus) if args.impl == "reference" else world)

    if args.impl == "reference":
        # reference arm: the reference's own CPU path restated (oracle/), all host threads, bounded sample.
        # The rate is per ray: the step below runs `cpu_sample_rays` rays of the SAME workload (same scans, network,
        # samples per ray), not the full rays_per_gpu batch - a full C2 batch takes ~1 min per step on the host.
        if rank != 0:
            return
        n = args.cpu_sample_rays
        rps, med = cpu_port_rays_per_sec(wl, n, args.steps, args.warmup)
        config["cpu_sample_rays"] = n
        config["note"] = (f"CPU arm: each step = one full mapping iteration over {n} rays of this workload (bounded sample); "
                          "rays/s is per ray and is compared with the GPU arm's full-batch rate")
        line = {"impl": "reference", "metric": "rays/sec", "value": rps, "unit": "rays/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": med * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rps, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                 "sample": f"{n} rays x {wl['S']} samples per step, full iteration (fwd+bwd+Adam), "
                                           f"torch CPU fp32, best of 8/16/32/64/all threads = {torch.get_num_threads()} "
                                           f"of {os.cpu_count()} host cores"},
                "e2e": {"value": rps, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    tm = Timing(dev, world)
    from loner_b200 import ops

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    main_run = run_workload(wl, tm, rank, args.steps, args.warmup, sections=True)
    e, N, n_per_kf, window = main_run["engine"], main_run["N"], main_run["n_per_kf"], main_run["window"]
    clock_info = clocks.stop(*main_run["wall"]) if rank == 0 else None
    loss_val = float(e.step(window, n_per_kf, optimize_poses=wl["poses"]).item())

    # ---- e2e: the step fed with HOST rays/depths (pinned), H2D every step, loss read back every step
    rays_h = e.last["rays"].detach().cpu().pin_memory()
    depths_h = e.last["depths"].detach().cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 10))
    ms2, _ = tm.timed(lambda: e.step_from_host(rays_h, depths_h), e2e_steps, 3)
    e2e_rps = N * world * e2e_steps / (ms2 * 1e-3)

    # ---- forward-only (test-mode render: sampler + MLP inference + volume render), same rays
    rays_d = e.last["rays"]
    ms3, _ = tm.timed(lambda: e.render(rays_d, seed=2), e2e_steps, 3)
    fwd_ms = ms3 / e2e_steps
    net_flags = getattr(e.net, "flags", 0)
    del e
    main_run["engine"] = None
    torch.cuda.empty_cache()

    extras = {}
    if not args.no_extras and args.workload == "c2":
        # the configuration the reference ships (HashGrid + 1x64): C2 size and the reference's default operating point
        hs = {}
        for name in ("c2hash", "refdefault"):
            r = run_workload(WORKLOADS[name], tm, rank, max(5, min(args.steps, 10)), 3, sections=True)
            Ph = min(r["N"], 16384) * WORKLOADS[name]["S"]
            sec = {k: v[0] for k, v in r["sections"].items()}
            l2 = l2_probe_peaks()
            hs[name] = {"rays_per_s": r["value"], "ms_per_step": r["ms_per_step"], "config": describe(WORKLOADS[name], world),
                        "sections_ms": {k: round(v, 4) for k, v in sec.items()},
                        # algorithmic L2 operations: 128 four-byte gathers per sample (forward; the backward re-gathers them)
                        # and 128 eight-byte vector reductions per sample (backward), against the rates measured for
                        # uniformly random addresses into tables of the same size (tests/gpu_probe_l2.py)
                        "roofline": {"bound": "l2 request rate",
                                     "fwd_gathers_G_per_s": round(128 * Ph / (sec["mlp_fwd"] * 1e-3) / 1e9, 1),
                                     "bwd_gathers_plus_reductions_G_per_s": round(256 * Ph / (sec["mlp_dgrad"] * 1e-3) / 1e9, 1),
                                     "peak_random_gathers_G_per_s": l2.get("random_4B_gathers_G_per_s"),
                                     "peak_random_reductions_G_per_s": l2.get("random_v2f32_reductions_G_per_s"),
                                     "peak_source": "profiles/r2_probe_l2.json (tests/gpu_probe_l2.py on a B200 of this pool)",
                                     "note": "ray-ordered samples share cells: gathers hit L1 and the coarse-level reductions are "
                                             "aggregated per warp run, so the achieved algorithmic rates may exceed the random-address peaks"}}
            r["engine"] = None
            torch.cuda.empty_cache()
        extras["hash"] = hs
        extras["api"] = api_optimizer_rate(wl, tm, max(5, min(args.steps, 20)))
        torch.cuda.empty_cache()
        if world > 1:
            from loner_b200 import engine as eng, parallel, synth
            swl = WORKLOADS["smoke"]
            wcs = synth.world_cube(swl["geom"])
            sscans, sposes = synth.make_window(swl["geom"], 3, seed=3, n_beams=16, n_azimuth=256)

            def make(distributed):
                cfg = eng.EngineConfig(scale=wcs.scale_factor, shift=wcs.shift, ray_range=(1.0, 50.0), n_neurons=128,
                                       n_hidden_layers=2, n_samples=128, chunk_rays=256)
                en = eng.MappingEngine(cfg, device=dev, distributed=distributed)
                for k in range(3):
                    en.add_keyframe(sscans[k].ray_directions, sscans[k].distances, synth.axis_angle_from_yaw_pose(sposes[k]))
                en.grid.copy_(synth.trained_occupancy_grid(swl["geom"])[0, 0])
                return en

            extras["grad_check"] = parallel.grad_check(make, [0, 1, 2], 160, 128, optimize_poses=True)
            extras["grad_check"]["what"] = ("k-GPU sharded step vs the same global ray set on one GPU: relative difference of "
                                            "loss, MLP gradient, pose gradient (3 keyframes x 160 rays per rank, 2x128, poses on)")
            for name in (["c5"] + (["c4"] if world == 4 else [])):
                r = run_workload(WORKLOADS[name], tm, rank, max(4, min(args.steps, 8)), 3)
                _, _, ft = flops_per_sample(WORKLOADS[name])
                extras[name] = {"rays_per_s": r["value"], "ms_per_step": r["ms_per_step"],
                                "scaling": "strong" if "rays_total" in WORKLOADS[name] else "weak",
                                "config": describe(WORKLOADS[name], world),
                                "achieved_tflops_per_gpu": round(ft * r["N"] * WORKLOADS[name]["S"] / (r["ms_per_step"] * 1e-3) / 1e12, 1)}
                r["engine"] = None
                torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    ms_per_step = main_run["ms_per_step"]
    value = main_run["value"]
    P = N * wl["S"]
    per_chunk = min(N, 16384) * wl["S"]
    secs = main_run["sections"]
    kern = kernel_table(wl, secs, per_chunk, peaks, net_flags)
    hashed = wl.get("encoding") == "HashGrid"
    dom = max((k for k in secs if k.startswith("mlp_")), key=lambda k: secs[k][0])
    if hashed:
        kd = kern[dom]
        roofline = {"kernel": "hash_bwd" if dom == "mlp_dgrad" else "hash_fwd", "bound": "hbm", "achieved": kd["gbs"],
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": kd["frac_of_hbm_peak"], "traffic": None,
                    "traffic_note": "algorithmic gather + atomic bytes per launch; they are served by L2 (fp16 table 14.8 MB, fp32 "
                                    "gradient table 29.7 MB), so this fraction of the HBM copy peak is a lower bound on how busy the "
                                    "memory system is, not an HBM utilisation",
                    "peak_source": peaks["source"], "algorithmic_bytes_per_sample": kd["bytes_per_sample"],
                    "samples_per_launch": per_chunk, "kernels": kern}
    else:
        traffic, src = ncu_traffic_bytes(dom)
        kd = kern[dom]
        roofline = {"kernel": dom, "bound": "tensor", "achieved": kd["tflops"], "peak": peaks["tflops_sustained"],
                    "unit": "TFLOP/s", "frac": kd["frac_of_sustained_tensor_peak"], "traffic": traffic,
                    "traffic_note": f"DRAM bytes of one C2-sized launch of this kernel from {src} (ncu --set full); the kernel's "
                                    "algorithmic HBM need is ~8 B/sample, the rest is the activation / gradient stash of the backward "
                                    "design (DESIGN.md section 4)",
                    "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                    "algorithmic_flops_per_sample": kd and (2 * (e_pad * wl["W"] + (wl["L"] - 1) * wl["W"] ** 2 + wl["W"]) if dom != "mlp_dgrad"
                                                             else 2 * ((wl["L"] - 1) * wl["W"] ** 2 + (e_pad * wl["W"] if wl["poses"] else 0))),
                    "samples_per_launch": per_chunk,
                    "step": {"algorithmic_tflop_per_step": round(f_train * P / 1e12, 3),
                             "achieved_tflops": round(f_train * P / (ms_per_step * 1e-3) / 1e12, 1),
                             "frac_of_sustained_peak": round(f_train * P / (ms_per_step * 1e-3) / 1e12 / peaks["tflops_sustained"], 4),
                             "frac_of_burst_peak": round(f_train * P / (ms_per_step * 1e-3) / 1e12 / peaks["tflops_burst"], 4)},
                    "kernels": kern}
    line = {"metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 (fp32 accumulate, fp32 master weights; render/loss fp32)",
            "data": "synthetic", "config": config, "loss": loss_val, "clocks": clock_info,
            "e2e": {"value": e2e_rps, "unit": "rays/s", "h2d_bytes_per_step": int(rays_h.numel() * 4 + depths_h.numel() * 4),
                    "d2h_bytes_per_step": 4, "api": "MappingEngine.step_from_host(rays[N,13], depths[N]) (pinned host)"},
            "gpu_launches": main_run["launches"], "roofline": roofline,
            "forward_only": {"rays_per_s": N * world / (fwd_ms * 1e-3), "ms": fwd_ms,
                             "tflops": round(f_fwd * P / (fwd_ms * 1e-3) / 1e12, 1),
                             "frac_of_sustained_peak": round(f_fwd * P / (fwd_ms * 1e-3) / 1e12 / peaks["tflops_sustained"], 4),
                             "what": "MappingEngine.render: OGM sampler + MLP inference + volume render, no stash"}}
    line.update(extras)
    if not args.no_cpu_baseline:
        n = args.cpu_sample_rays
        rps, med = cpu_port_rays_per_sec(wl, n, 3, 1)
        line["cpu_baseline"] = {"value": rps, "unit": "rays/s", "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{n} rays x {wl['S']} samples, 3 full iterations after 1 warm-up "
                                          f"(median {med:.2f} s), oracle port of the reference path, torch CPU fp32, "
                                          f"best thread count {torch.get_num_threads()} of {os.cpu_count()} host cores"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
