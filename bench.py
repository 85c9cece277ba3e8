#!/usr/bin/env python
"""bench.py — rays/sec of the LONER mapping step (BASELINE.json metric) on N B200s.

  python bench.py --gpus 1 --steps 20 --warmup 5                 # our arm (hand-written sm_100a kernels)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --steps 3 --warmup 1          # reference arm: CPU port on host cores

A step = one full mapping iteration over one batch of synthetic rays: ray pick + ray build +
occupancy-guided sampling + Frequency/MLP forward + volume render + JS-margin loss + backward +
Adam (+ occupancy-grid update every 10th step), exactly the loop body of
/root/reference/src/mapping/optimizer.py:276-384.  Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

WORKLOADS = {
    # BASELINE.json configs[1]: the configuration the metric is quoted on, per GPU
    "c2": dict(geom="canteen", K=1, rays_per_gpu=8192, S=512, W=256, L=4, poses=False,
               label="C2 canteen geometry, synthetic 64x1024 scan, 8192 rays/GPU x 512 samples, 4x256 MLP (Frequency-10)"),
    "c3": dict(geom="garden", K=8, rays_per_gpu=16384, S=512, W=256, L=4, poses=False,
               label="C3 garden geometry, 8-keyframe window, 16384 rays/GPU x 512 samples, 4x256 MLP"),
    "c5": dict(geom="canteen", K=16, rays_per_gpu=32768, S=512, W=256, L=4, poses=True,
               label="C5 weak scaling, 16-keyframe window, 32768 rays/GPU x 512 samples, joint pose+map"),
    "smoke": dict(geom="canteen", K=2, rays_per_gpu=512, S=128, W=128, L=2, poses=True, label="smoke"),
    # the sigma head the reference SHIPS (cfg/nerf_config/default_nerf_hash.yaml) at the C2 size
    "c2hash": dict(geom="canteen", K=1, rays_per_gpu=8192, S=512, W=64, L=1, poses=False, encoding="HashGrid",
                   label="C2 geometry and size with the reference's shipped sigma head: HashGrid (16 levels x 2, 2^18) + 1x64 MLP"),
}


def flops_per_sample(E_pad, W, L, poses):
    f_fwd = 2 * (E_pad * W + (L - 1) * W * W + W)
    f_train = 3 * f_fwd - (0 if poses else 2 * E_pad * W)
    return f_fwd, f_train


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


def ncu_traffic_bytes(kernel_substr):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch from the committed ncu --set full
    capture (profiles/r1_ncu_summary.csv, C2 workload); None if absent."""
    import csv
    path = os.path.join(REPO, "profiles", "r1_ncu_summary.csv")
    if not os.path.exists(path):
        return None
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    ki, ri, wi = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
    mult = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    # the training step runs the stash variant of the forward kernel: mlp_fwd_kernel<W, true>
    need = [kernel_substr + "_kernel"] + ([", 1>"] if kernel_substr == "mlp_fwd" else [])
    for r in rows[2:]:
        if all(n in r[ki] for n in need):
            return float(r[ri]) * mult.get(units[ri], 1.0) + float(r[wi]) * mult.get(units[wi], 1.0)
    return None


def build_engine(wl, device, seed):
    from loner_b200 import engine as eng
    from loner_b200 import synth
    wc = synth.world_cube(wl["geom"])
    cfg = eng.EngineConfig(scale=wc.scale_factor, shift=wc.shift, ray_range=synth.GEOMETRY[wl["geom"]]["ray_range"],
                           n_frequencies=10, n_neurons=wl["W"], n_hidden_layers=wl["L"], n_samples=wl["S"],
                           sampler="OGM", seed=seed, encoding=wl.get("encoding", "Frequency"))
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-rays", type=int, default=512)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    f_fwd, f_train = flops_per_sample(64, wl["W"], wl["L"], wl["poses"])
    config = {"workload": wl["label"], "geometry": wl["geom"], "keyframes": wl["K"], "rays_per_gpu": wl["rays_per_gpu"],
              "samples_per_ray": wl["S"], "mlp": f"{wl['L']}x{wl['W']}", "encoding": "Frequency(10) -> 64",
              "pose_optimisation": wl["poses"], "sampler": "OGM", "parallelism": f"ray-sharded dp{args.gpus}",
              "l2_policy": "activation stash (>8 GB/step) streams through HBM: inputs larger than L2, no flush needed"}

    if args.impl == "reference":
        # reference arm: the reference's own CPU path restated (oracle/), all host threads, bounded sample
        if rank != 0:
            return
        n = args.cpu_sample_rays
        rps, med = cpu_port_rays_per_sec(wl, n, args.steps, args.warmup)
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
    e = build_engine(wl, dev, seed=1000 + rank)
    window = list(range(wl["K"]))
    n_per_kf = wl["rays_per_gpu"] // wl["K"]
    N = n_per_kf * wl["K"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local_rank)
    if rank == 0:
        clocks.start()
    for _ in range(max(args.warmup, 3)):
        e.step(window, n_per_kf, optimize_poses=wl["poses"])
    barrier()
    e.timers = {}
    launches0 = e.launches
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    t_start.record()
    for _ in range(args.steps):
        loss = e.step(window, n_per_kf, optimize_poses=wl["poses"])
    t_end.record()
    barrier()
    wall1 = time.time()
    clock_info = clocks.stop(wall0, wall1) if rank == 0 else None
    ms = torch.tensor([t_start.elapsed_time(t_end)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    launches = e.launches - launches0
    sections = {k: sum(a.elapsed_time(b) for a, b in v) / len(v) for k, v in e.timers.items()}
    e.timers = None
    loss_val = float(loss.item())

    # ---- e2e: reference-facing call with HOST rays/depths (pinned), H2D every step, loss read back
    rays_h = e.last["rays"].detach().cpu().pin_memory()
    depths_h = e.last["depths"].detach().cpu().pin_memory()
    for _ in range(3):
        e.step_from_host(rays_h, depths_h)
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    ta, tb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ta.record()
    for _ in range(e2e_steps):
        e.step_from_host(rays_h, depths_h)
    tb.record()
    barrier()
    ms2 = torch.tensor([ta.elapsed_time(tb)], device=dev)
    if world > 1:
        dist.all_reduce(ms2, op=dist.ReduceOp.MAX)
    e2e_rps = N * world * e2e_steps / (float(ms2.item()) * 1e-3)

    # ---- forward-only (test-mode render: sampler + MLP inference + volume render), same rays
    rays_d = e.last["rays"]
    for _ in range(3):
        e.render(rays_d, seed=1)
    barrier()
    fa, fb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fa.record()
    for _ in range(e2e_steps):
        e.render(rays_d, seed=2)
    fb.record()
    barrier()
    ms3 = torch.tensor([fa.elapsed_time(fb)], device=dev)
    if world > 1:
        dist.all_reduce(ms3, op=dist.ReduceOp.MAX)
    fwd_ms = float(ms3.item()) / e2e_steps

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peaks = load_peaks()
    ms_per_step = ms_total / args.steps
    value = N * world / (ms_per_step * 1e-3)
    P = N * wl["S"]
    per_chunk = min(N, e.cfg.chunk_rays) * wl["S"]
    kern = {}
    flops = {"mlp_fwd": f_fwd, "mlp_dgrad": 2 * ((wl["L"] - 1) * wl["W"] ** 2 + (64 * wl["W"] if wl["poses"] else 0)),
             "mlp_wgrad": 2 * (64 * wl["W"] + (wl["L"] - 1) * wl["W"] ** 2 + wl["W"])}
    for k, t in sections.items():
        kern[k] = {"ms": round(t, 4)}
        if k in flops:
            kern[k]["tflops"] = round(flops[k] * per_chunk / (t * 1e-3) / 1e12, 1)
    hashed = wl.get("encoding") == "HashGrid"
    if hashed:          # E_pad = 32, no hidden-to-hidden layers; forward and the recomputing backward
        f_fwd = 2 * (32 * 64 + 64)
        f_train = 4 * f_fwd
        flops = {"mlp_fwd": f_fwd, "mlp_dgrad": 3 * f_fwd}
    dom = max((k for k in sections if k in flops), key=lambda k: sections[k])
    achieved = flops[dom] * per_chunk / (sections[dom] * 1e-3) / 1e12
    traffic = ncu_traffic_bytes(dom) if args.workload == "c2" else None
    roofline = {"kernel": dom, "bound": "tensor", "achieved": round(achieved, 1), "peak": peaks["tflops_sustained"],
                "unit": "TFLOP/s", "frac": round(achieved / peaks["tflops_sustained"], 4), "traffic": traffic,
                "traffic_note": "DRAM bytes of one launch from profiles/r1_ncu_summary.csv (ncu --set full); the "
                                "algorithmic HBM need of this kernel is 8 B/sample, the rest is the activation "
                                "stash the backward design requires (DESIGN.md section 4)",
                "peak_source": peaks["source"] + ", sustained bf16 (kernel timed inside a long step)",
                "algorithmic_flops_per_sample": flops[dom], "samples_per_launch": per_chunk,
                "step": {"algorithmic_tflop_per_step": round(f_train * P / 1e12, 3),
                         "achieved_tflops": round(f_train * P / (ms_per_step * 1e-3) / 1e12, 1),
                         "frac_of_sustained_peak": round(f_train * P / (ms_per_step * 1e-3) / 1e12 / peaks["tflops_sustained"], 4)},
                "kernels": kern}
    if hashed:
        # gather/scatter-bound: per sample the backward re-gathers 128 x 4 B and issues 128 x 8 B vector atomics
        # (the forward gathers 128 x 4 B); the 14.8 MB fp16 table and its 29.7 MB fp32 gradient live in L2
        nbytes = {"mlp_fwd": 512, "mlp_dgrad": 1536}[dom]
        gbs = nbytes * per_chunk / (sections[dom] * 1e-3) / 1e9
        roofline = {"kernel": "hash_bwd" if dom == "mlp_dgrad" else "hash_fwd", "bound": "hbm", "achieved": round(gbs, 1),
                    "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(gbs / peaks["hbm_gbs"], 4), "traffic": None,
                    "traffic_note": "algorithmic gather + atomic bytes per launch; they are served by L2 (table 14.8 MB, "
                                    "gradient table 29.7 MB), reported against the measured HBM copy peak",
                    "peak_source": peaks["source"], "algorithmic_bytes_per_sample": nbytes, "samples_per_launch": per_chunk,
                    "kernels": kern}
    line = {"metric": "rays/sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 (fp32 accumulate, fp32 master weights; render/loss fp32)",
            "data": "synthetic", "config": config, "loss": loss_val, "clocks": clock_info,
            "e2e": {"value": e2e_rps, "unit": "rays/s", "h2d_bytes_per_step": int(rays_h.numel() * 4 + depths_h.numel() * 4),
                    "d2h_bytes_per_step": 4, "api": "MappingEngine.step_from_host(rays[N,13], depths[N]) (pinned host)"},
            "gpu_launches": launches, "roofline": roofline,
            "forward_only": {"rays_per_s": N * world / (fwd_ms * 1e-3), "ms": fwd_ms,
                             "tflops": round(f_fwd * P / (fwd_ms * 1e-3) / 1e12, 1),
                             "frac_of_sustained_peak": round(f_fwd * P / (fwd_ms * 1e-3) / 1e12 / peaks["tflops_sustained"], 4),
                             "what": "MappingEngine.render: OGM sampler + MLP inference + volume render, no stash"}}
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
