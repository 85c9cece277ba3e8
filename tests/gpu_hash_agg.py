"""A/B of LONER_HASH_AGG_LEVELS on the engine's hash workloads (not a test): step time and backward section."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

tm = bench.Timing(torch.device("cuda", 0), 1)
for name in ("c2hash", "refdefault"):
    for agg in (0, 2, 5, 6, 8, 0):
        wl = dict(bench.WORKLOADS[name])
        wl["net_flags"] = agg << 4
        r = bench.run_workload(wl, tm, 0, 8, 3, sections=True)
        sec = {k: round(v[0], 4) for k, v in r["sections"].items()}
        print(json.dumps(dict(workload=name, agg_levels=agg or 4, ms_per_step=round(r["ms_per_step"], 4), bwd_ms=sec.get("mlp_dgrad"))), flush=True)
