"""A/B of kernel variants selected by environment variables (not a test): runs tests/gpu_microbench.py in
sub-processes, interleaved, and prints one line per run.   python tests/gpu_ab.py 2"""
import json, os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
rounds = int(sys.argv[1]) if len(sys.argv) > 1 else 2
variants = [dict(LONER_MMA_ORDER=o) for o in ("tile", "pair")]
for r in range(rounds):
    for v in variants:
        env = dict(os.environ, MB_SHORT="1", **v)
        out = subprocess.run([sys.executable, os.path.join(here, "gpu_microbench.py")], env=env, capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        res = json.loads(line[0]) if line else {"error": out.stderr[-300:]}
        print(r, v, {k: res.get(k) for k in ("mlp_fwd_stash", "mlp_fwd_infer", "mlp_dgrad", "mlp_dgrad_dpos", "mlp_wgrad_all", "error") if k in res}, flush=True)
