"""A/B of kernel variants selected by environment variables (not a test): runs tests/gpu_microbench.py in
sub-processes and prints one line per run.   python tests/gpu_ab.py"""
import json, os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
variants = [dict(), dict(LONER_MMA_ORDER="pair"), dict(LONER_WGRAD_GEN="1"), dict()]
for v in variants:
    env = dict(os.environ, MB_SHORT="1", **v)
    out = subprocess.run([sys.executable, os.path.join(here, "gpu_microbench.py")], env=env, capture_output=True, text=True)
    line = [l for l in out.stdout.splitlines() if l.startswith("{")]
    res = json.loads(line[0]) if line else {"error": out.stderr[-300:]}
    print(json.dumps({"env": v, **res}), flush=True)
