"""A/B of the kernel variants selected by loner_net_t.flags (not a test): runs tests/gpu_microbench.py in
sub-processes and prints one line per run.   python tests/gpu_ab.py
flags: 0 = CTA pairs + dZ_L rebuilt in wgrad + one issuing warp per tile in the training forward (production),
1 = single CTA, 2 = dZ_L stashed, 3 = round-1 pipeline, 32 = one issuing warp (forward and dgrad), 64 = one issuing warp in dgrad only,
128 = A_L stashed and dW_out accumulated from it (no fold), 256 = wgrad's layer-0 share by bytes."""
import json, os, subprocess, sys
here = os.path.dirname(os.path.abspath(__file__))
for flags in [int(f) for f in os.environ.get("AB_FLAGS", "0,256,128,32,1,2,3,256,0").split(",")]:
    env = dict(os.environ, MB_SHORT="1", MB_FLAGS=str(flags))
    try:
        out = subprocess.run([sys.executable, os.path.join(here, "gpu_microbench.py")], env=env, capture_output=True, text=True,
                             timeout=240)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        res = json.loads(line[0]) if line else {"error": (out.stdout + out.stderr)[-400:]}
    except subprocess.TimeoutExpired:
        res = {"error": "timeout"}
    print(json.dumps({"flags": flags, **res}), flush=True)
