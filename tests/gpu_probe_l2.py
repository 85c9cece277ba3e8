"""L2 random-reduction / random-gather throughput (the hash-grid head's access patterns); not a test.  Prints JSON."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import build, lib as L
lib = ctypes.CDLL(build.build_probe())
vp, ci, cu = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint
lib.loner_probe_atomics.argtypes = [vp, cu, ci, ci, vp]
lib.loner_probe_gather.argtypes = [vp, cu, ci, ci, vp, vp]
entries = 16 * (1 << 18)
grad = torch.zeros(entries, 2, device="cuda")
feat = torch.zeros(entries, dtype=torch.int32, device="cuda")
sink = torch.zeros(4, dtype=torch.int32, device="cuda")
blocks, per = 148 * 16, 256
n = blocks * 256 * per

def timed(fn):
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / 3

t_a = timed(lambda: lib.loner_probe_atomics(grad.data_ptr(), entries, per, blocks, L.stream_ptr()))
t_g = timed(lambda: lib.loner_probe_gather(feat.data_ptr(), entries, per, blocks, sink.data_ptr(), L.stream_ptr()))
print(json.dumps({"random_v2f32_reductions_G_per_s": round(n / t_a / 1e6, 1), "random_4B_gathers_G_per_s": round(n / t_g / 1e6, 1),
                  "table_entries": entries, "ops": n}))
