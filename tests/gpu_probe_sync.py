"""Latency of the MMA issuer's synchronisation primitives (tests/probes/probe_mma.cu, probe_sync_kernel); not a test."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import build, lib as L
lib = ctypes.CDLL(build.build_probe())
lib.loner_probe_sync.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
out = torch.zeros(8, dtype=torch.int64, device="cuda")
iters = 10000
for _ in range(2):
    assert lib.loner_probe_sync(iters, out.data_ptr(), L.stream_ptr()) == 0
    torch.cuda.synchronize()
names = ["mbarrier.try_wait (completed phase)", "mbarrier.test_wait", "ld.acquire.cta.shared", "ld.volatile.shared",
         "tcgen05.fence::after_thread_sync", "mbar_wait + __syncwarp"]
print(json.dumps({n: round(int(out[i]) / iters, 1) for i, n in enumerate(names)}))
