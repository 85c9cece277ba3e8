"""HBM write bandwidth of bulk shared->global copies vs copy size and copies in flight per SM (not a test)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import build, lib as L
lib = ctypes.CDLL(build.build_probe())
lib.loner_probe_bulk_store.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p]
blocks = 148
dst = torch.empty(blocks * 96 * 1024 * 1024, dtype=torch.uint8, device="cuda")      # 14 GB: larger than L2
cyc = torch.zeros(blocks, dtype=torch.int64, device="cuda")
src = torch.zeros(416 * 1024, dtype=torch.uint8, device="cuda")
for kb, depth, load_kb in ((16, 1, 0), (16, 4, 0), (64, 1, 0), (64, 2, 0), (64, 8, 0), (64, 2, 64), (64, 2, 128), (16, 4, 16), (16, 4, 32)):
    if True:
        iters = 96 * 1024 // kb // 2
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for rep in range(2):
            a.record()
            rc = lib.loner_probe_bulk_store(dst.data_ptr(), kb * 1024, iters, depth, blocks, cyc.data_ptr(), src.data_ptr(), load_kb * 1024, L.stream_ptr())
            b.record()
            torch.cuda.synchronize()
        assert rc == 0
        ms = a.elapsed_time(b)
        print(f"copy {kb:3d} KB  in flight {depth}  + {load_kb:3d} KB of L2->smem loads per copy ->  write {blocks * iters * kb * 1024 / ms / 1e9:5.2f} TB/s"
              f"  load {blocks * iters * load_kb * 1024 / ms / 1e9:5.2f} TB/s   ({ms:.3f} ms)")
