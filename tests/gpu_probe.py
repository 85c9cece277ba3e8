"""TMEM->register bandwidth probe (tcgen05.ld.32x32b.x32), per SM, vs warps and loads in flight."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import ctypes
from loner_b200 import build, lib as L
lib = ctypes.CDLL(build.build_probe())
lib.loner_probe_tmem.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
for warps in (4, 8, 16):
    for mode in (1, 2, 4):
        iters = 200
        cyc = torch.zeros(148, dtype=torch.int64, device="cuda")
        sink = torch.zeros(148 * warps * 32, dtype=torch.int32, device="cuda")
        L.check(lib.loner_probe_tmem(warps, iters, mode, cyc.data_ptr(), sink.data_ptr(), L.stream_ptr()), "probe")
        torch.cuda.synchronize()
        c = float(cyc.float().median())
        byts = warps * iters * 16 * 32 * 32 * 4
        print(f"warps={warps:2d} loads_in_flight={mode}  {byts / c:7.1f} B/clk/SM   ({c / (iters * 16):6.1f} clk per 4 KB warp-load)")
