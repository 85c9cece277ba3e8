"""tests/probes/probe_mma.cu driver (not a test): clk per tcgen05.mma (M128 N256 K16, fp16) on all SMs at once - alone
and next to weight-ring loads / epilogue-like shared-memory stores.  Nominal: 128 clk (8192 FLOP/clk/SM)."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import build, lib as L

lib = ctypes.CDLL(build.build_probe())
lib.loner_probe_mma.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
blocks = torch.cuda.get_device_properties(0).multi_processor_count
src = torch.zeros(12 * 32768, dtype=torch.uint8, device="cuda")
cyc = torch.zeros(2 * blocks, dtype=torch.int64, device="cuda")
iters = 4096
out = []
for load_kb, store_kb, lds, mode, what in (
        (0, 0, 0, 0, "MMAs alone"),
        (32, 16, 1, 0, "+ ring 32 KB + tcgen05.ld + stores of what was read per chunk time (the inference forward's traffic)"),
        (0, 0, 0, 1, "MMAs with a commit per chunk"),
        (0, 0, 0, 4, "MMAs alternating between two accumulators every 4 chunks"),
        (0, 0, 0, 2, "MMAs + 8 warps polling an mbarrier"),
        (32, 0, 0, 8, "MMAs reading B from the ring slots the bulk copies fill"),
        (32, 0, 0, 13, "ring B + commit per chunk + two accumulators"),
        (32, 0, 0, 15, "ring B + commit per chunk + two accumulators + polling warps"),
        (32, 16, 1, 13, "everything: ring B, commits, two accumulators, tcgen05.ld + stores"),
        (0, 0, 0, 0, "MMAs alone (again)")):
    for rep in range(2):
        rc = lib.loner_probe_mma(iters, load_kb, store_kb, lds, mode, src.data_ptr(), blocks, cyc.data_ptr(), L.stream_ptr())
        assert rc == 0, rc
        torch.cuda.synchronize()
    c = cyc[blocks:].float()          # duration of the MMA stream itself (the other warps are paced by the clock)
    tot = cyc[:blocks].float()
    r = dict(what=what, load_kb_per_chunk=load_kb, store_kb_per_chunk=store_kb, tmem_lds_per_chunk=lds, mode=mode, clk_per_mma_mean=round(float(c.mean()) / (iters * 4), 1),
             clk_per_mma_max=round(float(c.max()) / (iters * 4), 1), kernel_clk_per_mma=round(float(tot.mean()) / (iters * 4), 1), flop_per_clk_per_sm=round(2 * 128 * 256 * 16 * iters * 4 / float(c.mean())))
    out.append(r)
    print(json.dumps(r), flush=True)

# ---- CTA pairs: tcgen05.mma.cta_group::2 (M = 256 over two SMs)
lib.loner_probe_mma2.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
clusters = blocks // 2
cyc2 = torch.zeros(clusters, dtype=torch.int64, device="cuda")
for commit in (0, 1):
    for rep in range(2):
        rc = lib.loner_probe_mma2(iters, commit, clusters, cyc2.data_ptr(), L.stream_ptr())
        assert rc == 0, rc
        torch.cuda.synchronize()
    c = cyc2.float()
    print(json.dumps(dict(what="cta_group::2 MMAs (M256 N256 K16 over a CTA pair)" + (" with a multicast commit per chunk" if commit else ""),
                          clk_per_mma_mean=round(float(c.mean()) / (iters * 4), 1), clk_per_mma_max=round(float(c.max()) / (iters * 4), 1),
                          flop_per_clk_per_sm=round(2 * 256 * 256 * 16 * iters * 4 / float(c.mean()) / 2))), flush=True)
