"""Timeline of the pipelined forward kernel (not a test): loads tests/probes/libloner_trace.so (mlp.cu built with
-DLONER_TRACE), runs one C2-sized launch and prints, for CTA 0, when each role reached each point of a tile-step
(clock64 cycles relative to the first event).  Roles: 0 = MMA issuer, 1 / 2 = first / last epilogue warp, 3 = producer.
    MB_STASH=0|1 (inference / training forward), MB_FLAGS (loner_net_t.flags), MB_UNITS (units printed)"""
import ctypes as C, os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from loner_b200 import build, lib as L, synth, engine as eng, ops

T = C.CDLL(build.build_trace())
vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
T.loner_mlp_packed_bytes.restype = i64; T.loner_mlp_packed_bytes.argtypes = [vp]
T.loner_mlp_act_bytes.restype = i64; T.loner_mlp_act_bytes.argtypes = [vp, i64]
T.loner_mlp_pack.restype = C.c_int; T.loner_mlp_pack.argtypes = [vp, vp, vp, vp]
T.loner_mlp_fwd.restype = C.c_int; T.loner_mlp_fwd.argtypes = [vp, vp, vp, vp, vp, i32, i64, vp, vp, vp]
T.loner_trace_setup.restype = C.c_int; T.loner_trace_setup.argtypes = [vp, C.c_uint32]

N, S, W, Lh = int(os.environ.get("MB_N", 8192)), 512, 256, 4
STASH = bool(int(os.environ.get("MB_STASH", 0)))
FLAGS = int(os.environ.get("MB_FLAGS", 0))
dev = "cuda"
net = L.NetT(10, W, Lh, FLAGS)
ref = C.byref(net)
pnet = ops.Net(10, W, Lh, flags=FLAGS)
params = eng.xavier_uniform_flat(pnet.layer_shapes(), 1337).to(dev)
packed = torch.empty(T.loner_mlp_packed_bytes(ref), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
assert T.loner_mlp_pack(ref, params.data_ptr(), packed.data_ptr(), st) == 0
wc = synth.world_cube("canteen")
g = torch.Generator().manual_seed(0)
rays = torch.zeros(N, 13)
d = torch.randn(N, 3, generator=g); d = d / d.norm(dim=1, keepdim=True)
rays[:, 3:6] = d; rays[:, 11] = 1 / wc.scale_factor; rays[:, 12] = 50 / wc.scale_factor
rays = rays.to(dev)
z = (torch.rand(N, S, generator=g).sort(dim=1).values * 0.5 + 0.01).to(dev)
P = N * S
sigma = torch.empty(P, device=dev)
acts = torch.empty(T.loner_mlp_act_bytes(ref, P), dtype=torch.uint8, device=dev) if STASH else None

def run():
    rc = T.loner_mlp_fwd(ref, packed.data_ptr(), None, rays.data_ptr(), z.data_ptr(), S, P, sigma.data_ptr(),
                         acts.data_ptr() if STASH else None, st)
    assert rc == 0, rc
    torch.cuda.synchronize()

run()
CAP = 2048
buf = torch.zeros(2 * 4 * CAP * 2, dtype=torch.int64, device=dev)
assert T.loner_trace_setup(buf.data_ptr(), CAP) == 0
run()
assert T.loner_trace_setup(None, 0) == 0
raw = buf.cpu().view(2, 4, CAP, 2)
names = {0: {0: "mma.wait_a", 1: "mma.a_ready", 2: "mma.chunk", 3: "mma.issued", 4: "mma.chunk_issued", 5: "mma.a_seen", 6: "mma.chunk_seen"},
         1: {0: "epi0.wait_acc", 1: "epi0.acc_full", 2: "epi0.drained", 3: "epi0.handoff"},
         2: {0: "epi7.wait_acc", 1: "epi7.acc_full", 2: "epi7.drained", 3: "epi7.handoff"},
         3: {0: "prod.issue"}}
UNITS = [int(u) for u in os.environ.get("MB_UNITS", "3,4").split(",")]
for cta in (0, 1):
    ev = []
    for role in range(4):
        for i in range(CAP):
            tag, clk = int(raw[cta, role, i, 0]), int(raw[cta, role, i, 1])
            if clk == 0: break
            ev.append((clk, role, tag >> 48, (tag >> 32) & 0xFFFF, (tag >> 16) & 0xFFFF, (tag >> 8) & 0xFF, tag & 0xFF))
    if not ev: continue
    ev.sort()
    sel = [e for e in ev if e[3] in UNITS]
    if not sel: continue
    t0 = sel[0][0]
    print(f"==== CTA {cta}: {len(ev)} events; units {UNITS}; clk relative to the first listed event")
    for clk, role, e, unit, layer, tile, extra in sel:
        print(f"{clk - t0:8d}  u{unit} L{layer} {'XY'[tile]}  {names[role][e]}{' c%d' % extra if (role == 0 and e in (2, 4, 6)) or role == 3 else ''}")
    # summary: per (unit, layer, tile) intervals
    idx = {(r, e, u, l, t, x): c for c, r, e, u, l, t, x in ev}
    print("---- intervals (clk): unit layer tile | a_wait = mma.a_ready - mma.wait_a | issue = mma.issued - mma.a_ready | "
          "mma_done = epi0.acc_full - mma.issued | drain = epi0.drained - epi0.acc_full | handoff->a_ready(next layer)")
    for u in UNITS:
        for l in range(Lh):
            for t in range(2):
                k = lambda r, e: idx.get((r, e, u, l, t, 0))
                a0, a1, a3 = k(0, 0), k(0, 1), k(0, 3)
                e1, e2, e3 = k(1, 1), k(1, 2), k(1, 3)
                f1, f2, f3 = k(2, 1), k(2, 2), k(2, 3)
                nxt = idx.get((0, 1, u, l + 1, t, 0))
                if None in (a0, a1, a3, e1, e2, e3): continue
                print(f"u{u} L{l} {'XY'[t]} | a_wait {a1 - a0:6d} | issue {a3 - a1:6d} | mma_done {e1 - a3:6d} | drain e0 {e2 - e1:6d} e7 {(f2 - f1) if f1 and f2 else -1:6d} | "
                      f"arrive e0 {e3 - e2:5d} | handoff->a_ready {(nxt - max(e3, f3 or 0)) if nxt else -1:6d} | step {(nxt - a1) if nxt else -1:6d}")
