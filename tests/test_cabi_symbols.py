"""CPU-only: the C-ABI library builds/loads and exports every symbol include/loner_b200.h declares."""
import ctypes
import os
import re

from loner_b200 import build, lib

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(REPO, "include", "loner_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(loner_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    build.build()
    so = ctypes.CDLL(lib.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(so, n), f"{n} declared in include/loner_b200.h but not exported"
    assert set(lib.exported_symbols()) == set(names)


def test_no_compute_entry_points_need_a_gpu_to_query_sizes():
    l = lib.load()
    assert l.loner_version() == 100
    net = lib.NetT(10, 256, 4, 0)
    assert l.loner_mlp_param_count(ctypes.byref(net)) == 217088
    assert l.loner_mlp_packed_bytes(ctypes.byref(net)) == 2 * (64 * 256 * 2 + 3 * 256 * 256 * 2) + 256 * 4
    small = lib.NetT(10, 64, 2, 0)          # BASELINE config 1 (2 x 64): runs zero-padded on the 128-wide kernels
    assert l.loner_mlp_param_count(ctypes.byref(small)) == 64 * 64 + 64 * 64 + 16 * 64
    assert l.loner_mlp_packed_bytes(ctypes.byref(small)) == 2 * (64 * 128 * 2 + 128 * 128 * 2) + 128 * 4
    bad = lib.NetT(10, 96, 1, 0)
    assert l.loner_mlp_param_count(ctypes.byref(bad)) == -1
    assert l.loner_error_string(2).decode().startswith("configuration not supported")


def test_wgrad_plan_shares_every_sm_between_the_layers():
    """loner_mlp_bwd_scratch_bytes = dZ stash + one fp32 partial per wgrad CTA (mlp.cu plan_wgrad; 148 SMs on a B200 and in
    the no-GPU fallback): the CTA counts per layer follow from it.  4 x 256, E_pad = 64."""
    l = lib.load()
    P = 8192 * 512
    dz = (P // 128) * 16384 * 4 * 4                       # L * nb column-block images of 16 KB per tile

    def partial_floats(flags):
        net = lib.NetT(10, 256, 4, flags)
        return (l.loner_mlp_bwd_scratch_bytes(ctypes.byref(net), P) - dz) // 4

    def floats(c0, c_rest):                               # layer 0: [W x E_pad] (+ a dW_out row), others [W x W]
        return c0 * 64 * 256 + c_rest * 256 * 256 + c0 * 256

    assert partial_floats(0) == floats(37, 3 * 37)        # fold: equal shares (bytes per byte in flight)
    assert partial_floats(256) == floats(25, 3 * 41)      # LONER_NET_WG_PLAN_BYTES: 5 : 8 : 8 : 8
    assert partial_floats(128) == floats(40, 3 * 36)      # LONER_NET_STASH_AL: layer 0 also streams A_L, 9 : 8 : 8 : 8
    for flags in (0, 128, 256):                            # the stash layout itself does not depend on the variant
        net = lib.NetT(10, 256, 4, flags)
        assert l.loner_mlp_act_bytes(ctypes.byref(net), P) == (P // 128) * (16384 * 17 + 4 * 128 * 8 * 4)
