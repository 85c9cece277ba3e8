"""ctypes binding of the C ABI in include/loner_b200.h (the same stub INTEGRATION.md shows).

There is NO fallback: if libloner_b200.so is missing or a call returns non-zero this raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libloner_b200.so")

_c = ctypes
_vp, _i32, _i64, _u64, _f32 = _c.c_void_p, _c.c_int32, _c.c_int64, _c.c_uint64, _c.c_float


class NetT(_c.Structure):
    _fields_ = [("n_frequencies", _i32), ("n_neurons", _i32), ("n_hidden_layers", _i32), ("flags", _i32)]


class PickSegT(_c.Structure):
    _fields_ = [("kf", _i32), ("mode", _i32), ("base", _i64), ("size", _i64), ("map_off", _i64), ("out_begin", _i64)]


class HashNetT(_c.Structure):
    _fields_ = [("n_levels", _i32), ("n_features_per_level", _i32), ("log2_hashmap_size", _i32),
                ("base_resolution", _i32), ("per_level_scale", _f32), ("n_neurons", _i32),
                ("n_hidden_layers", _i32), ("flags", _i32)]


_SIGS = {
    "loner_version": (_c.c_int, []),
    "loner_sm_arch": (_c.c_int, []),
    "loner_error_string": (_c.c_char_p, [_c.c_int]),
    "loner_ray_pick": (_c.c_int, [_vp, _i32, _vp, _u64, _i64, _vp, _vp, _vp]),
    "loner_ray_build": (_c.c_int, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _f32, _f32, _f32, _vp, _vp, _vp, _vp, _vp]),
    "loner_ray_build_bwd": (_c.c_int, [_vp, _vp, _vp, _i64, _vp, _i32, _vp, _f32, _f32, _vp, _vp, _vp]),
    "loner_sample_uniform": (_c.c_int, [_vp, _i64, _i32, _f32, _vp, _u64, _vp, _vp]),
    "loner_sample_ogm": (_c.c_int, [_vp, _i64, _i32, _f32, _vp, _i32, _vp, _vp, _u64, _vp, _vp]),
    "loner_mlp_param_count": (_i64, [_vp]),
    "loner_mlp_packed_bytes": (_i64, [_vp]),
    "loner_mlp_act_bytes": (_i64, [_vp, _i64]),
    "loner_mlp_bwd_scratch_bytes": (_i64, [_vp, _i64]),
    "loner_mlp_pack": (_c.c_int, [_vp, _vp, _vp, _vp]),
    "loner_mlp_fwd": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _vp]),
    "loner_mlp_bwd": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _f32, _vp, _vp, _vp, _vp]),
    "loner_mlp_dgrad": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp, _f32, _vp, _vp, _vp]),
    "loner_mlp_wgrad": (_c.c_int, [_vp, _vp, _i64, _vp, _vp, _f32, _vp, _vp, _vp]),
    "loner_hash_param_count": (_i64, [_vp]),
    "loner_hash_table_entries": (_i64, [_vp]),
    "loner_hash_packed_bytes": (_i64, [_vp]),
    "loner_hash_bwd_scratch_bytes": (_i64, [_vp, _i64]),
    "loner_hash_pack": (_c.c_int, [_vp, _vp, _vp, _vp]),
    "loner_hash_fwd": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _vp]),
    "loner_hash_bwd": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp, _f32, _vp, _vp, _vp, _vp]),
    "loner_render_fwd": (_c.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _f32, _u64, _vp, _vp, _vp, _vp, _vp]),
    "loner_render_bwd": (_c.c_int, [_vp, _vp, _vp, _i64, _i32, _vp, _f32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "loner_render_loss": (_c.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _f32, _u64, _vp, _vp, _vp,
                                     _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "loner_loss_finalize": (_c.c_int, [_vp, _vp, _f32, _f32, _i32, _vp, _vp]),
    "loner_points_bwd": (_c.c_int, [_vp, _vp, _i64, _i32, _vp, _vp]),
    "loner_adam_step": (_c.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _f32, _f32, _f32, _f32, _vp]),
    "loner_ogm_grad": (_c.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _f32, _i32, _vp, _vp]),
    "loner_sgd_step": (_c.c_int, [_vp, _vp, _i64, _f32, _vp]),
    "loner_pose_matrices": (_c.c_int, [_vp, _vp, _i32, _vp, _f32, _vp, _vp, _vp]),
    "loner_pose_step": (_c.c_int, [_vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _f32, _f32, _f32, _f32, _i32, _vp, _vp]),
}

_lib = None


def load():
    """Loads the shared library; raises if it has not been built (python -m loner_b200.build)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} not found: build it with `python -m loner_b200.build` "
                               "(there is no CPU or PyTorch fallback for the hot path)")
        lib = _c.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)      # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def exported_symbols():
    return sorted(_SIGS)


def ptr(t):
    """Device (or host) pointer of a contiguous tensor, or NULL."""
    if t is None:
        return None
    assert t.is_contiguous(), "loner_b200 kernels need contiguous tensors"
    return t.data_ptr()


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


def check(code, what):
    if code != 0:
        msg = load().loner_error_string(code).decode()
        raise RuntimeError(f"{what} failed: {msg} (code {code})")


def host_floats(vals):
    arr = (_f32 * len(vals))(*[float(v) for v in vals])
    return arr
