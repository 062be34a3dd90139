"""ctypes loader of ``libnr3d_b200.so`` (the C-ABI declared in ``include/nr3d_b200.h``).

There is deliberately NO fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used by the callers only for device memory and streams; no torch type crosses this boundary.
"""
import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NR3D_B200_LIB") or os.path.join(_HERE, "lib", "libnr3d_b200.so")   # env override: A/B builds only

NR3D_MAX_LEVELS = 32
NR3D_MAX_DIMS = 4
NR3D_MAX_PSEUDO_LEVELS = 256

# dtype codes (include/nr3d_b200.h)
F32, F16, F64, I32, I64, I16, I8, U8 = range(8)
_DTYPE_CODE = {
    torch.float32: F32, torch.float16: F16, torch.float64: F64, torch.int32: I32, torch.int64: I64,
    torch.int16: I16, torch.int8: I8, torch.uint8: U8, torch.bool: U8,
}


class LotdMetaStruct(ctypes.Structure):
    """Mirror of ``nr3d_lotd_meta``."""
    _fields_ = [
        ("n_levels", ctypes.c_uint32), ("n_pseudo_levels", ctypes.c_uint32), ("n_feat_per_pseudo_lvl", ctypes.c_uint32),
        ("n_dims_to_encode", ctypes.c_uint32), ("n_encoded_dims", ctypes.c_uint32), ("n_params", ctypes.c_uint32),
        ("interpolation_type", ctypes.c_uint32), ("hash_only", ctypes.c_uint32),
        ("level_res", (ctypes.c_uint32 * NR3D_MAX_DIMS) * NR3D_MAX_LEVELS),
        ("level_n_feats", ctypes.c_uint32 * NR3D_MAX_LEVELS),
        ("level_types", ctypes.c_uint32 * NR3D_MAX_LEVELS),
        ("level_n_params", ctypes.c_uint32 * NR3D_MAX_LEVELS),
        ("level_sizes", ctypes.c_uint32 * NR3D_MAX_LEVELS),
        ("level_offsets", ctypes.c_uint32 * (NR3D_MAX_LEVELS + 1)),
        ("map_levels", ctypes.c_uint32 * NR3D_MAX_PSEUDO_LEVELS),
        ("map_cnt", ctypes.c_uint32 * NR3D_MAX_PSEUDO_LEVELS),
    ]


class ForestMetaStruct(ctypes.Structure):
    """Mirror of ``nr3d_forest_meta``."""
    _fields_ = [("octree", ctypes.c_void_p), ("exsum", ctypes.c_void_p), ("block_ks", ctypes.c_void_p), ("n_trees", ctypes.c_uint32),
                ("level", ctypes.c_uint32), ("level_poffset", ctypes.c_uint32), ("continuity_enabled", ctypes.c_uint32)]


_lock = threading.Lock()
_lib = None

_vp, _i32, _u32, _i64, _u64, _f32, _f64 = (ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32, ctypes.c_int64,
                                           ctypes.c_uint64, ctypes.c_float, ctypes.c_double)
_meta_p = ctypes.POINTER(LotdMetaStruct)
_u64_p = ctypes.POINTER(ctypes.c_uint64)
_forest_p = ctypes.POINTER(ForestMetaStruct)
_f32_3 = ctypes.POINTER(ctypes.c_float)    # 3 host floats

# name -> argtypes, exactly the declarations of include/nr3d_b200.h
SIGNATURES = {
    "nr3d_lotd_meta_create": [_i32, _i32, _vp, _vp, _vp, _u32, _i32, _meta_p],
    "nr3d_lotd_fwd": [_meta_p, _i32, _i32, _u64, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _i64, _i64, _vp, _i64, _i64, _vp],
    "nr3d_lotd_bwd_param": [_meta_p, _i32, _i32, _u64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp],
    "nr3d_lotd_bwd_param_scenes": [_meta_p, _i32, _i32, _u64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _u32, _u32, _i32, _vp, _vp],
    "nr3d_lotd_bwd_input": [_meta_p, _i32, _i32, _u64, _vp, _i64, _i64, _vp, _i64, _i64, _vp, _vp],
    "nr3d_lotd_bwd_bwd_input": [_meta_p, _i32, _i32, _u64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _i64, _i64, _vp, _vp, _u32,
                                _i32, _vp, _vp, _vp, _vp],
    "nr3d_lotd_grid_index": [_meta_p, _i32, _u64, _vp, _vp, _vp, _u32, _i32, _vp, _vp],
    "nr3d_lotd_forest_fwd": [_meta_p, _forest_p, _i32, _i32, _u64, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp, _vp],
    "nr3d_lotd_forest_bwd_param": [_meta_p, _forest_p, _i32, _i32, _u64, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp],
    "nr3d_lotd_forest_bwd_bwd_dx": [_meta_p, _forest_p, _i32, _i32, _u64, _vp, _vp, _i64, _i64, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp],
    "nr3d_lotd_sort_points": [_u64, _vp, _vp, _u32, _u32, _i32, _vp, _vp, _vp, _u64_p, _vp],
    "nr3d_lotd_sort_set_two_level_min": [_u64],
    "nr3d_lotd_sort_ws_reset": [_u64, _i32, _u32, _u32, _vp, _u64, _vp],
    "nr3d_lotd_sort_points_mapped": [_u64, _vp, _vp, _u32, _u32, _i32, _f32, _f32, _i32, _vp, _vp, _vp, _u64_p, _vp],
    "nr3d_lotd_fwd_sorted": [_meta_p, _i32, _u64, _vp, _vp, _u32, _vp, _i32, _vp, _i64, _i64, _vp],
    "nr3d_lotd_bwd_param_sorted": [_meta_p, _i32, _u64, _vp, _vp, _u32, _vp, _i64, _i64, _i32, _u32, _u32, _vp, _vp],
    "nr3d_lotd_fwd_dydx_sorted": [_meta_p, _i32, _u64, _vp, _vp, _u32, _vp, _i32, _vp, _vp, _vp],
    "nr3d_lotd_bwd_param2_sorted": [_meta_p, _i32, _u64, _vp, _vp, _u32, _vp, _i64, _i64, _vp, _i32, _vp, _vp],
    "nr3d_lotd_density_head_fwd_sorted": [_meta_p, _i32, _u64, _vp, _vp, _u32, _vp, _i32, _vp, _f32, _vp, _vp, _vp],
    "nr3d_lotd_density_head_bwd_sorted": [_meta_p, _i32, _u64, _vp, _vp, _u32, _vp, _vp, _vp, _vp, _f32, _i32, _vp, _vp],
    "nr3d_march_count": [_u64, _vp, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _u32,
                         _vp, _vp],
    "nr3d_march_pack": [_u64, _vp, _vp, _vp, _vp, _u64_p, _vp],
    "nr3d_march_fill": [_u64, _vp, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _u32,
                        _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_march_record": [_u64, _vp, _vp, _vp, _vp, _vp, _u32, _i32, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _f32, _u32,
                          _vp, _vp, _u64, _vp],
    "nr3d_march_compact": [_u64, _vp, _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_forest_march_count": [_u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32_3, _f32_3, _vp, _i32, _i32, _i32, _f32, _f32, _f32,
                                _u32, _vp, _vp],
    "nr3d_forest_march_fill": [_u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _f32_3, _f32_3, _vp, _i32, _i32, _i32, _f32, _f32, _f32,
                               _u32, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_march_samples": [_u64, _vp, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
    "nr3d_density_alpha_fwd": [_u64, _u32, _vp, _i64, _vp, _f32, _vp, _vp, _vp],
    "nr3d_density_alpha_bwd": [_u64, _u32, _vp, _vp, _vp, _vp, _vp, _f32, _vp, _vp],
    "nr3d_pack_sum": [_i32, _u64, _u32, _u64, _vp, _vp, _vp, _vp],
    "nr3d_pack_cumsum": [_i32, _u64, _u32, _u64, _vp, _vp, _i32, _i32, _vp, _vp],
    "nr3d_pack_cumprod": [_i32, _u64, _u32, _u64, _vp, _vp, _i32, _i32, _i32, _vp, _vp],
    "nr3d_pack_diff": [_i32, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_pack_backward_diff": [_i32, _u64, _u32, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_pack_binary": [_i32, _i32, _u64, _u32, _vp, _vp, _vp, _vp, _vp],
    "nr3d_pack_weighted_sums_fwd": [_u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_pack_weighted_sums_bwd": [_u64, _u64, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_pack_alpha_to_vw_fwd": [_i32, _u64, _u64, _vp, _vp, _f32, _f32, _vp, _vp, _vp, _vp],
    "nr3d_pack_alpha_to_vw_bwd": [_i32, _u64, _u64, _vp, _vp, _vp, _vp, _f32, _f32, _vp, _vp],
    "nr3d_pack_infos_from_counts": [_u64, _vp, _vp, _vp, _vp, _u64_p, _vp],
    "nr3d_pack_interleave_linstep": [_i32, _u64, _vp, _vp, _vp, _f64, _f64, _vp, _vp, _vp],
    "nr3d_pack_sample_step_count": [_i32, _u64, _vp, _vp, _u32, _f64, _f64, _f64, _vp, _vp],
    "nr3d_pack_sample_step_fill": [_i32, _u64, _vp, _vp, _f64, _f64, _f64, _vp, _vp, _vp, _vp],
    "nr3d_pack_mark_boundaries": [_i32, _u64, _vp, _vp, _vp],
    "nr3d_pack_searchsorted": [_i32, _u64, _vp, _vp, _vp, _u32, _vp, _vp, _vp],
    "nr3d_pack_invert_cdf": [_i32, _u64, _vp, _vp, _vp, _vp, _u32, _vp, _vp, _vp],
    "nr3d_pack_merge_sorted_aligned": [_i32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_pack_sort": [_i32, _u64, _vp, _vp, _vp, _vp],
    "nr3d_pack_matmul": [_i32, _u64, _u32, _u32, _vp, _vp, _vp, _vp, _vp],
    "nr3d_lotd_fused_density_fwd": [_vp, _u64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
    "nr3d_lotd_fused_density_bwd": [_vp, _u64, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp],
    "nr3d_occ_scatter_max": [_u64, _vp, _vp, _vp, _u64, _vp, _u32, _vp, _vp, _vp],
    "nr3d_occ_apply": [_u64, _vp, _vp, _f32, _i32, _f32, _vp, _vp, _vp],
    "nr3d_occ_binarize": [_u64, _vp, _f32, _i32, _f32, _vp, _vp, _vp],
    "nr3d_occ_sample_in_voxels": [_u64, _vp, _vp, _u64, _vp, _vp, _vp, _vp, _vp],
    "nr3d_occ_query": [_u64, _vp, _vp, _u64, _u32, _vp, _vp, _vp, _vp],
    "nr3d_pack_seg_sample_count": [_i32, _u64, _vp, _vp, _vp, _vp, _vp, _u32, _f64, _f64, _f64, _vp, _vp],
    "nr3d_pack_seg_sample_fill": [_i32, _u64, _vp, _vp, _vp, _vp, _vp, _vp, _f64, _f64, _f64, _vp, _vp, _vp, _vp, _vp],
    "nr3d_pack_mark_consecutive_segments": [_u64, _vp, _vp, _vp, _i32, _vp, _vp, _vp],
}


def get_lib():
    """Load (once) and return the ctypes handle.  Raises ImportError if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"nr3d_lib_b200: {LIB_PATH} is missing. Build it with `python nr3d_lib_b200/csrc/build.py` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU / PyTorch fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        lib.nr3d_last_error.restype = ctypes.c_char_p
        lib.nr3d_last_error.argtypes = []
        lib.nr3d_version.restype = ctypes.c_int
        lib.nr3d_launch_count.restype = ctypes.c_uint64
        for name, argtypes in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def check(rc):
    if rc != 0:
        msg = get_lib().nr3d_last_error()
        raise RuntimeError(msg.decode("utf-8", "replace") if msg else f"nr3d_b200 call failed with status {rc}")


def ptr(t):
    return None if t is None else t.data_ptr()


def dtype_code(dt):
    try:
        return _DTYPE_CODE[dt]
    except KeyError:
        raise RuntimeError(f"nr3d_lib_b200: unsupported dtype {dt}")


def require_cuda(*tensors, who=""):
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(f"{who}: expected all tensors to be CUDA tensors (the B200 build has no CPU path)")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"{who}: expected all tensors on the same GPU, got {dev} and {t.device}")
    return dev


def stream_of(device):
    return torch.cuda.current_stream(device).cuda_stream


def launch_count():
    return int(get_lib().nr3d_launch_count())
