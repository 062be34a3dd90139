"""Drop-in replacement for ``nr3d_lib.bindings._lotd`` (reference: csrc/lotd/src/lotd.cpp:23-110).

Same names, argument order, defaults, validation rules, returned shapes / dtypes / stride patterns as the
reference's pybind module, implemented on top of the C-ABI library ``libnr3d_b200.so``.
"""
import ctypes
import os
import enum
import threading
from typing import List, Optional, Sequence, Tuple, Union

import torch

from .. import _lib

__all__ = ["LoDType", "InterpolationType", "LoDMeta", "lod_fwd", "lod_bwd", "lod_bwd_bwd_input", "lod_get_grid_index"]


class LoDType(enum.IntEnum):
    # values of the C++ enum incl. the un-exported VecZMatXoY=2 (csrc/lotd/include/lotd/lotd_types.h:16-26)
    Dense = 0
    VectorMatrix = 1
    CP = 3
    CPfast = 4
    NPlaneMul = 5
    NPlaneSum = 6
    Hash = 7


class InterpolationType(enum.IntEnum):
    Linear = 0
    Smoothstep = 1


# py::enum_::export_values() (lotd.cpp:68,73)
Dense, VectorMatrix, CP, CPfast, NPlaneMul, NPlaneSum, Hash = (LoDType.Dense, LoDType.VectorMatrix, LoDType.CP,
                                                              LoDType.CPfast, LoDType.NPlaneMul, LoDType.NPlaneSum, LoDType.Hash)
Linear, Smoothstep = InterpolationType.Linear, InterpolationType.Smoothstep

_TYPE_FROM_STR = {  # case-insensitive names + aliases (lotd_types.h:42-62)
    "dense": 0, "hash": 7, "nplane": 6, "nplanesum": 6, "nplanemul": 5, "vectormatrix": 1, "vm": 1,
    "veczmatxoy": 2, "cpfast": 4, "cp": 3,
}


def _string_to_lod_type(s: str) -> int:
    try:
        return _TYPE_FROM_STR[str(s).lower()]
    except KeyError:
        raise RuntimeError(f"LoTDEncoding: Invalid lod type: {s}")


class LoDMeta:
    """Host-side description of a LoTD encoding (reference: LoDMeta, lotd_torch_api.h:81-135; pybind lotd.cpp:75-109)."""

    _READONLY = ("level_res", "level_res_multidim", "level_n_params", "level_n_feats", "level_types", "level_sizes",
                 "level_types_str", "level_offsets", "map_levels", "map_cnt", "n_levels", "n_pseudo_levels",
                 "n_feat_per_pseudo_lvl", "n_dims_to_encode", "n_encoded_dims", "n_params", "interpolation_type")

    def __init__(self, n_input_dims: int, lod_res: Union[Sequence[int], Sequence[Sequence[int]]], lod_n_feats: Sequence[int],
                 lod_types: Sequence[str], hashmap_size: Optional[int] = None, use_smooth_step: Optional[bool] = None,
                 *, lod_res_multidim=None):
        if lod_res_multidim is not None:
            lod_res = lod_res_multidim
        n_input_dims = int(n_input_dims)
        lod_res = list(lod_res)
        lod_n_feats = [int(v) for v in lod_n_feats]
        lod_types = [str(t) for t in lod_types]
        if not (len(lod_res) == len(lod_n_feats) == len(lod_types)):
            raise RuntimeError("LoTDEncoding: Expect los_res, lod_n_feats, lod_str_types to have the same length")
        if n_input_dims not in (2, 3, 4):
            raise RuntimeError("LoTDEncoding: `n_input_dim` must be 2/3/4.")
        res_md = []
        for r in lod_res:
            if isinstance(r, (list, tuple)) or (hasattr(r, "__len__") and not isinstance(r, str)):
                r = [int(v) for v in r]
                if len(r) != n_input_dims:
                    raise RuntimeError("LoTDEncoding: each entry of `lod_res_multidim` must have `n_input_dims` items")
            else:
                r = [int(r)] * n_input_dims
            res_md.append(r)
        L = len(res_md)
        types = [_string_to_lod_type(t) for t in lod_types]
        flat_res = (ctypes.c_int32 * max(1, L * n_input_dims))(*[v for r in res_md for v in r])
        nf = (ctypes.c_int32 * max(1, L))(*lod_n_feats)
        tp = (ctypes.c_int32 * max(1, L))(*types)
        st = _lib.LotdMetaStruct()
        _lib.check(_lib.get_lib().nr3d_lotd_meta_create(
            n_input_dims, L, ctypes.cast(flat_res, ctypes.c_void_p), ctypes.cast(nf, ctypes.c_void_p),
            ctypes.cast(tp, ctypes.c_void_p), int(hashmap_size or 0), int(bool(use_smooth_step)), ctypes.byref(st)))
        object.__setattr__(self, "_c", st)
        D = n_input_dims
        ro = dict(
            level_res_multidim=[[int(st.level_res[l][d]) for d in range(D)] for l in range(L)],
            level_n_feats=[int(st.level_n_feats[l]) for l in range(L)],
            level_types=[int(st.level_types[l]) for l in range(L)],
            level_types_str=list(lod_types),
            level_n_params=[int(st.level_n_params[l]) for l in range(L)],
            level_sizes=[int(st.level_sizes[l]) for l in range(L)],
            level_offsets=[int(st.level_offsets[l]) for l in range(L + 1)],
            map_levels=[int(st.map_levels[p]) for p in range(st.n_pseudo_levels)],
            map_cnt=[int(st.map_cnt[p]) for p in range(st.n_pseudo_levels)],
            n_levels=int(st.n_levels), n_pseudo_levels=int(st.n_pseudo_levels),
            n_feat_per_pseudo_lvl=int(st.n_feat_per_pseudo_lvl), n_dims_to_encode=int(st.n_dims_to_encode),
            n_encoded_dims=int(st.n_encoded_dims), n_params=int(st.n_params),
            interpolation_type=InterpolationType(int(st.interpolation_type)),
        )
        ro["level_res"] = [r[0] if all(v == r[0] for v in r) else 0 for r in ro["level_res_multidim"]]
        for k, v in ro.items():
            object.__setattr__(self, k, v)
        # read-write configuration flags (lotd.cpp:105-109, defaults lotd_torch_api.h:99-103)
        object.__setattr__(self, "c_hash_only", bool(st.hash_only))
        object.__setattr__(self, "c_profile", False)
        object.__setattr__(self, "c_bmm_backend", True)
        object.__setattr__(self, "c_prefetch", True)
        object.__setattr__(self, "c_permute_dydx", True)
        # B200-only knob (no reference counterpart): walk the points in cell-sorted order (lotd_fast.cu).  On by default: whenever a call
        # is eligible (Dense/Hash-only meta, D=3, 2 / 4 / 8 features per pseudo level, fp32 points, no user `batch_offsets`) lod_fwd returns
        # contiguous row-major [N, n_enc] / [N, n_enc, D] tensors instead of the reference's transposed feature-major views; values are the
        # same.  `meta.c_sort_points = False` (or NR3D_B200_SORT_POINTS=0 in the environment) restores the reference's strides and the
        # generic kernels for callers that depend on the memory layout.
        object.__setattr__(self, "c_sort_points", os.environ.get("NR3D_B200_SORT_POINTS", "1") not in ("", "0", "false", "False"))
        object.__setattr__(self, "_ctor", (n_input_dims, res_md, lod_n_feats, lod_types, hashmap_size, use_smooth_step))

    def __setattr__(self, key, value):
        if key in self._READONLY:
            raise AttributeError(f"LoDMeta.{key} is read-only")
        object.__setattr__(self, key, bool(value) if key.startswith("c_") else value)

    def __reduce__(self):
        n, res, nf, tp, hs, ss = self._ctor
        return (LoDMeta, (n, res, nf, tp, hs, ss))

    def __repr__(self):
        return (f"LoDMeta(n_dims={self.n_dims_to_encode}, n_levels={self.n_levels}, n_encoded_dims={self.n_encoded_dims}, "
                f"n_params={self.n_params}, types={self.level_types_str})")


# --------------------------------------------------------------------------------------------------------------
# argument validation shared by all ops (reference: lotd_torch_api.cu:244-292, 412-490, 592-681, 781-825)
# --------------------------------------------------------------------------------------------------------------
def _is_forest(meta):
    return isinstance(meta, (tuple, list))


def _split_metas(lod_meta, fn, input, batch_offsets):
    """`metas=(LoDMeta, ForestMeta)` overloads (csrc/lotd/src/lotd.cpp:45-58): returns (LoDMeta, nr3d_forest_meta struct | None) after
    the reference's forest argument checks (lotd_torch_api.cu:333-359)."""
    if not _is_forest(lod_meta):
        return lod_meta, None
    if len(lod_meta) != 2:
        raise RuntimeError(f"{fn}: expected metas=(LoDMeta, ForestMeta)")
    meta, forest = lod_meta
    if not isinstance(meta, LoDMeta):
        raise RuntimeError(f"{fn}: incompatible function arguments: lod_meta must be a LoDMeta")
    if meta.n_dims_to_encode != 3:
        raise RuntimeError("LoTDEncoding::fwd: lotd-forest only supports `n_dims_to_encode`==3")
    octree, exsum, block_ks = forest.octree, forest.exsum, forest.block_ks
    for name, t, dim, dt in (("forest.octree", octree, 1, torch.uint8), ("forest.exsum", exsum, 1, torch.int32),
                             ("forest.block_ks", block_ks, 2, torch.int16)):
        if t is None or t.dim() != dim:
            raise RuntimeError(f"{fn}: Expected {dim}-dimensional tensor for argument '{name}'")
        if t.dtype != dt:
            raise RuntimeError(f"{fn}: Expected '{name}' to have scalar type {dt}, got {t.dtype}")
        if not t.is_contiguous():
            raise RuntimeError(f"{fn}: Expected contiguous tensor for argument '{name}'")
    n_trees = int(forest.n_trees)
    if tuple(block_ks.shape) != (n_trees, 3):
        raise RuntimeError(f"{fn}: Expected tensor of size [{n_trees}, 3] for argument 'forest.block_ks', got {tuple(block_ks.shape)}")
    _lib.require_cuda(input, octree, exsum, block_ks, who=fn)
    if batch_offsets is not None and tuple(batch_offsets.shape) != (n_trees,):
        raise RuntimeError(f"{fn}: Expected tensor of size [{n_trees}] for argument 'batch_offset'")
    for tp in meta.level_types:
        if tp not in (int(LoDType.Dense), int(LoDType.VectorMatrix), int(LoDType.NPlaneMul), int(LoDType.CP), int(LoDType.Hash)):
            raise RuntimeError("LoTDEncoding: lotd-forest supports Dense / VM / NPlaneMul / CP / Hash levels only")
    fm = _lib.ForestMetaStruct(octree.data_ptr(), exsum.data_ptr(), block_ks.data_ptr(), n_trees, int(forest.level), int(forest.level_poffset),
                               1 if getattr(forest, "continuity_enabled", True) else 0)
    fm._keepalive = (octree, exsum, block_ks)
    return meta, fm


def _check_common(fn, meta: LoDMeta, input, params, batch_inds, batch_offsets, batch_data_size):
    if not isinstance(meta, LoDMeta):
        raise TypeError(f"{fn}: `lod_meta` must be a nr3d_lib_b200 LoDMeta, got {type(meta)}")
    if input.dim() != 2:
        raise RuntimeError(f"{fn}: Expected 2-dimensional tensor, but got {input.dim()}-dimensional tensor for argument 'x'")
    if input.dtype not in (torch.float16, torch.float32):
        raise RuntimeError(f"{fn}: Expected 'x' to have scalar type Half or Float, got {input.dtype}")
    if not input.is_contiguous():
        raise RuntimeError(f"{fn}: Expected contiguous tensor for argument 'x'")
    N = input.shape[0]
    if input.shape[1] != meta.n_dims_to_encode:
        raise RuntimeError(f"{fn}: Expected tensor to have size {meta.n_dims_to_encode} at dimension 1, but got size "
                           f"{input.shape[1]} for argument 'x'")
    if params is not None:
        if params.dim() != 1:
            raise RuntimeError(f"{fn}: Expected 1-dimensional tensor, but got {params.dim()}-dimensional tensor for argument 'grid'")
        if params.dtype not in (torch.float16, torch.float32):
            raise RuntimeError(f"{fn}: Expected 'grid' to have scalar type Half or Float, got {params.dtype}")
        if not params.is_contiguous():
            raise RuntimeError(f"{fn}: Expected contiguous tensor for argument 'grid'")
        if meta.n_params == 0 or params.shape[0] % meta.n_params != 0:
            raise RuntimeError(f"LoTDEncoding::{fn}: Expect size of `params`={params.shape[0]} to be an integral multiple of "
                               f"`n_param`={meta.n_params}")
    if batch_inds is not None:
        if batch_inds.dim() != 1 or batch_inds.dtype != torch.int64 or not batch_inds.is_contiguous() or batch_inds.shape[0] != N:
            raise RuntimeError(f"{fn}: `batch_inds` must be a contiguous int64 tensor of shape [{N}]")
    if batch_offsets is not None:
        if batch_offsets.dim() != 1 or batch_offsets.dtype != torch.int64 or not batch_offsets.is_contiguous():
            raise RuntimeError(f"{fn}: `batch_offsets` must be a contiguous 1-D int64 tensor")
    bds = 0
    if batch_data_size is not None:
        bds = int(batch_data_size)
        if not (bds == 0 or (N % bds) == 0):
            raise RuntimeError(f"LoTDEncoding::{fn}: Expect nonzero `batch_data_size`={bds} to be a divisor of `batch_size`={N}")
    dev = _lib.require_cuda(input, params, batch_inds, batch_offsets, who=fn)
    return N, bds, dev


# Sorted records of the fast path.  lod_fwd and lod_bwd of one step see the same points, and the reference's autograd wrappers give us no way
# to hand the records from one call to the other, so the shim keeps ONE set of buffers per (device, stream): sorted records, scene ids and the
# sort's stateful workspace.  Whether the records still belong to the points of the current call is decided ON THE DEVICE by a fingerprint of
# the points (nr3d_lotd_sort_points, lotd_sort.cu) -- tensor identity / version counters are not trusted, so `x.data.mul_()`, raw kernels
# writing into the buffer or a re-clamped copy of the same points (lotd.py:211) all do the right thing.  Buffers are only reused on the
# stream that produced them.  Memory held: 16 B + 4 B per point and 2 x 4 B per sort bin per (device, stream); clear_sort_cache() frees it.
_sort_cache = {}
_sort_lock = threading.Lock()
_sort_calls = {}     # (device, stream) -> number of _sorted_points calls so far: any of them may have re-sorted the slot's records


def clear_sort_cache():
    """Free the sorted-record buffers (the next call sorts unconditionally)."""
    with _sort_lock:
        _sort_cache.clear()
        for k in _sort_calls:
            _sort_calls[k] += 1


def _n_scenes(meta, params):
    return params.shape[0] // meta.n_params if (params is not None and meta.n_params) else 1


def _sorted_eligible(meta, input, params, batch_inds, batch_offsets, bds):
    if not (getattr(meta, "c_sort_points", False) and meta.c_hash_only and meta.n_dims_to_encode == 3
            and meta.n_feat_per_pseudo_lvl in (2, 4, 8) and params is not None and params.dtype in (torch.float32, torch.float16)
            and input.dtype == torch.float32 and batch_offsets is None and input.shape[0] > 0):
        return False
    ns = _n_scenes(meta, params)
    if ns < 1 or ns >= 65535 or ns * meta.n_params >= 2 ** 32 or params.data_ptr() % 16 != 0:
        return False
    if ns > 1 and batch_inds is None and not bds:
        return False        # several tables but no scene assignment: the generic kernels read scene 0 like the reference
    return True


def _sorted_points(x: torch.Tensor, batch_inds: Optional[torch.Tensor] = None, bds: int = 0, n_scenes: int = 1, expect_new: bool = False,
                   coord_map=None):
    """(xs, scenes): float4 records (x, y, z, original index) in (scene, cell) order and -- for batched calls -- the uint16 scene of every
    record.  Sorts on the current stream unless the device-side fingerprint says the cached records belong to these very points.
    `expect_new` (forward calls: a step brings new points): sort unconditionally and take the fingerprint inside the histogram pass, which
    saves the separate fingerprint pass; the backward of the step then finds the records current.
    `coord_map = (scale, shift, clamp01)`: the records hold fma(x, scale, shift) (clamped to [1e-6, 1 - 1e-6] if clamp01) -- for callers whose
    points live in another box (ray samples in [-1, 1]^3) and who would otherwise spend two passes over [N, 3] on the conversion."""
    dev = x.device
    lib = _lib.get_lib()
    N = x.shape[0]
    st = _lib.stream_of(dev)
    batched = batch_inds is not None or bool(bds)
    ns = n_scenes if batched else 1
    key = (dev.index, int(st or 0))
    cfg = (N, ns, batched)
    with _sort_lock:
        ent = _sort_cache.get(key)
        force = 1 if expect_new else 0
        if ent is None or ent[0] != cfg:
            # New configuration (ray samples: every call has its own point count).  The workspace is kept while it is large enough -- only its
            # stateful head (header + counters) is put back to zero, not the hundreds of MB of rank / record scratch behind it.
            nbytes = ctypes.c_uint64(0)
            _lib.check(lib.nr3d_lotd_sort_points(N, None, _lib.ptr(batch_inds), int(bds), ns, 1, None, None, None, ctypes.byref(nbytes), None))
            xs = torch.empty([N, 4], dtype=torch.float32, device=dev)      # (x, y, z, original index bits) per sorted point
            scenes = torch.empty([N], dtype=torch.int16, device=dev) if batched else None
            ws = ent[3] if ent is not None else None
            if ws is None or ws.numel() < nbytes.value:
                ws = ent = None                                               # (release the old one first)
                _sort_cache.pop(key, None)
                ws = torch.empty([nbytes.value + nbytes.value // 8], dtype=torch.uint8, device=dev)   # 1/8 headroom for slightly larger calls
            _lib.check(lib.nr3d_lotd_sort_ws_reset(N, 1 if batch_inds is not None else 0, int(bds), ns, ws.data_ptr(), ws.numel(), st))
            ent = (cfg, xs, scenes, ws, ws.numel())
            _sort_cache[key] = ent
            force = 1
        _, xs, scenes, ws, nb = ent
        nbytes = ctypes.c_uint64(nb)
        sc, sh, cl = (1.0, 0.0, 0) if coord_map is None else (float(coord_map[0]), float(coord_map[1]), int(bool(coord_map[2])))
        _lib.check(lib.nr3d_lotd_sort_points_mapped(N, x.data_ptr(), _lib.ptr(batch_inds), int(bds), ns, force, sc, sh, cl, xs.data_ptr(),
                                                    _lib.ptr(scenes), ws.data_ptr(), ctypes.byref(nbytes), st))
        _sort_calls[key] = _sort_calls.get(key, 0) + 1
    return xs, scenes


def _records_token(x: torch.Tensor):
    """Token of the records the LAST _sorted_points call on x's device / current stream produced (take it right after that call)."""
    key = (x.device.index, int(_lib.stream_of(x.device) or 0))
    with _sort_lock:
        ent = _sort_cache.get(key)
        return None if ent is None else (key, _sort_calls.get(key, 0), ent[1].data_ptr())


def _records_if_untouched(token, dev):
    """(xs, scenes) of a token if the current stream of `dev` is the one the records were made on and NO sort call has been made on that
    (device, stream) slot since -- the records are then the token's for certain and a caller that owns its points (pipeline.
    _EncodeDensityAlpha: forward and backward of one autograd node) can skip the fingerprint pass.  None otherwise: call _sorted_points again."""
    if token is None:
        return None
    key, calls, ptr = token
    if key != (dev.index, int(_lib.stream_of(dev) or 0)):
        return None
    with _sort_lock:
        ent = _sort_cache.get(key)
        if ent is None or _sort_calls.get(key, 0) != calls or ent[1].data_ptr() != ptr:
            return None
        return ent[1], ent[2]


# Multi-GPU hook (no reference counterpart; the reference's DDP support is "discarded", nr3d_lib/config.py:74-75).  When set, an eligible
# lod_bwd scatters the FINE levels first and calls `hook(dL_dparam, begin, end)` as soon as a contiguous part [begin, end) of the table is
# final on the current stream, so that the caller (nr3d_lib_b200.dist.GradReducer) can start the all-reduce of that part while the
# remaining levels are still being scattered.  The hook is called once per part; the parts cover the whole table.
_grad_bucket_hook = None
_grad_bucket_split = 8      # pseudo levels [split, n) form the first part; 8 * 2 features = 64 B = two full sectors of a dL_dy row


def set_grad_bucket_hook(fn, split: Optional[int] = None):
    global _grad_bucket_hook, _grad_bucket_split
    _grad_bucket_hook = fn
    if split is not None:
        _grad_bucket_split = int(split)


# Second multi-GPU hook: who allocates dL/dparams.  `fn(numel, dtype, device)` must return a ZEROED contiguous 1-D tensor; a reducer that owns a
# symmetric-memory (NVLS multicast) buffer hands it out here so that the scatter writes straight into the buffer the all-reduce runs on
# (nr3d_lib_b200.dist.GradReducer(mode="symm")).  None: torch.zeros.
_param_grad_allocator = None


def set_param_grad_allocator(fn):
    global _param_grad_allocator
    _param_grad_allocator = fn


def _dydx_view(dy_dx, N, meta):
    """[N, n_enc, D] view of a dy_dx tensor in either reference layout; returns (tensor, stride_n, stride_j)."""
    D, E = meta.n_dims_to_encode, meta.n_encoded_dims
    v = dy_dx.view(N, E, D) if dy_dx.dim() != 3 else dy_dx
    if tuple(v.shape) != (N, E, D):
        raise RuntimeError(f"lod_bwd: `dy_dx` has shape {tuple(dy_dx.shape)}, expected [{N}, {E}, {D}] or [{N}, {E * D}]")
    if v.stride(2) != 1:
        v = v.contiguous()
    return v, v.stride(0), v.stride(1)


def lod_fwd(lod_meta, input: torch.Tensor, params: torch.Tensor, batch_inds: Optional[torch.Tensor] = None,
            batch_offsets: Optional[torch.Tensor] = None, batch_data_size: Optional[int] = None,
            max_level: Optional[int] = None, need_input_grad: Optional[bool] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """== lotd::torch::lod_fwd / lod_forest_fwd (csrc/lotd/src/lotd_torch_api.cu:232-395)."""
    meta, forest = _split_metas(lod_meta, "fwd", input, batch_offsets)
    N, bds, dev = _check_common("fwd", meta, input, params, batch_inds, batch_offsets, batch_data_size)
    D, E = meta.n_dims_to_encode, meta.n_encoded_dims
    max_level = meta.n_levels if max_level is None else int(max_level)
    need_input_grad = input.requires_grad if need_input_grad is None else bool(need_input_grad)
    if max_level <= -1:
        return (torch.zeros([N, E], dtype=params.dtype, device=dev), torch.zeros([N, E * D], dtype=input.dtype, device=dev))
    dy_dx = None
    ds_n = ds_f = 0
    if forest is not None:   # row-major outputs (lotd_forest.h:198-207); every element is written by the kernel
        with torch.cuda.device(dev):
            y = torch.empty([N, E], dtype=params.dtype, device=dev)
            if need_input_grad:
                dy_dx = torch.empty([N, E * D], dtype=input.dtype, device=dev)
            _lib.check(_lib.get_lib().nr3d_lotd_forest_fwd(
                ctypes.byref(meta._c), ctypes.byref(forest), _lib.dtype_code(input.dtype), _lib.dtype_code(params.dtype), N, _lib.ptr(input),
                _lib.ptr(params), _lib.ptr(batch_inds), _lib.ptr(batch_offsets), bds, max_level, y.data_ptr(), _lib.ptr(dy_dx),
                _lib.stream_of(dev)))
        return y, dy_dx
    with torch.cuda.device(dev):
        if _sorted_eligible(meta, input, params, batch_inds, batch_offsets, bds):
            # fast path: row-major y (and dy_dx: same shapes as the reference returns, contiguous instead of permuted views)
            ns = _n_scenes(meta, params)
            xs, scenes = _sorted_points(input, batch_inds, bds, ns, expect_new=True)
            y = torch.empty([N, E], dtype=params.dtype, device=dev)
            if not need_input_grad:
                _lib.check(_lib.get_lib().nr3d_lotd_fwd_sorted(
                    ctypes.byref(meta._c), _lib.dtype_code(params.dtype), N, xs.data_ptr(), _lib.ptr(scenes), ns, params.data_ptr(), max_level,
                    y.data_ptr(), E, 1, _lib.stream_of(dev)))
                return y, None
            dy_dx = torch.empty([N, E, D] if meta.c_permute_dydx else [N, E * D], dtype=input.dtype, device=dev)
            _lib.check(_lib.get_lib().nr3d_lotd_fwd_dydx_sorted(
                ctypes.byref(meta._c), _lib.dtype_code(params.dtype), N, xs.data_ptr(), _lib.ptr(scenes), ns, params.data_ptr(), max_level,
                y.data_ptr(), dy_dx.data_ptr(), _lib.stream_of(dev)))
            return y, dy_dx
        if input.dtype == torch.float16 and params.dtype != torch.float16:
            raise RuntimeError("LoTDEncoding: Input type combination not supported. Supported types are: "
                               "<input,param> -> (half, half), (float, half), (float, float)")
        # the kernels write dy_dx in fp32; half points (the <half, half, half> combination) get it converted afterwards (same strides)
        if meta.c_hash_only:
            # feature-major storage returned through a transposed view (lotd_torch_api.cu:303,317)
            y_store = torch.empty([E, N], dtype=params.dtype, device=dev)
            y, ys_n, ys_f = y_store.t(), 1, N
            if need_input_grad:
                if meta.c_permute_dydx:
                    store = torch.empty([E, N, D], dtype=torch.float32, device=dev)
                    dy_dx, ds_n, ds_f = store.permute(1, 0, 2), D, N * D
                else:
                    dy_dx = torch.empty([N, E * D], dtype=torch.float32, device=dev)
                    ds_n, ds_f = E * D, D
        else:
            y = torch.empty([N, E], dtype=params.dtype, device=dev)
            ys_n, ys_f = E, 1
            if need_input_grad:
                dy_dx = torch.empty([N, E * D], dtype=torch.float32, device=dev)
                ds_n, ds_f = E * D, D
        _lib.check(_lib.get_lib().nr3d_lotd_fwd(
            ctypes.byref(meta._c), _lib.dtype_code(input.dtype), _lib.dtype_code(params.dtype), N, _lib.ptr(input), _lib.ptr(params),
            _lib.ptr(batch_inds), _lib.ptr(batch_offsets), bds, max_level, y.data_ptr(), ys_n, ys_f,
            _lib.ptr(dy_dx), ds_n, ds_f, _lib.stream_of(dev)))
        if dy_dx is not None and input.dtype != torch.float32:
            dy_dx = dy_dx.to(input.dtype)
    return y, dy_dx


def lod_bwd(lod_meta, dL_dy: torch.Tensor, input: torch.Tensor, params: torch.Tensor, dy_dx: Optional[torch.Tensor] = None,
            batch_inds: Optional[torch.Tensor] = None, batch_offsets: Optional[torch.Tensor] = None,
            batch_data_size: Optional[int] = None, max_level: Optional[int] = None, need_input_grad: Optional[bool] = None,
            need_param_grad: Optional[bool] = None) -> Tuple[Optional[torch.Tensor], Optional[torch.Tensor]]:
    """== lotd::torch::lod_bwd / lod_forest_bwd (csrc/lotd/src/lotd_torch_api.cu:397-573)."""
    meta, forest = _split_metas(lod_meta, "bwd", input, batch_offsets)
    N, bds, dev = _check_common("bwd", meta, input, params, batch_inds, batch_offsets, batch_data_size)
    D, E = meta.n_dims_to_encode, meta.n_encoded_dims
    if dL_dy.dim() != 2 or tuple(dL_dy.shape) != (N, E):
        raise RuntimeError(f"bwd: Expected tensor of size [{N}, {E}] for argument 'dL_dy', got {tuple(dL_dy.shape)}")
    if dL_dy.dtype != params.dtype:
        raise RuntimeError(f"bwd: Expected 'dL_dy' ({dL_dy.dtype}) to have the same type as 'grid' ({params.dtype})")
    _lib.require_cuda(dL_dy, input, dy_dx, who="bwd")
    if dy_dx is not None and dy_dx.dtype != input.dtype:
        raise RuntimeError(f"bwd: Expected 'dy_dx' ({dy_dx.dtype}) to have the same type as 'x' ({input.dtype})")
    max_level = meta.n_levels if max_level is None else int(max_level)
    need_input_grad = input.requires_grad if need_input_grad is None else bool(need_input_grad)
    need_param_grad = params.requires_grad if need_param_grad is None else bool(need_param_grad)
    dL_dx = dL_dparam = None
    lib = _lib.get_lib()
    with torch.cuda.device(dev):
        if need_input_grad:
            if dy_dx is None:
                raise RuntimeError("LoTDEncoding::bwd: need `dy_dx` to comput `dL_dx`.")
            # (fp32 inside; half points get dL_dx converted on return)
            dL_dx = torch.zeros([N, D], dtype=torch.float32, device=dev) if max_level <= -1 else torch.empty([N, D], dtype=torch.float32, device=dev)
        if need_param_grad:
            dL_dparam = None
            if _param_grad_allocator is not None:
                dL_dparam = _param_grad_allocator(params.shape[0], params.dtype, dev)
            if dL_dparam is None:
                dL_dparam = torch.zeros([params.shape[0]], dtype=params.dtype, device=dev)
        if max_level <= -1:
            return (None if dL_dx is None else dL_dx.to(input.dtype)), dL_dparam
        st = _lib.stream_of(dev)
        if need_input_grad:
            dv, ds_n, ds_f = _dydx_view(dy_dx if dy_dx.dtype == torch.float32 else dy_dx.float(), N, meta)
            _lib.check(lib.nr3d_lotd_bwd_input(
                ctypes.byref(meta._c), _lib.dtype_code(input.dtype), _lib.dtype_code(params.dtype), N, dL_dy.data_ptr(),
                dL_dy.stride(0), dL_dy.stride(1), dv.data_ptr(), ds_n, ds_f, dL_dx.data_ptr(), st))
        if need_param_grad and forest is not None:
            _lib.check(lib.nr3d_lotd_forest_bwd_param(
                ctypes.byref(meta._c), ctypes.byref(forest), _lib.dtype_code(input.dtype), _lib.dtype_code(params.dtype), N, dL_dy.data_ptr(),
                dL_dy.stride(0), dL_dy.stride(1), None, input.data_ptr(), params.data_ptr(), _lib.ptr(batch_inds), _lib.ptr(batch_offsets),
                bds, max_level, dL_dparam.data_ptr(), st))
        elif need_param_grad and _sorted_eligible(meta, input, params, batch_inds, batch_offsets, bds):
            ns = _n_scenes(meta, params)
            xs, scenes = _sorted_points(input, batch_inds, bds, ns)
            P = meta.n_pseudo_levels
            hook, split = _grad_bucket_hook, _grad_bucket_split
            groups = [(0, P)]
            if hook is not None and ns == 1 and 0 < split < P and meta.map_levels[split] != meta.map_levels[split - 1]:
                groups = [(split, P), (0, split)]         # fine levels first: their part of the table is reduced while the coarse levels run
            for b, e in groups:
                _lib.check(lib.nr3d_lotd_bwd_param_sorted(
                    ctypes.byref(meta._c), _lib.dtype_code(params.dtype), N, xs.data_ptr(), _lib.ptr(scenes), ns, dL_dy.data_ptr(),
                    dL_dy.stride(0), dL_dy.stride(1), max_level, b, e, dL_dparam.data_ptr(), st))
                if hook is not None:
                    lo = meta.level_offsets[meta.map_levels[b]] if len(groups) > 1 else 0
                    hi = meta.level_offsets[meta.map_levels[e - 1] + 1] if len(groups) > 1 else dL_dparam.shape[0]
                    hook(dL_dparam, lo, hi)
        elif need_param_grad:
            _lib.check(lib.nr3d_lotd_bwd_param_scenes(
                ctypes.byref(meta._c), _lib.dtype_code(input.dtype), _lib.dtype_code(params.dtype), N, dL_dy.data_ptr(),
                dL_dy.stride(0), dL_dy.stride(1), None, input.data_ptr(), params.data_ptr(), _lib.ptr(batch_inds),
                _lib.ptr(batch_offsets), bds, params.shape[0] // meta.n_params, max_level, dL_dparam.data_ptr(), st))
    if dL_dx is not None and dL_dx.dtype != input.dtype:
        dL_dx = dL_dx.to(input.dtype)
    return dL_dx, dL_dparam


def lod_bwd_bwd_input(lod_meta, dL_ddLdx: torch.Tensor, dL_dy: torch.Tensor, input: torch.Tensor, params: torch.Tensor,
                      dy_dx: Optional[torch.Tensor] = None, batch_inds: Optional[torch.Tensor] = None,
                      batch_offsets: Optional[torch.Tensor] = None, batch_data_size: Optional[int] = None,
                      max_level: Optional[int] = None, need_dLdinput_ddLdoutput: Optional[bool] = None,
                      need_dLdinput_dparams: Optional[bool] = None, need_dLdinput_dinput: Optional[bool] = None):
    """== lotd::torch::lod_bwd_bwd_input / lod_forest_bwd_bwd_input (csrc/lotd/src/lotd_torch_api.cu:575-769) -> (dL_ddLdy, dL_dparams, dL_dx)."""
    meta, forest = _split_metas(lod_meta, "bwd_bwd_input", input, batch_offsets)
    N, bds, dev = _check_common("bwd_bwd_input", meta, input, params, batch_inds, batch_offsets, batch_data_size)
    D, E = meta.n_dims_to_encode, meta.n_encoded_dims
    if dL_ddLdx.dim() != 2 or tuple(dL_ddLdx.shape) != (N, D) or not dL_ddLdx.is_contiguous():
        raise RuntimeError(f"bwd_bwd_input: Expected contiguous tensor of size [{N}, {D}] for argument 'dL_ddLdx'")
    if dL_ddLdx.dtype != input.dtype:
        raise RuntimeError("bwd_bwd_input: Expected 'dL_ddLdx' to have the same type as 'x'")
    if dL_dy.dim() != 2 or tuple(dL_dy.shape) != (N, E):
        raise RuntimeError(f"bwd_bwd_input: Expected tensor of size [{N}, {E}] for argument 'dL_dy'")
    if dL_dy.dtype != params.dtype:
        raise RuntimeError("bwd_bwd_input: Expected 'dL_dy' to have the same type as 'grid'")
    if input.dtype != torch.float32:
        raise RuntimeError("LoTDEncoding: Input type combination not supported. Supported types are: "
                           "<input,param> -> (half, half), (float, half), (float, float)")
    _lib.require_cuda(dL_ddLdx, dL_dy, dy_dx, who="bwd_bwd_input")
    max_level = meta.n_levels if max_level is None else int(max_level)
    need_dLdy = dL_dy.requires_grad if need_dLdinput_ddLdoutput is None else bool(need_dLdinput_ddLdoutput)
    need_input = input.requires_grad if need_dLdinput_dinput is None else bool(need_dLdinput_dinput)
    need_param = params.requires_grad if need_dLdinput_dparams is None else bool(need_dLdinput_dparams)
    dL_ddLdy = dL_dx = dL_dparams = None
    with torch.cuda.device(dev):
        if need_dLdy:
            if dy_dx is None:
                raise RuntimeError("LoTDEncoding::bwd_bwd_input: need `dy_dx` to compute `dL_d(dLdy)`.")
            dL_ddLdy = torch.zeros([N, E], dtype=dL_dy.dtype, device=dev) if max_level <= -1 else torch.empty([N, E], dtype=dL_dy.dtype, device=dev)
        if need_input:
            dL_dx = torch.zeros([N, D], dtype=input.dtype, device=dev)
        if need_param:
            dL_dparams = torch.zeros([params.shape[0]], dtype=params.dtype, device=dev)
        if max_level <= -1 or not (need_dLdy or need_input or need_param):
            return dL_ddLdy, dL_dparams, dL_dx
        dv, ds_n, ds_f = (None, 0, 0)
        if need_dLdy:
            dv, ds_n, ds_f = _dydx_view(dy_dx, N, meta)
        lib, st = _lib.get_lib(), _lib.stream_of(dev)
        idt, pdt = _lib.dtype_code(input.dtype), _lib.dtype_code(params.dtype)
        if forest is not None:
            if need_dLdy:      # contraction with dy_dx: the same kernel as the single-block encoder
                _lib.check(lib.nr3d_lotd_bwd_bwd_input(
                    ctypes.byref(meta._c), idt, pdt, N, dL_ddLdx.data_ptr(), dL_dy.data_ptr(), dL_dy.stride(0), dL_dy.stride(1), input.data_ptr(),
                    params.data_ptr(), _lib.ptr(dv), ds_n, ds_f, None, None, 0, max_level, _lib.ptr(dL_ddLdy), None, None, st))
            if need_param:
                _lib.check(lib.nr3d_lotd_forest_bwd_param(
                    ctypes.byref(meta._c), ctypes.byref(forest), idt, pdt, N, dL_dy.data_ptr(), dL_dy.stride(0), dL_dy.stride(1),
                    dL_ddLdx.data_ptr(), input.data_ptr(), params.data_ptr(), _lib.ptr(batch_inds), _lib.ptr(batch_offsets), bds, max_level,
                    dL_dparams.data_ptr(), st))
            if need_input:
                _lib.check(lib.nr3d_lotd_forest_bwd_bwd_dx(
                    ctypes.byref(meta._c), ctypes.byref(forest), idt, pdt, N, dL_ddLdx.data_ptr(), dL_dy.data_ptr(), dL_dy.stride(0),
                    dL_dy.stride(1), input.data_ptr(), params.data_ptr(), _lib.ptr(batch_inds), _lib.ptr(batch_offsets), bds, max_level,
                    dL_dx.data_ptr(), st))
            return dL_ddLdy, dL_dparams, dL_dx
        dparam_generic = dL_dparams
        if need_param and _sorted_eligible(meta, input, params, batch_inds, batch_offsets, bds):
            ns = _n_scenes(meta, params)     # second-order scatter on the fast path; the two other outputs keep the generic kernels
            xs, scenes = _sorted_points(input, batch_inds, bds, ns)
            _lib.check(lib.nr3d_lotd_bwd_param2_sorted(
                ctypes.byref(meta._c), pdt, N, xs.data_ptr(), _lib.ptr(scenes), ns, dL_dy.data_ptr(), dL_dy.stride(0), dL_dy.stride(1), dL_ddLdx.data_ptr(),
                max_level, dL_dparams.data_ptr(), st))
            dparam_generic = None
        if dparam_generic is not None:      # second-order scatter with the scene count made explicit (shared-memory privatisation)
            _lib.check(lib.nr3d_lotd_bwd_param_scenes(
                ctypes.byref(meta._c), idt, pdt, N, dL_dy.data_ptr(), dL_dy.stride(0), dL_dy.stride(1), dL_ddLdx.data_ptr(), input.data_ptr(),
                params.data_ptr(), _lib.ptr(batch_inds), _lib.ptr(batch_offsets), bds, params.shape[0] // meta.n_params, max_level,
                dparam_generic.data_ptr(), st))
        if need_dLdy or need_input:
            _lib.check(lib.nr3d_lotd_bwd_bwd_input(
                ctypes.byref(meta._c), idt, pdt, N, dL_ddLdx.data_ptr(),
                dL_dy.data_ptr(), dL_dy.stride(0), dL_dy.stride(1), input.data_ptr(), params.data_ptr(), _lib.ptr(dv), ds_n, ds_f,
                _lib.ptr(batch_inds), _lib.ptr(batch_offsets), bds, max_level, _lib.ptr(dL_ddLdy), None,
                _lib.ptr(dL_dx), st))
    return dL_ddLdy, dL_dparams, dL_dx


def lod_get_grid_index(lod_meta, input: torch.Tensor, batch_inds: Optional[torch.Tensor] = None,
                       batch_offsets: Optional[torch.Tensor] = None, batch_data_size: Optional[int] = None,
                       max_level: Optional[int] = None) -> torch.Tensor:
    """== lotd::torch::lod_get_grid_index (csrc/lotd/src/lotd_torch_api.cu:771-855): int64 [N, n_enc, 2^D]."""
    if _is_forest(lod_meta):
        raise RuntimeError("LoTDEncoding::lod_get_grid_index: Not implemented for forest for now")
    meta = lod_meta
    if not input.is_contiguous():
        input = input.contiguous()
    N, bds, dev = _check_common("get_grid_index", meta, input, None, batch_inds, batch_offsets, batch_data_size)
    for tp in meta.level_types:
        if tp not in (int(LoDType.Dense), int(LoDType.Hash)):
            raise RuntimeError("LoTDEncoding::get_grid_index: Only support Dense/Hash type.")
    max_level = meta.n_levels if max_level is None else int(max_level)
    with torch.cuda.device(dev):
        out = torch.zeros([N, meta.n_encoded_dims, 1 << meta.n_dims_to_encode], dtype=torch.int64, device=dev)
        if max_level <= -1:
            return out
        _lib.check(_lib.get_lib().nr3d_lotd_grid_index(
            ctypes.byref(meta._c), _lib.dtype_code(input.dtype), N, input.data_ptr(), _lib.ptr(batch_inds),
            _lib.ptr(batch_offsets), bds, max_level, out.data_ptr(), _lib.stream_of(dev)))
    return out
