"""Drop-in replacement for ``nr3d_lib.bindings._occ_grid`` (reference: csrc/occ_grid/src/occ_grid.cpp:22-33).

``ray_marching`` / ``batched_ray_marching`` keep the reference's signature and return list
(``[packed_info i32[R,2], t_starts f32[S,1], t_ends f32[S,1], ridx i32[S], (bidx i32[S],) gidx i32[S] | None]``).
The march and the exclusive scan run on the device; the only host synchronisation is the single read of the total sample
count that sizes the outputs (the reference also needs it, ray_marching.cu:209).  Calls whose scratch fits MARCH_SCRATCH_BYTES
march every ray once (record + compact) instead of twice (count + fill).
"""
import ctypes
import enum
import os
from typing import List, Optional

import torch

from .. import _lib


class ContractionType(enum.IntEnum):  # csrc/occ_grid/include/occ_grid/cpp_api.h:14-19
    AABB = 0
    UN_BOUNDED_TANH = 1
    UN_BOUNDED_SPHERE = 2


AABB, UN_BOUNDED_TANH, UN_BOUNDED_SPHERE = ContractionType.AABB, ContractionType.UN_BOUNDED_TANH, ContractionType.UN_BOUNDED_SPHERE


def _check(name, t, dim, dtype=None):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor")
    if not t.is_contiguous():
        raise RuntimeError(f"{name} must be contiguous")
    if t.dim() != dim:
        raise RuntimeError(f"Expected {name}.ndimension() == {dim}")
    if dtype is not None and t.dtype != dtype:
        raise RuntimeError(f"expected scalar type {dtype} for {name} but found {t.dtype}")


# Scratch budget of the single-pass marcher (16 bytes per ray and step of `max_steps`): calls that fit march every ray ONCE (record + compact),
# larger ones take the reference's two passes (count + fill).  Same outputs either way.  NR3D_B200_MARCH_SCRATCH_GB=0 forces the two passes.
MARCH_SCRATCH_BYTES = int(float(os.environ.get("NR3D_B200_MARCH_SCRATCH_GB", "4")) * 2 ** 30)


def _march(rays_o, rays_d, t_min, t_max, batch_inds, batch_data_size, n_batches, roi, grid, res, type_, step_size, max_step_size,
           dt_gamma, max_steps, return_gidx, batched):
    dev = _lib.require_cuda(rays_o, rays_d, t_min, t_max, batch_inds, roi, grid, who="ray_marching")
    lib = _lib.get_lib()
    R = rays_o.shape[0]
    grid_u8 = grid.view(torch.uint8) if grid.dtype == torch.bool else grid
    if grid_u8.dtype != torch.uint8:
        raise RuntimeError("expected scalar type Bool for grid_binary")
    common = (R, rays_o.data_ptr(), rays_d.data_ptr(), t_min.data_ptr(), t_max.data_ptr(), _lib.ptr(batch_inds), int(batch_data_size),
              int(n_batches), roi.data_ptr(), grid_u8.data_ptr(), int(res[0]), int(res[1]), int(res[2]), int(type_), float(step_size),
              float(max_step_size), float(dt_gamma), int(max_steps))
    with torch.cuda.device(dev):
        st = _lib.stream_of(dev)
        num_steps = torch.empty([R], dtype=torch.int32, device=dev)
        packed_info = torch.empty([R, 2], dtype=torch.int32, device=dev)
        total = torch.zeros([1], dtype=torch.int64, device=dev)
        rec_bytes = R * int(max_steps) * 16
        records = None
        if 0 < rec_bytes <= MARCH_SCRATCH_BYTES:
            records = torch.empty([rec_bytes], dtype=torch.uint8, device=dev)
            _lib.check(lib.nr3d_march_record(*common, num_steps.data_ptr(), records.data_ptr(), rec_bytes, st))
        elif R > 0:
            _lib.check(lib.nr3d_march_count(*common, num_steps.data_ptr(), st))
        nbytes = ctypes.c_uint64(0)
        _lib.check(lib.nr3d_march_pack(R, None, None, None, None, ctypes.byref(nbytes), None))
        ws = torch.empty([max(8, nbytes.value)], dtype=torch.uint8, device=dev)
        nbytes = ctypes.c_uint64(ws.numel())
        _lib.check(lib.nr3d_march_pack(R, num_steps.data_ptr(), packed_info.data_ptr(), total.data_ptr(), ws.data_ptr(),
                                       ctypes.byref(nbytes), st))
        S = int(total.item())
        t_starts = torch.empty([S, 1], dtype=torch.float32, device=dev)
        t_ends = torch.empty([S, 1], dtype=torch.float32, device=dev)
        ridx = torch.empty([S], dtype=torch.int32, device=dev)
        bidx = torch.empty([S], dtype=torch.int32, device=dev) if batched else None
        gidx = torch.empty([S], dtype=torch.int32, device=dev) if return_gidx else None
        if R > 0 and S > 0 and records is not None:
            _lib.check(lib.nr3d_march_compact(R, _lib.ptr(batch_inds), int(batch_data_size), records.data_ptr(), packed_info.data_ptr(),
                                              t_starts.data_ptr(), t_ends.data_ptr(), ridx.data_ptr(), _lib.ptr(bidx), _lib.ptr(gidx), st))
        elif R > 0 and S > 0:
            _lib.check(lib.nr3d_march_fill(*common, packed_info.data_ptr(), t_starts.data_ptr(), t_ends.data_ptr(), ridx.data_ptr(),
                                           _lib.ptr(bidx), _lib.ptr(gidx), st))
    if batched:
        return [packed_info, t_starts, t_ends, ridx, bidx, gidx]
    return [packed_info, t_starts, t_ends, ridx, gidx]


def ray_marching(rays_o: torch.Tensor, rays_d: torch.Tensor, t_min: torch.Tensor, t_max: torch.Tensor, roi: torch.Tensor,
                 grid_binary: torch.Tensor, type: ContractionType, step_size: float, max_step_size: float, dt_gamma: float,
                 max_steps: int, return_gidx: bool) -> List[Optional[torch.Tensor]]:
    """== ray_marching (csrc/occ_grid/src/ray_marching.cu:136-243)."""
    _check("rays_o", rays_o, 2, torch.float32); _check("rays_d", rays_d, 2, torch.float32)
    _check("t_min", t_min, 1, torch.float32); _check("t_max", t_max, 1, torch.float32)
    _check("roi", roi, 1, torch.float32); _check("grid_binary", grid_binary, 3)
    if rays_o.shape[1] != 3 or rays_d.shape[1] != 3 or roi.shape[0] != 6:
        raise RuntimeError("ray_marching: expected rays_o/rays_d of shape [n_rays, 3] and roi of shape [6]")
    return _march(rays_o, rays_d, t_min, t_max, None, 0, 1, roi, grid_binary, grid_binary.shape, type, step_size, max_step_size,
                  dt_gamma, max_steps, return_gidx, batched=False)


def batched_ray_marching(rays_o: torch.Tensor, rays_d: torch.Tensor, t_min: torch.Tensor, t_max: torch.Tensor,
                         batch_inds: Optional[torch.Tensor], batch_data_size: Optional[int], roi: torch.Tensor,
                         grid_binary: torch.Tensor, type: ContractionType, step_size: float, max_step_size: float, dt_gamma: float,
                         max_steps: int, return_gidx: bool) -> List[Optional[torch.Tensor]]:
    """== batched_ray_marching (csrc/occ_grid/src/batched_marching.cu:154-287)."""
    _check("rays_o", rays_o, 2, torch.float32); _check("rays_d", rays_d, 2, torch.float32)
    _check("t_min", t_min, 1, torch.float32); _check("t_max", t_max, 1, torch.float32)
    _check("roi", roi, 2, torch.float32); _check("grid_binary", grid_binary, 4)
    if rays_o.shape[1] != 3 or rays_d.shape[1] != 3 or roi.shape[1] != 6 or grid_binary.shape[0] != roi.shape[0]:
        raise RuntimeError("batched_ray_marching: expected rays [n_rays,3], roi [B,6], grid_binary [B,rx,ry,rz]")
    R = rays_o.shape[0]
    if batch_inds is not None:
        _check("batch_inds", batch_inds, 1, torch.int32)
        if batch_inds.shape[0] != R:
            raise RuntimeError("batched_ray_marching: batch_inds must have one entry per ray")
    bds = int(batch_data_size or 0)
    if not (bds == 0 or R % bds == 0):
        raise RuntimeError(f"batched_ray_marching: Expect nonzero `batch_data_size`={bds} to be a divisor of `n_rays`={R}")
    return _march(rays_o, rays_d, t_min, t_max, batch_inds, bds, grid_binary.shape[0], roi, grid_binary, grid_binary.shape[1:], type,
                  step_size, max_step_size, dt_gamma, max_steps, return_gidx, batched=True)


class ForestMeta:
    """Plain-Python stand-in of ``nr3d_lib.bindings._forest.ForestMeta`` (csrc/forest/forest_cpp_api.h:16-36, bound in
    csrc/forest/forest.cpp:23-34): the same read-write attributes, default-constructed empty."""

    def __init__(self):
        self.octree: Optional[torch.Tensor] = None      # uint8 [n_nodes]  kaolin SPC octree bytes
        self.exsum: Optional[torch.Tensor] = None       # int32 [n_nodes+1] exclusive sum of the child counts
        self.block_ks: Optional[torch.Tensor] = None    # int16 [n_trees, 3] integer block coordinates
        self.world_block_size = [1.0, 1.0, 1.0]
        self.world_origin = [0.0, 0.0, 0.0]
        self.resolution = [0, 0, 0]
        self.n_trees = 0
        self.level = 0
        self.level_poffset = 0
        self.continuity_enabled = True


def forest_ray_marching(forest, rays_o: torch.Tensor, rays_d: torch.Tensor, t_min: torch.Tensor, t_max: torch.Tensor,
                        seg_block_inds: torch.Tensor, seg_entries: torch.Tensor, seg_exits: torch.Tensor, seg_pack_infos: torch.Tensor,
                        grid_binary: torch.Tensor, step_size: float, max_step_size: float, dt_gamma: float, max_steps: int,
                        return_gidx: bool) -> List[Optional[torch.Tensor]]:
    """== forest_ray_marching (csrc/occ_grid/src/forest_marching.cu:152-303): returns
    ``[packed_info i32[R,2], t_starts f32[S,1], t_ends f32[S,1], ridx i32[S], blidx i32[S], gidx i32[S] | None]``.
    ``forest`` is any object with ``block_ks`` (int16 [n_trees,3]), ``world_origin`` and ``world_block_size``."""
    _check("rays_o", rays_o, 2, torch.float32); _check("rays_d", rays_d, 2, torch.float32)
    _check("t_min", t_min, 1, torch.float32); _check("t_max", t_max, 1, torch.float32)
    _check("seg_block_inds", seg_block_inds, 1, torch.int32); _check("seg_entries", seg_entries, 1, torch.float32)
    _check("seg_exits", seg_exits, 1, torch.float32); _check("seg_pack_infos", seg_pack_infos, 2, torch.int32)
    _check("grid_binary", grid_binary, 4)
    block_ks = forest.block_ks
    _check("forest.block_ks", block_ks, 2, torch.int16)
    R = rays_o.shape[0]
    if rays_o.shape != rays_d.shape or rays_o.shape[1] != 3 or t_min.shape != t_max.shape or t_min.shape[0] != R:
        raise RuntimeError("forest_ray_marching: expected rays_o/rays_d [n_rays,3] and t_min/t_max [n_rays]")
    if seg_block_inds.shape != seg_entries.shape or seg_block_inds.shape != seg_exits.shape:
        raise RuntimeError("forest_ray_marching: seg_block_inds, seg_entries and seg_exits must have the same size")
    if seg_pack_infos.shape[0] != R or seg_pack_infos.shape[1] != 2:
        raise RuntimeError("forest_ray_marching: expected seg_pack_infos of shape [n_rays, 2]")
    if block_ks.shape[1] != 3 or grid_binary.shape[0] < block_ks.shape[0]:
        raise RuntimeError("forest_ray_marching: expected forest.block_ks [n_trees,3] and grid_binary [n_trees,rx,ry,rz]")
    dev = _lib.require_cuda(rays_o, rays_d, t_min, t_max, seg_block_inds, seg_entries, seg_exits, seg_pack_infos, grid_binary, block_ks,
                            who="forest_ray_marching")
    lib = _lib.get_lib()
    grid_u8 = grid_binary.view(torch.uint8) if grid_binary.dtype == torch.bool else grid_binary
    if grid_u8.dtype != torch.uint8:
        raise RuntimeError("expected scalar type Bool for grid_binary")
    origin = (ctypes.c_float * 3)(*[float(v) for v in forest.world_origin])
    bsize = (ctypes.c_float * 3)(*[float(v) for v in forest.world_block_size])
    res = grid_binary.shape[1:]
    common = (R, rays_o.data_ptr(), rays_d.data_ptr(), t_min.data_ptr(), t_max.data_ptr(), seg_block_inds.data_ptr(), seg_entries.data_ptr(),
              seg_exits.data_ptr(), seg_pack_infos.data_ptr(), block_ks.data_ptr(), origin, bsize, grid_u8.data_ptr(), int(res[0]), int(res[1]),
              int(res[2]), float(step_size), float(max_step_size), float(dt_gamma), int(max_steps))
    with torch.cuda.device(dev):
        st = _lib.stream_of(dev)
        num_steps = torch.empty([R], dtype=torch.int32, device=dev)
        packed_info = torch.empty([R, 2], dtype=torch.int32, device=dev)
        total = torch.zeros([1], dtype=torch.int64, device=dev)
        if R > 0:
            _lib.check(lib.nr3d_forest_march_count(*common, num_steps.data_ptr(), st))
        nbytes = ctypes.c_uint64(0)
        _lib.check(lib.nr3d_march_pack(R, None, None, None, None, ctypes.byref(nbytes), None))
        ws = torch.empty([max(8, nbytes.value)], dtype=torch.uint8, device=dev)
        nbytes = ctypes.c_uint64(ws.numel())
        _lib.check(lib.nr3d_march_pack(R, num_steps.data_ptr(), packed_info.data_ptr(), total.data_ptr(), ws.data_ptr(),
                                       ctypes.byref(nbytes), st))
        S = int(total.item())
        t_starts = torch.empty([S, 1], dtype=torch.float32, device=dev)
        t_ends = torch.empty([S, 1], dtype=torch.float32, device=dev)
        ridx = torch.empty([S], dtype=torch.int32, device=dev)
        blidx = torch.empty([S], dtype=torch.int32, device=dev)
        gidx = torch.empty([S], dtype=torch.int32, device=dev) if return_gidx else None
        if R > 0 and S > 0:
            _lib.check(lib.nr3d_forest_march_fill(*common, packed_info.data_ptr(), t_starts.data_ptr(), t_ends.data_ptr(), ridx.data_ptr(),
                                                  blidx.data_ptr(), _lib.ptr(gidx), st))
    return [packed_info, t_starts, t_ends, ridx, blidx, gidx]


def march_samples(rays_o: torch.Tensor, rays_d: torch.Tensor, t_starts: torch.Tensor, t_ends: Optional[torch.Tensor], ridx: torch.Tensor):
    """One-pass post-processing of the marcher's output (not a reference export; replaces the torch composition of
    occgrid_raymarch.py:96-107): returns ``(samples [S,3], deltas [S] | None)``."""
    dev = _lib.require_cuda(rays_o, rays_d, t_starts, t_ends, ridx, who="march_samples")
    if rays_o.dtype != torch.float32 or not rays_o.is_contiguous() or not rays_d.is_contiguous() or rays_o.shape != rays_d.shape:
        raise RuntimeError("march_samples: rays_o / rays_d must be contiguous float32 [n_rays, 3]")
    S = ridx.shape[0]
    t0 = t_starts.reshape(-1)
    t1 = None if t_ends is None else t_ends.reshape(-1)
    if t0.shape[0] != S or not t0.is_contiguous() or not ridx.is_contiguous() or ridx.dtype not in (torch.int32, torch.int64):
        raise RuntimeError("march_samples: expected contiguous t_starts [S] and int32 / int64 ridx [S]")
    with torch.cuda.device(dev):
        samples = torch.empty([S, 3], dtype=torch.float32, device=dev)
        deltas = None if t1 is None else torch.empty([S], dtype=torch.float32, device=dev)
        _lib.check(_lib.get_lib().nr3d_march_samples(S, rays_o.data_ptr(), rays_d.data_ptr(), t0.data_ptr(), _lib.ptr(t1), ridx.data_ptr(),
                                                     _lib.dtype_code(ridx.dtype), samples.data_ptr(), _lib.ptr(deltas), _lib.stream_of(dev)))
    return samples, deltas
