"""ctypes front-end of oracle/march_oracle.c -- TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "march_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libmarch_oracle.so")
_lib = None


def build(force=False):
    """gcc -O2 -ffp-contract=off (no compiler-introduced FMAs; the oracle spells out the one that matters)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


def _get():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def ray_marching(rays_o, rays_d, t_min, t_max, roi, grid, contraction=0, step_size=1e-3, max_step_size=1e10, dt_gamma=0.0,
                 max_steps=512, batch_inds=None, batch_data_size=0):
    """Returns dict(packed_info int32 [R,2], t_starts, t_ends float32 [S], ridx, bidx, gidx int32 [S]).

    ``grid`` is bool/uint8 [rx,ry,rz] (single) or [B,rx,ry,rz] (batched; roi then is [B,6])."""
    lib = _get()
    rays_o = np.ascontiguousarray(rays_o, dtype=np.float32)
    rays_d = np.ascontiguousarray(rays_d, dtype=np.float32)
    t_min = np.ascontiguousarray(t_min, dtype=np.float32)
    t_max = np.ascontiguousarray(t_max, dtype=np.float32)
    roi = np.ascontiguousarray(roi, dtype=np.float32)
    grid = np.ascontiguousarray(grid).astype(np.uint8)
    res = grid.shape[-3:]
    bi = None if batch_inds is None else np.ascontiguousarray(batch_inds, dtype=np.int32)
    R = rays_o.shape[0]
    num_steps = np.zeros(R, dtype=np.int32)
    common = [ctypes.c_uint64(R), _p(rays_o), _p(rays_d), _p(t_min), _p(t_max), _p(bi), ctypes.c_uint32(int(batch_data_size)),
              _p(roi), _p(grid), ctypes.c_int(res[0]), ctypes.c_int(res[1]), ctypes.c_int(res[2]), ctypes.c_int(int(contraction)),
              ctypes.c_float(step_size), ctypes.c_float(max_step_size), ctypes.c_float(dt_gamma)]
    lib.march_oracle_count(*common, ctypes.c_uint32(int(max_steps)), _p(num_steps))
    cum = np.cumsum(num_steps.astype(np.int64))
    packed_info = np.stack([cum - num_steps, num_steps], 1).astype(np.int32)
    S = int(cum[-1]) if R else 0
    t_starts = np.zeros(S, dtype=np.float32)
    t_ends = np.zeros(S, dtype=np.float32)
    ridx = np.zeros(S, dtype=np.int32)
    bidx = np.zeros(S, dtype=np.int32)
    gidx = np.zeros(S, dtype=np.int32)
    packed_info = np.ascontiguousarray(packed_info)
    lib.march_oracle_fill(*common, _p(packed_info), _p(t_starts), _p(t_ends), _p(ridx), _p(bidx), _p(gidx))
    return dict(packed_info=packed_info, t_starts=t_starts, t_ends=t_ends, ridx=ridx, bidx=bidx, gidx=gidx)


def forest_ray_marching(rays_o, rays_d, t_min, t_max, seg_block_inds, seg_entries, seg_exits, seg_pack_infos, block_ks, world_origin,
                        world_block_size, grid, step_size=1e-3, max_step_size=1e10, dt_gamma=0.0, max_steps=512):
    """Forest (multi-block) marcher, csrc/occ_grid/src/forest_marching.cu.  ``grid`` bool/uint8 [n_trees,rx,ry,rz], ``block_ks`` int16
    [n_trees,3].  Returns dict(packed_info int32 [R,2], t_starts, t_ends float32 [S], ridx, blidx, gidx int32 [S])."""
    lib = _get()
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    rays_o, rays_d, t_min, t_max, seg_entries, seg_exits = map(f32, (rays_o, rays_d, t_min, t_max, seg_entries, seg_exits))
    world_origin, world_block_size = f32(world_origin), f32(world_block_size)
    seg_block_inds = np.ascontiguousarray(seg_block_inds, dtype=np.int32)
    seg_pack_infos = np.ascontiguousarray(seg_pack_infos, dtype=np.int32)
    block_ks = np.ascontiguousarray(block_ks, dtype=np.int16)
    grid = np.ascontiguousarray(grid).astype(np.uint8)
    res = grid.shape[-3:]
    R = rays_o.shape[0]
    num_steps = np.zeros(R, dtype=np.int32)
    common = [ctypes.c_uint64(R), _p(rays_o), _p(rays_d), _p(t_min), _p(t_max), _p(seg_block_inds), _p(seg_entries), _p(seg_exits),
              _p(seg_pack_infos), _p(block_ks), _p(world_origin), _p(world_block_size), _p(grid), ctypes.c_int(res[0]), ctypes.c_int(res[1]),
              ctypes.c_int(res[2]), ctypes.c_float(step_size), ctypes.c_float(max_step_size), ctypes.c_float(dt_gamma)]
    lib.forest_march_oracle_count(*common, ctypes.c_uint32(int(max_steps)), _p(num_steps))
    cum = np.cumsum(num_steps.astype(np.int64))
    packed_info = np.ascontiguousarray(np.stack([cum - num_steps, num_steps], 1).astype(np.int32))
    S = int(cum[-1]) if R else 0
    t_starts, t_ends = np.zeros(S, dtype=np.float32), np.zeros(S, dtype=np.float32)
    ridx, blidx, gidx = (np.zeros(S, dtype=np.int32) for _ in range(3))
    lib.forest_march_oracle_fill(*common, _p(packed_info), _p(t_starts), _p(t_ends), _p(ridx), _p(blidx), _p(gidx))
    return dict(packed_info=packed_info, t_starts=t_starts, t_ends=t_ends, ridx=ridx, blidx=blidx, gidx=gidx)
