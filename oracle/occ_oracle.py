"""TEST INFRASTRUCTURE ONLY -- CPU restatement (numpy) of the reference's occupancy value-grid maintenance.

Follows nr3d_lib/models/accelerations/occgrid/utils.py:18-133 (sample_pts_in_voxels arithmetic, binarize,
update_[batched_]occ_val_grid[_idx]_) and ema_single.py:214-218 (query).  The reference composes torch ops with
`torch_scatter.scatter_max` -- a third-party dependency that is NOT in this image (README.md:29, no pinned version); its
published semantics are restated here: `out` participates in the maximum (include-self), results are written back for the
touched indices only.  Pinned by tests/golden/occ_update.npz, produced by running the reference's own utils.py in this
container on CPU with torch.Tensor.scatter_reduce_(amax, include_self=True) standing in for scatter_max
(tests/golden/make_golden_occ.py).  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
import numpy as np

F = np.float32


def cells_of(pts, res):
    """((pts/2 + 0.5) * res).long().clamp(0, res-1)   (utils.py:109) -- every op rounded to fp32 separately."""
    pts = np.asarray(pts, dtype=F)
    res = np.asarray(res, dtype=np.int64)
    u = ((pts / F(2.0) + F(0.5)).astype(F) * res.astype(F)).astype(F)
    return np.clip(np.trunc(u).astype(np.int64), 0, res - 1)


def update_idx(grid, gidx, vals, ema_decay=1.0, bidx=None):
    """utils.py:93-103 / 113-126: grid[c] = max(ema * grid[c], max of vals scattered to c) for the touched cells; in place."""
    grid = np.asarray(grid)
    assert grid.dtype == F
    batched = grid.ndim == 4
    res = grid.shape[-3:]
    gidx = np.asarray(gidx, dtype=np.int64).reshape(-1, 3) if not (batched and bidx is None) else np.asarray(gidx, dtype=np.int64)
    vals = np.asarray(vals, dtype=F)
    if batched and bidx is None:      # occ_val [B, n], gidx [B, n, 3]
        B, n = vals.reshape(grid.shape[0], -1).shape
        bidx = np.repeat(np.arange(B), n)
        gidx, vals = gidx.reshape(-1, 3), vals.reshape(-1)
    flat = gidx[:, 0] * (res[1] * res[2]) + gidx[:, 1] * res[2] + gidx[:, 2]
    if batched:
        flat = np.asarray(bidx, dtype=np.int64).reshape(-1) * (res[0] * res[1] * res[2]) + flat
    g = grid.reshape(-1)
    out = (F(ema_decay) * g).astype(F)
    np.maximum.at(out, flat, vals.reshape(-1))
    g[flat] = out[flat]
    return grid


def update_pts(grid, pts, vals, ema_decay=1.0, bidx=None):
    """utils.py:105-110 / 128-133."""
    res = grid.shape[-3:]
    return update_idx(grid, cells_of(pts, res), vals, ema_decay, bidx)


def binarize(grid, thre, consider_mean=False, eps=1e-5):
    """utils.py:84-87; the mean is taken in float64 and rounded to fp32 (torch's fp32 cascade sum differs by ~1 ulp)."""
    grid = np.asarray(grid, dtype=F)
    thr = F(thre)
    if consider_mean:
        thr = min(F(F(grid.astype(np.float64).mean()) - F(eps)), thr)
    return grid > thr


def sample_pts(gidx, res, offsets, vidx=None):
    """utils.py:31,35: ((gidx[v] + offsets) / res) * 2 - 1, each op rounded to fp32."""
    gidx = np.asarray(gidx, dtype=np.int64)
    offsets = np.asarray(offsets, dtype=F)
    r = np.asarray(res, dtype=F)
    if vidx is None:
        n_per_vox = offsets.shape[1]
        vidx = np.repeat(np.arange(gidx.shape[0]), n_per_vox)
        offsets = offsets.reshape(-1, 3)
    a = (gidx[vidx].astype(F) + offsets).astype(F)
    return ((a / r).astype(F) * F(2.0) - F(1.0)).astype(F), np.asarray(vidx, dtype=np.int64)


def query(occ, pts, bidx=None):
    """ema_single.py:214-218 (+ batched)."""
    occ = np.asarray(occ)
    c = cells_of(np.asarray(pts).reshape(-1, 3), occ.shape[-3:])
    if occ.ndim == 4:
        return occ[np.asarray(bidx).reshape(-1), c[:, 0], c[:, 1], c[:, 2]]
    return occ[c[:, 0], c[:, 1], c[:, 2]]
