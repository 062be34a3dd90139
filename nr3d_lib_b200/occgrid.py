"""Occupancy value-grid maintenance on the B200 kernels.

Same function names, arguments and in-place semantics as the reference's
``nr3d_lib/models/accelerations/occgrid/utils.py`` (:18-133), which composes ~15 torch passes and
``torch_scatter.scatter_max``; here every update is two launches (``csrc/occ_update.cu``).  ``OccGridEma`` is a slim
counterpart of ``ema_single.py:25-218`` (state + init / step / query / sample) without the nn.Module, logging and
checkpoint plumbing, which is outside the hot path.

Random numbers (voxel picks, in-voxel offsets) are drawn with the same torch calls in the same order as the reference, so
a seeded run produces the reference's sample positions bit for bit.  No CPU fallback: CUDA tensors only.
"""
from math import prod
from typing import Optional, Tuple

import torch

from . import _lib

__all__ = ["sample_pts_in_voxels", "sample_pts_from_offsets", "binarize", "update_occ_val_grid_idx_", "update_occ_val_grid_", "update_batched_occ_val_grid_idx_",
           "update_batched_occ_val_grid_", "update_and_binarize_", "query_occ_grid", "OccGridEma"]

err_msg_empty_occ = ("Occupancy grid becomes empty during training. Your model/algorithm/training settings might be incorrect. "
                     "Please check configs and tensorboard.")

_scratch = {}


def _scratch_for(dev, n, stream):
    """uint32 scratch, zero between calls (nr3d_occ_apply resets what nr3d_occ_scatter_max touched).  One buffer per (device, stream, size):
    two grids of equal size updated on different streams never share it.  A failed update zeroes it again (see the caller)."""
    key = (dev.index, int(stream or 0), n)
    buf = _scratch.get(key)
    if buf is None:
        for k in [k for k in _scratch if k[:2] == key[:2]]:
            del _scratch[k]   # one grid shape at a time per stream is the common case; do not hoard
        buf = torch.zeros([n], dtype=torch.int32, device=dev)
        _scratch[key] = buf
    return buf


def _res_array(res3):
    import ctypes
    return (ctypes.c_int32 * 3)(*[int(r) for r in res3])


def _check_grid(fn, grid):
    if grid.dtype != torch.float32 or not grid.is_contiguous():
        raise RuntimeError(f"{fn}: `occ_val_grid` must be a contiguous float32 tensor, got {grid.dtype}, contiguous={grid.is_contiguous()}")
    return _lib.require_cuda(grid, who=fn)


def sample_pts_from_offsets(gidx: torch.Tensor, resolution, offsets: torch.Tensor, vidx: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The arithmetic of sample_pts_in_voxels (utils.py:31,35): pts = ((gidx[v] + offsets) / resolution) * 2 - 1.
    offsets [n,3] with vidx [n] (random voxel per point), or offsets [num_voxels, n_per_vox, 3] (every voxel, v = i // n_per_vox)."""
    fn = "sample_pts_in_voxels"
    if gidx.dim() != 2 or gidx.shape[-1] != 3 or offsets.dtype != torch.float32:
        raise RuntimeError(f"{fn}: the B200 build supports 3-D voxel indices and float32 points (got gidx {tuple(gidx.shape)}, {offsets.dtype})")
    dev = _lib.require_cuda(gidx, offsets, who=fn)
    gidx, offsets = gidx.contiguous().long(), offsets.contiguous()
    res = _res_array(resolution.tolist() if isinstance(resolution, torch.Tensor) else resolution)
    with torch.cuda.device(dev):
        if vidx is not None:
            n, n_per_vox, vidx_in, vidx_out = offsets.shape[0], 0, vidx.contiguous().long(), None
        else:
            n, n_per_vox, vidx_in = offsets.shape[0] * offsets.shape[1], offsets.shape[1], None
            vidx_out = torch.empty([n], device=dev, dtype=torch.long)
        pts = torch.empty([n, 3], device=dev, dtype=torch.float32)
        _lib.check(_lib.get_lib().nr3d_occ_sample_in_voxels(n, gidx.data_ptr(), _lib.ptr(vidx_in), n_per_vox, offsets.data_ptr(), res,
                                                             pts.data_ptr(), _lib.ptr(vidx_out), _lib.stream_of(dev)))
    return pts, (vidx_in if vidx is not None else vidx_out)


def sample_pts_in_voxels(gidx: torch.Tensor, num_pts: int, resolution: torch.Tensor, device=None, dtype=torch.float) -> Tuple[torch.Tensor, torch.Tensor]:
    """== sample_pts_in_voxels (utils.py:18-39): normalised points in [-1,1] inside the given voxels + voxel index per point.
    Draws the same torch random numbers in the same order as the reference."""
    assert gidx.dim() == 2, "Only support gidx with shape [N,num_dim]"
    device = device or gidx.device
    num_voxels = gidx.shape[0]
    if num_pts / num_voxels < 2.0:
        vidx = torch.randint(num_voxels, [num_pts, ], device=device)
        offsets = torch.rand([num_pts, gidx.shape[-1]], device=device, dtype=dtype)
        return sample_pts_from_offsets(gidx, resolution, offsets, vidx)
    n_per_vox = int(num_pts // num_voxels) + 1
    offsets = torch.rand([num_voxels, n_per_vox, gidx.shape[-1]], device=device, dtype=dtype)
    return sample_pts_from_offsets(gidx, resolution, offsets)


def binarize(occ_val: torch.Tensor, occ_threshold: float, consider_mean=False, eps=1e-5) -> torch.Tensor:
    """== binarize (utils.py:84-87) -> bool tensor of the grid's shape."""
    dev = _check_grid("binarize", occ_val)
    with torch.cuda.device(dev):
        occ = torch.empty(occ_val.shape, dtype=torch.bool, device=dev)
        total = occ_val.sum(dtype=torch.float64).reshape(1) if consider_mean else None
        _lib.check(_lib.get_lib().nr3d_occ_binarize(occ_val.numel(), occ_val.data_ptr(), float(occ_threshold), int(bool(consider_mean)),
                                                    float(eps), _lib.ptr(total), occ.data_ptr(), _lib.stream_of(dev)))
    return occ


def _update(fn, grid, *, pts=None, gidx=None, bidx=None, batch_data_size=0, occ_val, ema_decay, B, res3,
            occ_out=None, occ_thre=0.0, consider_mean=False, eps=1e-5, extra=None):
    """scatter (+ optional second scatter of `extra` = (gidx [M,3], bidx-or-None, vals [M]) into the same scratch) -> apply."""
    dev = _check_grid(fn, grid)
    occ_val = occ_val.flatten().to(grid).contiguous()
    N = occ_val.shape[0]
    if pts is not None:
        pts = pts.reshape(-1, 3).to(torch.float32).contiguous()
        if pts.shape[0] != N:
            raise RuntimeError(f"{fn}: {pts.shape[0]} points but {N} values")
    else:
        gidx = gidx.reshape(-1, 3).long().contiguous()
        if gidx.shape[0] != N:
            raise RuntimeError(f"{fn}: {gidx.shape[0]} voxel indices but {N} values")
    if bidx is not None:
        bidx = bidx.flatten().long().contiguous()
        if bidx.shape[0] != N:
            raise RuntimeError(f"{fn}: {bidx.shape[0]} batch indices but {N} values")
    _lib.require_cuda(*(t for t in (pts, gidx, bidx, occ_val) if t is not None), who=fn)
    lib = _lib.get_lib()
    n_cells = grid.numel()
    with torch.cuda.device(dev):
        st = _lib.stream_of(dev)
        scratch = _scratch_for(dev, n_cells, st)
        try:
            _lib.check(lib.nr3d_occ_scatter_max(N, _lib.ptr(pts), _lib.ptr(gidx), _lib.ptr(bidx), int(batch_data_size), occ_val.data_ptr(), int(B),
                                                _res_array(res3), scratch.data_ptr(), st))
            if extra is not None:
                e_gidx, e_bidx, e_val = extra
                e_gidx, e_val = e_gidx.reshape(-1, 3).long().contiguous(), e_val.flatten().to(grid).contiguous()
                e_bidx = None if e_bidx is None else e_bidx.flatten().long().contiguous()
                _lib.check(lib.nr3d_occ_scatter_max(e_val.shape[0], None, e_gidx.data_ptr(), _lib.ptr(e_bidx), 0, e_val.data_ptr(), int(B),
                                                    _res_array(res3), scratch.data_ptr(), st))
            fused = occ_out is not None and not consider_mean
            total = torch.zeros([1], dtype=torch.float64, device=dev) if (occ_out is not None and consider_mean) else None
            _lib.check(lib.nr3d_occ_apply(n_cells, grid.data_ptr(), scratch.data_ptr(), float(ema_decay), int(fused), float(occ_thre),
                                          _lib.ptr(occ_out) if fused else None, _lib.ptr(total), st))
        except BaseException:
            scratch.zero_()   # nr3d_occ_apply did not run (or failed): do not leave stale maxima for the next update
            raise
        if total is not None:
            _lib.check(lib.nr3d_occ_binarize(n_cells, grid.data_ptr(), float(occ_thre), 1, float(eps), total.data_ptr(), occ_out.data_ptr(), st))


def update_occ_val_grid_idx_(occ_val_grid: torch.Tensor, gidx: torch.Tensor, occ_val: torch.Tensor, ema_decay: float = 1.0):
    """== update_occ_val_grid_idx_ (utils.py:93-103), in place."""
    _update("update_occ_val_grid_idx_", occ_val_grid, gidx=gidx, occ_val=occ_val, ema_decay=ema_decay, B=1, res3=occ_val_grid.shape)


def update_occ_val_grid_(occ_val_grid: torch.Tensor, pts: torch.Tensor, occ_val: torch.Tensor, ema_decay: float = 1.0):
    """== update_occ_val_grid_ (utils.py:105-110), in place; pts in [-1,1]."""
    _update("update_occ_val_grid_", occ_val_grid, pts=pts, occ_val=occ_val, ema_decay=ema_decay, B=1, res3=occ_val_grid.shape)


def update_batched_occ_val_grid_idx_(occ_val_grid: torch.Tensor, bidx: Optional[torch.Tensor] = None, gidx: torch.Tensor = ..., occ_val: torch.Tensor = ...,
                                     ema_decay: float = 1.0):
    """== update_batched_occ_val_grid_idx_ (utils.py:113-126): bidx given -> flat points; bidx None -> occ_val [B, num_pts], gidx [B, num_pts, 3]."""
    B = occ_val_grid.shape[0]
    bds = 0 if bidx is not None else prod(gidx.shape[1:-1])
    _update("update_batched_occ_val_grid_idx_", occ_val_grid, gidx=gidx, bidx=bidx, batch_data_size=bds, occ_val=occ_val, ema_decay=ema_decay, B=B,
            res3=occ_val_grid.shape[1:])


def update_batched_occ_val_grid_(occ_val_grid: torch.Tensor, pts: torch.Tensor, bidx: Optional[torch.Tensor] = None, occ_val: torch.Tensor = ...,
                                 ema_decay: float = 1.0):
    """== update_batched_occ_val_grid_ (utils.py:128-133)."""
    B = occ_val_grid.shape[0]
    bds = 0 if bidx is not None else prod(pts.shape[1:-1])
    _update("update_batched_occ_val_grid_", occ_val_grid, pts=pts, bidx=bidx, batch_data_size=bds, occ_val=occ_val, ema_decay=ema_decay, B=B,
            res3=occ_val_grid.shape[1:])


def update_and_binarize_(occ_val_grid: torch.Tensor, occ_grid: torch.Tensor, *, pts=None, gidx=None, bidx=None, occ_val: torch.Tensor, ema_decay: float,
                         occ_threshold: float, consider_mean=False, eps=1e-5, extra=None):
    """The reference's `_step_update_occ` tail (ema_single.py:201-202, ema_batched.py:259-260) in one call: EMA-max update of
    `occ_val_grid` and refresh of the bool `occ_grid`, both in place (2 launches, 3 with the mean-relative threshold).
    `extra` = (gidx, bidx-or-None, vals): additional (voxel, value) pairs merged into the same update (collected samples)."""
    if occ_grid.dtype != torch.bool or occ_grid.shape != occ_val_grid.shape or not occ_grid.is_contiguous():
        raise RuntimeError("update_and_binarize_: `occ_grid` must be a contiguous bool tensor of the value grid's shape")
    batched = occ_val_grid.dim() == 4
    B = occ_val_grid.shape[0] if batched else 1
    src = pts if pts is not None else gidx
    bds = prod(src.shape[1:-1]) if (batched and bidx is None) else 0
    _update("update_and_binarize_", occ_val_grid, pts=pts, gidx=gidx, bidx=bidx, batch_data_size=bds, occ_val=occ_val, ema_decay=ema_decay, B=B,
            res3=occ_val_grid.shape[-3:], occ_out=occ_grid, occ_thre=occ_threshold, consider_mean=consider_mean, eps=eps, extra=extra)


def query_occ_grid(occ_grid: torch.Tensor, pts: torch.Tensor, bidx: Optional[torch.Tensor] = None) -> torch.Tensor:
    """== OccGridEma.query / OccGridEmaBatched.query (ema_single.py:214-218): bool per point, pts in [-1,1]."""
    fn = "query_occ_grid"
    if occ_grid.dtype != torch.bool or not occ_grid.is_contiguous():
        raise RuntimeError(f"{fn}: `occ_grid` must be a contiguous bool tensor")
    dev = _lib.require_cuda(occ_grid, pts, who=fn)
    batched = occ_grid.dim() == 4
    B = occ_grid.shape[0] if batched else 1
    prefix = pts.shape[:-1]
    flat = pts.reshape(-1, 3).to(torch.float32).contiguous()
    bds = 0
    if batched and bidx is None:
        bds = prod(pts.shape[1:-1])
    if bidx is not None:
        bidx = bidx.flatten().long().contiguous()
    with torch.cuda.device(dev):
        out = torch.empty([flat.shape[0]], dtype=torch.bool, device=dev)
        _lib.check(_lib.get_lib().nr3d_occ_query(flat.shape[0], flat.data_ptr(), _lib.ptr(bidx), int(bds), int(B), _res_array(occ_grid.shape[-3:]),
                                                 occ_grid.data_ptr(), out.data_ptr(), _lib.stream_of(dev)))
    return out.view(prefix)


class OccGridEma:
    """Single-scene EMA occupancy grid (behaviour: ema_single.py:25-218).  `occ_val_fn` maps network values to occupancy."""

    def __init__(self, resolution, occ_val_fn=None, occ_thre: float = 0.3, ema_decay: float = 0.95, occ_thre_consider_mean: bool = False,
                 n_steps_between_update: int = 16, n_steps_warmup: int = 256, should_collect_samples: bool = False, device=None):
        device = torch.device(device if device is not None else "cuda")
        if device.type != "cuda":
            raise RuntimeError("OccGridEma: a CUDA device is required (no CPU path in the B200 build)")
        self.resolution = torch.tensor([int(r) for r in resolution], dtype=torch.long, device=device)
        self.occ_val_fn = occ_val_fn if occ_val_fn is not None else (lambda v: v)
        self.occ_thre, self.ema_decay, self.occ_thre_consider_mean = float(occ_thre), float(ema_decay), bool(occ_thre_consider_mean)
        self.n_steps_between_update, self.n_steps_warmup = int(n_steps_between_update), int(n_steps_warmup)
        self.should_collect_samples = bool(should_collect_samples)
        res_l = self.resolution.tolist()
        self.occ_val_grid = torch.zeros(res_l, dtype=torch.float32, device=device)
        self.occ_grid = torch.ones(res_l, dtype=torch.bool, device=device)
        self._occ_val_grid_pcl = torch.zeros(res_l, dtype=torch.float32, device=device) if should_collect_samples else None
        self.gidx_full = torch.stack(torch.meshgrid([torch.arange(r, device=device) for r in res_l], indexing="ij"), -1).view(-1, 3)
        self.is_initialized = False

    @torch.no_grad()
    def init_from_constant(self, constant_value: float):
        self.occ_val_grid.fill_(constant_value)
        self.occ_grid = binarize(self.occ_val_grid, self.occ_thre, self.occ_thre_consider_mean)
        self.is_initialized = True

    @torch.no_grad()
    def init_from_net(self, val_query_fn, num_steps=4, num_pts: int = 2 ** 18):
        for _ in range(num_steps):
            gidx_empty = self.occ_grid.logical_not().nonzero().long()
            if gidx_empty.shape[0] > 0:
                pts = sample_pts_in_voxels(gidx_empty, num_pts, self.resolution)[0]
                self._update(pts=pts, occ_val=self.occ_val_fn(val_query_fn(pts)), ema_decay=1.0)
        self.is_initialized = True

    def _update(self, *, pts=None, gidx=None, occ_val, ema_decay, extra=None):
        update_and_binarize_(self.occ_val_grid, self.occ_grid, pts=pts, gidx=gidx, occ_val=occ_val, ema_decay=ema_decay,
                             occ_threshold=self.occ_thre, consider_mean=self.occ_thre_consider_mean, extra=extra)

    @torch.no_grad()
    def step(self, cur_it: int, val_query_fn, num_steps=4, num_pts: int = 2 ** 18) -> bool:
        if not (cur_it > 0 and cur_it % self.n_steps_between_update == 0):
            return False
        pts_list = []
        if cur_it < self.n_steps_warmup:
            for _ in range(num_steps):
                pts_list.append(sample_pts_in_voxels(self.gidx_full, num_pts, self.resolution)[0])
        else:
            gidx_nonempty = self.occ_grid.nonzero().long()
            gidx_empty = self.occ_grid.logical_not().nonzero().long()
            assert gidx_nonempty.numel() > 0, err_msg_empty_occ
            for _ in range(num_steps):
                pts_list.append(sample_pts_in_voxels(self.gidx_full, num_pts // 2, self.resolution)[0])
                if gidx_empty.numel() > 0:
                    pts_list.append(sample_pts_in_voxels(gidx_empty, num_pts // 4, self.resolution)[0])
                pts_list.append(sample_pts_in_voxels(gidx_nonempty, num_pts // 4, self.resolution)[0])
        pts = torch.cat(pts_list, 0)
        self.step_update_occ(pts, val_query_fn(pts))
        return True

    @torch.no_grad()
    def step_update_occ(self, pts: torch.Tensor, val: torch.Tensor):
        """== _step_update_occ (ema_single.py:186-202)."""
        occ_val = self.occ_val_fn(val.flatten())
        extra = None
        if self.should_collect_samples:
            idx_pcl = self._occ_val_grid_pcl.nonzero().long()
            if idx_pcl.numel() > 0:     # the collected samples join the update as extra (voxel, value) pairs
                extra = (idx_pcl, None, self._occ_val_grid_pcl[tuple(idx_pcl.t())])
            self._occ_val_grid_pcl.zero_()
        self._update(pts=pts, occ_val=occ_val, ema_decay=self.ema_decay, extra=extra)

    @torch.no_grad()
    def collect_samples(self, pts: torch.Tensor, val: torch.Tensor):
        if self.should_collect_samples:
            update_occ_val_grid_(self._occ_val_grid_pcl, pts, self.occ_val_fn(val), ema_decay=1.0)

    @torch.no_grad()
    def sample_pts_in_occupied(self, num_pts: int) -> torch.Tensor:
        gidx_nonempty = self.occ_grid.nonzero().long()
        assert gidx_nonempty.numel() > 0, err_msg_empty_occ
        return sample_pts_in_voxels(gidx_nonempty, num_pts, self.resolution)[0]

    @torch.no_grad()
    def query(self, pts: torch.Tensor) -> torch.Tensor:
        return query_occ_grid(self.occ_grid, pts)
