"""Host-side mirror of the reference's occupancy-grid ray-march wrappers
(nr3d_lib/graphics/raymarch/occgrid_raymarch.py:25-221 and the ``RaymarchRet*`` dataclasses of raymarch/__init__.py).
"""
from dataclasses import dataclass, fields
from enum import Enum
from typing import Literal, Optional, Union

import torch

from .bindings import _occ_grid as _backend
from .pack_ops import get_pack_infos_from_boundary, mark_pack_boundaries, packed_diff

__all__ = ['ContractionType', 'RaymarchRetSingle', 'RaymarchRetBatched', 'RaymarchRetForest', 'occgrid_raymarch', 'occgrid_raymarch_batched',
           'occgrid_raymarch_forest']


class ContractionType(Enum):
    AABB = int(_backend.ContractionType.AABB)
    UN_BOUNDED_TANH = int(_backend.ContractionType.UN_BOUNDED_TANH)
    UN_BOUNDED_SPHERE = int(_backend.ContractionType.UN_BOUNDED_SPHERE)


@dataclass
class _RetBase:
    num_hit_rays: int
    ridx_hit: Optional[torch.Tensor]        # [num_hit_rays]      indices of the rays that produced samples
    samples: Optional[torch.Tensor]         # [num_samples, 3]    sample positions
    depth_samples: Optional[torch.Tensor]   # [num_samples]       sample depths
    deltas: Optional[torch.Tensor]          # [num_samples]       interval lengths
    ridx: Optional[torch.Tensor]            # [num_samples]       ray index of each sample
    pack_infos: Optional[torch.Tensor]      # [num_hit_rays, 2]   (first sample, n samples) per hit ray

    def __iter__(self):
        return iter(tuple(getattr(self, f.name) for f in fields(self)))

    def __getitem__(self, name: str):
        return getattr(self, name)


@dataclass
class RaymarchRetSingle(_RetBase):
    gidx: Optional[torch.Tensor]            # [num_samples] voxel index of each sample
    gidx_pack_infos: Optional[torch.Tensor]


@dataclass
class RaymarchRetBatched(_RetBase):
    bidx: Optional[torch.Tensor]            # [num_samples] batch index of each sample
    gidx: Optional[torch.Tensor]
    gidx_pack_infos: Optional[torch.Tensor]


@dataclass
class RaymarchRetForest(_RetBase):
    blidx: Optional[torch.Tensor]            # [num_samples] block index of each sample
    blidx_pack_infos: Optional[torch.Tensor]  # [num_block_packs, 2] one pack per run of samples in the same block
    gidx: Optional[torch.Tensor]
    gidx_pack_infos: Optional[torch.Tensor]


_CONTRACTION = {'aabb': ContractionType.AABB, 'sphere': ContractionType.UN_BOUNDED_SPHERE, 'tanh': ContractionType.UN_BOUNDED_TANH}


def _contraction(name: str):
    try:
        return _backend.ContractionType(_CONTRACTION[name.lower()].value)
    except KeyError:
        raise RuntimeError(f"Invalid constraction={name}")


def _finish(rays_o, rays_d, pack_infos, t_starts, t_ends, ridx, perturb, perturb_before_march):
    ridx_hit = pack_infos[..., 1].nonzero().long()[..., 0].contiguous()
    if ridx_hit.numel() == 0:
        return None
    pack_infos = pack_infos[ridx_hit].contiguous().long()
    t_ends.squeeze_(-1), t_starts.squeeze_(-1)
    # one pass instead of index_select x2 + addcmul + sub (reference occgrid_raymarch.py:96-107); same values
    samples, deltas = _backend.march_samples(rays_o.contiguous(), rays_d.contiguous(), t_starts, t_ends, ridx)
    t_samples = t_starts
    if perturb and not perturb_before_march:
        t_samples = torch.addcmul(t_starts, torch.rand_like(deltas), deltas)
        deltas = packed_diff(t_samples, pack_infos)  # last delta of each pack defaults to zero
    return ridx_hit, pack_infos, deltas, t_samples, samples


def occgrid_raymarch(occ_grid: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, near: Union[torch.Tensor, float],
                     far: Union[torch.Tensor, float], *, constraction: Literal['aabb', 'tanh', 'sphere'] = 'aabb', perturb=False,
                     perturb_before_march=False, roi: torch.Tensor = None, step_size: float = 1e-3, max_step_size: float = 1e10,
                     dt_gamma: float = 0.0, max_steps: int = 512, step_size_factor=1.0) -> RaymarchRetSingle:
    """March rays through a binary occupancy grid [rx, ry, rz] (reference occgrid_raymarch.py:25-112)."""
    step_size *= step_size_factor
    dt_gamma *= step_size_factor
    device, dtype = rays_o.device, rays_o.dtype
    near = rays_o.new_full(rays_o.shape[:-1], near) if not isinstance(near, torch.Tensor) else near
    far = rays_o.new_full(rays_o.shape[:-1], far) if not isinstance(far, torch.Tensor) else far
    if roi is None:
        roi = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=dtype, device=device)
    ctype = _contraction(constraction)
    if perturb and perturb_before_march:
        near = near + step_size * torch.rand_like(near)
    pack_infos, t_starts, t_ends, ridx, gidx = _backend.ray_marching(rays_o, rays_d, near, far, roi, occ_grid, ctype, step_size,
                                                                     max_step_size, dt_gamma, max_steps, True)
    ridx, gidx = ridx.long(), gidx.long()
    fin = _finish(rays_o, rays_d, pack_infos, t_starts, t_ends, ridx, perturb, perturb_before_march)
    if fin is None:
        return RaymarchRetSingle(0, None, None, None, None, None, None, None, None)
    ridx_hit, pack_infos, deltas, _t_samples, samples = fin
    # NOTE: the reference returns `t_starts` (not the perturbed depths) as depth_samples in the single-block variant
    return RaymarchRetSingle(ridx_hit.numel(), ridx_hit, samples, t_starts, deltas, ridx, pack_infos, gidx, None)


def occgrid_raymarch_batched(occ_grid: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, rays_bidx: torch.Tensor = None,
                             near: Union[torch.Tensor, float] = ..., far: Union[torch.Tensor, float] = ..., *,
                             constraction: Literal['aabb', 'tanh', 'sphere'] = 'aabb', perturb=False, perturb_before_march=False,
                             roi: torch.Tensor = None, step_size: float = 1e-3, max_step_size: float = 1e10, dt_gamma: float = 0.0,
                             max_steps: int = 512, step_size_factor=1.0) -> RaymarchRetBatched:
    """March rays through batched occupancy grids [B, rx, ry, rz] (reference occgrid_raymarch.py:114-221)."""
    step_size *= step_size_factor
    dt_gamma *= step_size_factor
    assert occ_grid.dim() == 4, "Requires batched occ grid input of shape [B,Nx,Ny,Nz]"
    B = occ_grid.shape[0]
    device, dtype = rays_o.device, rays_o.dtype
    near = rays_o.new_full(rays_o.shape[:-1], near) if not isinstance(near, torch.Tensor) else near
    far = rays_o.new_full(rays_o.shape[:-1], far) if not isinstance(far, torch.Tensor) else far
    if rays_bidx is None:
        assert rays_o.dim() == 3 and rays_o.shape[0] == B, "When not given rays_bidx, inputs should be batched"
        batch_data_size = rays_o.shape[1]
        rays_o, rays_d = rays_o.flatten(0, -2), rays_d.flatten(0, -2)
        near, far = near.flatten(), far.flatten()
    else:
        assert rays_o.dim() == 2 and [*rays_o.shape[:-1]] == [*rays_bidx.shape], \
            "When given rays_bidx, inputs should have the same size with rays_bidx"
        rays_bidx = rays_bidx.int().contiguous()
        batch_data_size = 0
    if roi is None:
        roi = torch.tensor([-1, -1, -1, 1, 1, 1], dtype=dtype, device=device).tile(B, 1)
    elif roi.dim() == 1:
        roi = roi.tile(B, 1)
    else:
        assert roi.dim() == 2 and roi.shape[0] == B
    ctype = _contraction(constraction)
    if perturb and perturb_before_march:
        near = near + step_size * torch.rand_like(near)
    pack_infos, t_starts, t_ends, ridx, bidx, gidx = _backend.batched_ray_marching(
        rays_o.contiguous(), rays_d.contiguous(), near.contiguous(), far.contiguous(), rays_bidx, batch_data_size, roi.contiguous(),
        occ_grid, ctype, step_size, max_step_size, dt_gamma, max_steps, True)
    ridx, bidx, gidx = ridx.long(), bidx.long(), gidx.long()
    fin = _finish(rays_o, rays_d, pack_infos, t_starts, t_ends, ridx, perturb, perturb_before_march)
    if fin is None:
        return RaymarchRetBatched(0, None, None, None, None, None, None, None, None, None)
    ridx_hit, pack_infos, deltas, t_samples, samples = fin
    return RaymarchRetBatched(ridx_hit.numel(), ridx_hit, samples, t_samples, deltas, ridx, pack_infos, bidx, gidx, None)


def occgrid_raymarch_forest(forest_meta, occ_grid: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor, near: Union[torch.Tensor, float],
                            far: Union[torch.Tensor, float], seg_block_inds: torch.Tensor, seg_entries: torch.Tensor, seg_exits: torch.Tensor,
                            seg_pack_infos: torch.Tensor, *, perturb=False, perturb_before_march=False, step_size: float = 1e-3,
                            max_step_size: float = 1e10, dt_gamma: float = 0.0, max_steps: int = 512, step_size_factor=1.0) -> RaymarchRetForest:
    """March rays through a forest of occupancy grids [n_blocks, rx, ry, rz] along their block segments
    (reference occgrid_raymarch.py:223-272)."""
    step_size *= step_size_factor
    dt_gamma *= step_size_factor
    near = rays_o.new_full(rays_o.shape[:-1], near) if not isinstance(near, torch.Tensor) else near
    far = rays_o.new_full(rays_o.shape[:-1], far) if not isinstance(far, torch.Tensor) else far
    if perturb and perturb_before_march:
        near = near + step_size * torch.rand_like(near)
    pack_infos, t_starts, t_ends, ridx, blidx, _ = _backend.forest_ray_marching(
        forest_meta, rays_o, rays_d, near, far, seg_block_inds.int().contiguous(), seg_entries.contiguous(), seg_exits.contiguous(),
        seg_pack_infos.int().contiguous(), occ_grid, step_size, max_step_size, dt_gamma, max_steps, False)
    ridx, blidx = ridx.long(), blidx.long()
    fin = _finish(rays_o, rays_d, pack_infos, t_starts, t_ends, ridx, perturb, perturb_before_march)
    if fin is None:
        return RaymarchRetForest(0, None, None, None, None, None, None, None, None, None, None)
    ridx_hit, pack_infos, deltas, t_samples, samples = fin
    blidx_pack_infos = get_pack_infos_from_boundary(mark_pack_boundaries(blidx))
    return RaymarchRetForest(ridx_hit.numel(), ridx_hit, samples, t_samples, deltas, ridx, pack_infos, blidx, blidx_pack_infos, None, None)
