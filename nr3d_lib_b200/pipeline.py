"""The hot path end to end, as the reference's NeRF ray query strings it together
(nr3d_lib/graphics/nerf/nerf_ray_query.py:29-188, models/fields/nerf/renderer_mixin.py:150-313):

    occupancy march -> sample positions -> LoTD encode -> density -> alpha -> packed alpha-composite -> per-ray sums

The decoder MLP is out of scope; a softplus of the feature sum stands in for the density head
(SURVEY.md section 8d, config C3).  Everything here is a composition of the three operator families of this package,
so it doubles as the integration test target and as the M2 ("march + encode + composite rays/s") workload.
"""
from dataclasses import dataclass
from typing import Optional

import torch
import torch.nn.functional as F

from .lotd import LoTD
from .occgrid_raymarch import RaymarchRetSingle, occgrid_raymarch
from .pack_ops import packed_alpha_to_vw, packed_sum


@dataclass
class RenderOut:
    march: RaymarchRetSingle
    depth: Optional[torch.Tensor]   # [num_hit_rays] expected depth
    acc: Optional[torch.Tensor]     # [num_hit_rays] accumulated opacity
    weights: Optional[torch.Tensor]  # [num_samples]


def density_proxy(features: torch.Tensor, gain: float = 20.0) -> torch.Tensor:
    """Stand-in for the density head of the decoder: softplus of the (scaled) feature sum."""
    return F.softplus(features.sum(-1) * gain)


def march_encode_composite(encoder: LoTD, params: torch.Tensor, occ_grid: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor,
                           near: torch.Tensor, far: torch.Tensor, *, step_size: float = 0.01, max_steps: int = 512, gain: float = 20.0,
                           early_stop_eps: float = 1e-4, alpha_thre: float = 0.0) -> RenderOut:
    """One differentiable render step for rays already intersected with the [-1,1]^3 box (near / far given)."""
    ret = occgrid_raymarch(occ_grid, rays_o, rays_d, near, far, step_size=step_size, max_steps=max_steps)
    if ret.num_hit_rays == 0:
        return RenderOut(ret, None, None, None)
    x01 = ret.samples * 0.5 + 0.5                      # [-1,1] -> [0,1]  (LoTDEncoding.forward, lotd_encoding.py:162)
    h = encoder(x01, params)                           # [S, n_enc]
    sigma = density_proxy(h.float(), gain)
    alpha = 1.0 - torch.exp(-sigma * ret.deltas)       # nerf_ray_query.py:182
    w = packed_alpha_to_vw(alpha, ret.pack_infos, early_stop_eps, alpha_thre)
    depth = packed_sum(w * ret.depth_samples, ret.pack_infos)
    acc = packed_sum(w, ret.pack_infos)
    return RenderOut(ret, depth, acc, w)
