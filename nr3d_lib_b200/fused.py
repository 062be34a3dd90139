"""Fused LoTD encoder + density decoder, forward only.

Replaces the composition in ``LoTDNeRF.query_density`` / ``forward_density``
(nr3d_lib/models/fields/nerf/lotd_nerf.py:136-178): ``encoding(x) -> density_decoder(h) -> activation(out[..., 0])`` with
a one-hidden-layer decoder (``models/blocks/mlp.py``: Linear(32, 64) -> ReLU -> Linear(64, 1 + n_extra), n_extra <= 15).
One kernel (``csrc/lotd_fused.cu``): the features never reach HBM, both layers run on the tensor cores (tcgen05, bf16
operands, fp32 accumulation).  No autograd: this is the no-grad query used by occupancy-grid updates and rendering.
"""
import ctypes
from typing import Optional, Tuple

import torch

from . import _lib
from .bindings import _lotd

_ACT = {"identity": 0, "none": 0, "exp": 1, "softplus": 2, "relu": 3}


def pack_kmajor_bf16(w: torch.Tensor, rows: int) -> torch.Tensor:
    """[rows_in, K] weight matrix -> bf16 bytes in the K-major core-matrix order tcgen05.mma reads from shared memory:
    [K/8][rows/8][8 rows][8 k] (core matrix = 8 rows x 16 bytes); rows_in < rows is zero padded."""
    n, k = w.shape
    if k % 8 or rows % 8 or n > rows:
        raise RuntimeError(f"pack_kmajor_bf16: shape {tuple(w.shape)} does not fit [{rows}, K % 8 == 0]")
    full = torch.zeros([rows, k], dtype=torch.float32, device=w.device)
    full[:n] = w.float()
    return full.to(torch.bfloat16).view(rows // 8, 8, k // 8, 8).permute(2, 0, 1, 3).contiguous()


class FusedDensityDecoder:
    """density = activation( (relu(h W1^T + b1) W2^T + b2)[..., 0] ),  h = LoTD(x).  Weights are torch Linear-style [out, in]."""

    def __init__(self, meta, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor, b2: Optional[torch.Tensor], activation: str = "exp"):
        if activation not in _ACT:
            raise RuntimeError(f"FusedDensityDecoder: activation must be one of {sorted(_ACT)}, got {activation!r}")
        if not (meta.c_hash_only and meta.n_dims_to_encode == 3 and meta.n_feat_per_pseudo_lvl == 2 and meta.n_encoded_dims == 32):
            raise RuntimeError("FusedDensityDecoder: needs a Dense/Hash-only LoDMeta with D=3, F=2 and 32 encoded dims")
        if tuple(w1.shape) != (64, 32) or w2.dim() != 2 or w2.shape[1] != 64 or not (1 <= w2.shape[0] <= 16):
            raise RuntimeError(f"FusedDensityDecoder: expected w1 [64, 32] and w2 [<=16, 64], got {tuple(w1.shape)} and {tuple(w2.shape)}")
        self.dev = _lib.require_cuda(w1, w2, b1, b2, who="FusedDensityDecoder")
        self.meta, self.activation, self.n_out = meta, _ACT[activation], w2.shape[0]
        self.w1p, self.w2p = pack_kmajor_bf16(w1, 64), pack_kmajor_bf16(w2, 16)
        self.b1 = None if b1 is None else b1.detach().float().contiguous()
        b2p = torch.zeros([16], dtype=torch.float32, device=self.dev)
        if b2 is not None:
            b2p[: self.n_out] = b2.detach().float()
        self.b2 = b2p

    @torch.no_grad()
    def query_density(self, x: torch.Tensor, params: torch.Tensor, max_level: Optional[int] = None, return_output: bool = False
                      ) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
        """x [N, 3] fp32 in [0, 1], params fp32 [n_params] -> (sigma [N], decoder output [N, n_out] or None)."""
        fn = "FusedDensityDecoder.query_density"
        if x.dim() != 2 or x.shape[1] != 3 or x.dtype != torch.float32 or not x.is_contiguous():
            raise RuntimeError(f"{fn}: expected a contiguous fp32 [N, 3] tensor for `x`")
        if params.dtype != torch.float32 or params.shape[0] != self.meta.n_params or not params.is_contiguous():
            raise RuntimeError(f"{fn}: expected contiguous fp32 params of size n_params={self.meta.n_params}")
        dev = _lib.require_cuda(x, params, self.w1p, self.b2, self.b1, who=fn)    # ... and on the decoder's device
        N = x.shape[0]
        ml = self.meta.n_levels if max_level is None else int(max_level)
        with torch.cuda.device(dev):
            sigma = torch.empty([N], dtype=torch.float32, device=dev)
            out16 = torch.empty([N, 16], dtype=torch.float32, device=dev) if return_output else None
            if N:
                # same clamp as the unfused path (reference lotd.py:211 / nr3d_lib_b200/lotd.py): points on or outside the box boundary index
                # the last cell instead of reading Dense levels out of bounds
                xs, _ = _lotd._sorted_points(x.clamp(1.0e-6, 1.0 - 1.0e-6))
                _lib.check(_lib.get_lib().nr3d_lotd_fused_density_fwd(
                    ctypes.byref(self.meta._c), N, xs.data_ptr(), params.data_ptr(), ml, self.w1p.data_ptr(), _lib.ptr(self.b1),
                    self.w2p.data_ptr(), self.b2.data_ptr(), self.activation, sigma.data_ptr(), _lib.ptr(out16), _lib.stream_of(dev)))
        return sigma, (None if out16 is None else out16[:, : self.n_out])
