"""Fused LoTD encoder + density decoder: no-grad query (`FusedDensityDecoder`) and the trainable module
(`FusedDensityMLP` / `fused_density`, forward + backward kernels).

Replaces the composition in ``LoTDNeRF.query_density`` / ``forward_density``
(nr3d_lib/models/fields/nerf/lotd_nerf.py:136-178): ``encoding(x) -> density_decoder(h) -> activation(out[..., 0])`` with
a one-hidden-layer decoder (``models/blocks/mlp.py``: Linear(32, 64) -> ReLU -> Linear(64, 1 + n_extra), n_extra <= 15).
One kernel each way (``csrc/lotd_fused.cu``, ``csrc/lotd_fused_bwd.cu``): the features, the hidden activations and their gradients never
reach HBM, all GEMMs run on the tensor cores (tcgen05, bf16 operands, fp32 accumulation in TMEM), the backward ends in the run-merged
LoTD scatter.  ``FusedDensityDecoder`` is the no-grad query used by occupancy-grid updates and rendering; ``FusedDensityMLP`` trains.
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
                xs, _ = _lotd._sorted_points(x.clamp(1.0e-6, 1.0 - 1.0e-6), expect_new=True)
                _lib.check(_lib.get_lib().nr3d_lotd_fused_density_fwd(
                    ctypes.byref(self.meta._c), N, xs.data_ptr(), params.data_ptr(), ml, self.w1p.data_ptr(), _lib.ptr(self.b1),
                    self.w2p.data_ptr(), self.b2.data_ptr(), self.activation, sigma.data_ptr(), _lib.ptr(out16), _lib.stream_of(dev)))
        return sigma, (None if out16 is None else out16[:, : self.n_out])


class _FusedDensityFn(torch.autograd.Function):
    """sigma (, out16) = decoder(LoTD(x)) with gradients to the LoTD tables and the decoder weights (not to x)."""

    @staticmethod
    def forward(ctx, x, params, w1, b1, w2, b2, meta, activation, want_out):
        fn = "fused_density"
        if x.dim() != 2 or x.shape[1] != 3 or x.dtype != torch.float32:
            raise RuntimeError(f"{fn}: expected an fp32 [N, 3] tensor for `x`")
        if params.dtype != torch.float32 or params.dim() != 1 or params.shape[0] != meta.n_params or not params.is_contiguous():
            raise RuntimeError(f"{fn}: expected contiguous fp32 params of size n_params={meta.n_params}")
        if not (meta.c_hash_only and meta.n_dims_to_encode == 3 and meta.n_feat_per_pseudo_lvl == 2 and meta.n_encoded_dims == 32):
            raise RuntimeError(f"{fn}: needs a Dense/Hash-only LoDMeta with D=3, F=2 and 32 encoded dims")
        if tuple(w1.shape) != (64, 32) or w2.dim() != 2 or w2.shape[1] != 64 or not (1 <= w2.shape[0] <= 16):
            raise RuntimeError(f"{fn}: expected w1 [64, 32] and w2 [<=16, 64], got {tuple(w1.shape)} and {tuple(w2.shape)}")
        dev = _lib.require_cuda(x, params, w1, b1, w2, b2, who=fn)
        N, n_out = x.shape[0], w2.shape[0]
        xc = x.detach().clamp(1.0e-6, 1.0 - 1.0e-6).contiguous()
        with torch.cuda.device(dev):
            w1p, w2p = pack_kmajor_bf16(w1.detach(), 64), pack_kmajor_bf16(w2.detach(), 16)
            b1f = None if b1 is None else b1.detach().float().contiguous()
            b2p = torch.zeros([16], dtype=torch.float32, device=dev)
            if b2 is not None:
                b2p[:n_out] = b2.detach().float()
            sigma = torch.empty([N], dtype=torch.float32, device=dev)
            out16 = torch.empty([N, 16], dtype=torch.float32, device=dev) if want_out else None
            if N:
                xs, _ = _lotd._sorted_points(xc, expect_new=True)
                _lib.check(_lib.get_lib().nr3d_lotd_fused_density_fwd(
                    ctypes.byref(meta._c), N, xs.data_ptr(), params.data_ptr(), meta.n_levels, w1p.data_ptr(), _lib.ptr(b1f), w2p.data_ptr(),
                    b2p.data_ptr(), activation, sigma.data_ptr(), _lib.ptr(out16), _lib.stream_of(dev)))
        ctx.save_for_backward(xc, params, w1, b1, w2, sigma)
        ctx.meta, ctx.activation, ctx.n_out, ctx.has_b2 = meta, activation, n_out, b2 is not None
        out = None if out16 is None else out16[:, :n_out]
        return sigma, out

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_sigma, d_out):
        xc, params, w1, b1, w2, sigma = ctx.saved_tensors
        meta, n_out = ctx.meta, ctx.n_out
        dev, N = xc.device, xc.shape[0]
        if d_sigma is None and d_out is None:
            return (None,) * 9
        with torch.cuda.device(dev):
            g_p = torch.zeros_like(params)
            g = torch.zeros([64 * 32 + 64 + 16 * 64 + 16], dtype=torch.float32, device=dev)      # dW1 | db1 | dW2 (16 rows) | db2 (16)
            g_w1, g_b1, g_w2, g_b2 = g[:2048].view(64, 32), g[2048:2112], g[2112:3136].view(16, 64), g[3136:3152]
            if N:
                d16 = None
                if d_out is not None:
                    d16 = torch.zeros([N, 16], dtype=torch.float32, device=dev)
                    d16[:, :n_out] = d_out
                ds = None if d_sigma is None else d_sigma.float().contiguous()
                w2_16 = torch.zeros([16, 64], dtype=torch.float32, device=dev)
                w2_16[:n_out] = w2.detach().float()
                w1p = pack_kmajor_bf16(w1.detach(), 64)                     # [64 n][32 k]   layer 1
                w2tp = pack_kmajor_bf16(w2_16.t().contiguous(), 64)         # [64 n][16 k]   dH = G W2
                w1tp = pack_kmajor_bf16(w1.detach().t().contiguous(), 32)   # [32 n][64 k]   dF = DH W1
                b1f = None if b1 is None else b1.detach().float().contiguous()
                xs, _ = _lotd._sorted_points(xc)       # the forward's records unless other points went through this stream since
                _lib.check(_lib.get_lib().nr3d_lotd_fused_density_bwd(
                    ctypes.byref(meta._c), N, xs.data_ptr(), params.data_ptr(), meta.n_levels, w1p.data_ptr(), w2tp.data_ptr(), w1tp.data_ptr(),
                    _lib.ptr(b1f), ctx.activation, sigma.data_ptr(), _lib.ptr(ds), _lib.ptr(d16), g_p.data_ptr(), g_w1.data_ptr(), g_b1.data_ptr(),
                    g_w2.data_ptr(), g_b2.data_ptr(), _lib.stream_of(dev)))
        return (None, g_p, g_w1.to(w1.dtype), None if b1 is None else g_b1.to(b1.dtype), g_w2[:n_out].to(w2.dtype),
                g_b2[:n_out].clone() if ctx.has_b2 else None, None, None, None)


def fused_density(x: torch.Tensor, params: torch.Tensor, w1: torch.Tensor, b1: Optional[torch.Tensor], w2: torch.Tensor,
                  b2: Optional[torch.Tensor], meta, activation: str = "exp", return_output: bool = False):
    """Differentiable fused query: (sigma [N], decoder output [N, n_out] or None).  Gradients flow to `params` (LoTD tables), w1, b1, w2, b2.
    Replaces encoding -> density_decoder -> activation of LoTDNeRF.forward_density (lotd_nerf.py:136-178) in a training step."""
    if activation not in _ACT:
        raise RuntimeError(f"fused_density: activation must be one of {sorted(_ACT)}, got {activation!r}")
    return _FusedDensityFn.apply(x, params, w1, b1, w2, b2, meta, _ACT[activation], bool(return_output))


class FusedDensityMLP(torch.nn.Module):
    """Trainable density decoder Linear(32, 64) -> ReLU -> Linear(64, n_out) -> activation(out[..., 0]) fused with the LoTD encoder both ways.
    Initialisation follows torch.nn.Linear; `forward(x, params)` returns sigma (and the raw decoder output with `return_output=True`)."""

    def __init__(self, meta, n_out: int = 1, activation: str = "exp", bias: bool = True, device=None):
        super().__init__()
        if not (1 <= n_out <= 16):
            raise RuntimeError("FusedDensityMLP: 1 <= n_out <= 16")
        l1, l2 = torch.nn.Linear(32, 64, bias=bias, device=device), torch.nn.Linear(64, n_out, bias=bias, device=device)
        self.meta, self.activation = meta, activation
        self.w1, self.w2 = torch.nn.Parameter(l1.weight.detach().clone()), torch.nn.Parameter(l2.weight.detach().clone())
        self.b1 = torch.nn.Parameter(l1.bias.detach().clone()) if bias else None
        self.b2 = torch.nn.Parameter(l2.bias.detach().clone()) if bias else None

    def forward(self, x: torch.Tensor, params: torch.Tensor, return_output: bool = False):
        sigma, out = fused_density(x, params, self.w1, self.b1, self.w2, self.b2, self.meta, self.activation, return_output)
        return (sigma, out) if return_output else sigma
