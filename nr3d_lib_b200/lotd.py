"""Host-side mirror of the reference's LoTD operator interface
(nr3d_lib/models/grid_encodings/lotd/lotd.py:31-458): same class / function names, argument meaning and error
behaviour, bound to the B200 kernels through ``nr3d_lib_b200.bindings._lotd``.

Written for this repository (not a copy): the three autograd Functions share one set of helpers.
"""
from enum import Enum
from math import prod
from typing import List, Literal, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn
from torch.autograd.function import FunctionCtx, once_differentiable

from .bindings import _lotd as _backend

__all__ = ['LoDType', 'generate_meta', 'LoTDFunction', 'LoTDFunctionFwdDydx', 'LoTDFunctionBwdDydx', 'lotd_encoding',
           'lotd_encoding_fwd_dydx', 'lotd_encoding_bwd_dydx', 'lotd_get_grid_index', 'LoTD']

_X_EPS = 1.0e-6  # inputs are clamped to [eps, 1-eps] before encoding (reference lotd.py:68,144,206)


class LoDType(Enum):  # reference lotd.py:31-38
    Dense = int(_backend.LoDType.Dense)
    VectorMatrix = int(_backend.LoDType.VectorMatrix)
    CP = int(_backend.LoDType.CP)
    CPfast = int(_backend.LoDType.CPfast)
    NPlaneMul = int(_backend.LoDType.NPlaneMul)
    NPlaneSum = int(_backend.LoDType.NPlaneSum)
    Hash = int(_backend.LoDType.Hash)


def generate_meta(n_input_dim: int, lod_res: List[int], lod_n_feats: Union[int, List[int]], lod_types: Union[str, List[str]],
                  hashmap_size: int = None, use_smooth_step=False):
    """reference lotd.py:40-45"""
    if isinstance(lod_n_feats, int):
        lod_n_feats = [lod_n_feats] * len(lod_res)
    if isinstance(lod_types, str):
        lod_types = [lod_types] * len(lod_res)
    return _backend.LoDMeta(n_input_dim, lod_res, lod_n_feats, lod_types, hashmap_size, use_smooth_step)


def _prep(x: torch.Tensor, bidx: Optional[torch.Tensor]):
    prefix = x.shape[:-1]
    x = x.clamp(_X_EPS, 1 - _X_EPS)
    bidx = None if bidx is None else bidx.contiguous().long().flatten()
    return prefix, x, bidx


def _stash(ctx, meta, prefix, batch_data_size, loss_scale, max_level):
    ctx.meta, ctx.prefix, ctx.batch_data_size, ctx.loss_scale, ctx.max_level = meta, prefix, batch_data_size, loss_scale, max_level


def _scaled(t, s):
    """t * s without the extra pass over the [N, n_enc] gradient when the loss scale is 1 (fp32 tables): same values."""
    return t if s == 1.0 else t * s


def _unscaled(t, s):
    return t if (t is None or s == 1.0) else t / s


def _first_order_backward(ctx, dL_dy, need_dx: bool, need_dgrid: bool):
    x, grid, dy_dx, bidx, batch_offsets = ctx.saved_tensors
    ls = ctx.loss_scale
    dL_dx, dL_dgrid = _backend.lod_bwd(ctx.meta, _scaled(dL_dy.flatten(0, -2), ls), x.flatten(0, -2), grid, dy_dx,
                                       None if bidx is None else bidx.flatten(), batch_offsets, ctx.batch_data_size, ctx.max_level,
                                       need_dx, need_dgrid)
    dL_dx = None if dL_dx is None else _unscaled(dL_dx.unflatten(0, ctx.prefix), ls)
    dL_dgrid = _unscaled(dL_dgrid, ls)
    return dL_dx, dL_dgrid


class LoTDFunction(torch.autograd.Function):
    """First-order LoTD encoding (reference lotd.py:48-119)."""

    @staticmethod
    def forward(ctx: FunctionCtx, meta, x: torch.Tensor, grid: torch.Tensor, bidx=None, batch_offsets=None, batch_data_size=None,
                loss_scale=1.0, max_level=None):
        ctx.set_materialize_grads(False)
        prefix, x, bidx = _prep(x, bidx)
        y, dy_dx = _backend.lod_fwd(meta, x.flatten(0, -2), grid, bidx, batch_offsets, batch_data_size, max_level, ctx.needs_input_grad[1])
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[2]:
            ctx.save_for_backward(x, grid, dy_dx, bidx, batch_offsets)
            _stash(ctx, meta, prefix, batch_data_size, loss_scale, max_level)
        return y.unflatten(0, prefix)

    @staticmethod
    @once_differentiable
    def backward(ctx: FunctionCtx, dL_dy):
        dL_dx = dL_dgrid = None
        if dL_dy is not None and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
            dL_dx, dL_dgrid = _first_order_backward(ctx, dL_dy, ctx.needs_input_grad[1], ctx.needs_input_grad[2])
        return None, dL_dx, dL_dgrid, None, None, None, None, None


class LoTDFunctionFwdDydx(torch.autograd.Function):
    """Forward that also returns dy/dx; pair it with LoTDFunctionBwdDydx for nablas (reference lotd.py:121-191)."""

    @staticmethod
    def forward(ctx: FunctionCtx, meta, x, grid, bidx=None, batch_offsets=None, batch_data_size=None, loss_scale=1.0, max_level=None,
                need_dL_dinput: Optional[bool] = None):
        if need_dL_dinput is None:
            need_dL_dinput = torch.is_grad_enabled() and x.requires_grad
        ctx.set_materialize_grads(False)
        prefix, x, bidx = _prep(x, bidx)
        y, dy_dx = _backend.lod_fwd(meta, x.flatten(0, -2), grid, bidx, batch_offsets, batch_data_size, max_level, True)
        ctx.save_for_backward(x, grid, dy_dx, bidx, batch_offsets)
        _stash(ctx, meta, prefix, batch_data_size, loss_scale, max_level)
        ctx.need_dL_dinput = need_dL_dinput
        ctx.mark_non_differentiable(dy_dx)
        return y.unflatten(0, prefix), dy_dx

    @staticmethod
    @once_differentiable
    def backward(ctx: FunctionCtx, dL_dy, _unused):
        dL_dx = dL_dgrid = None
        if dL_dy is not None:
            with torch.no_grad():
                dL_dx, dL_dgrid = _first_order_backward(ctx, dL_dy, ctx.need_dL_dinput, ctx.needs_input_grad[2])
        return None, dL_dx, dL_dgrid, None, None, None, None, None, None


class LoTDFunctionBwdDydx(torch.autograd.Function):
    """dL/dx from (dL/dy, dy/dx) with second-order backward onto dL/dy, params (and x) (reference lotd.py:193-268)."""

    @staticmethod
    def forward(ctx: FunctionCtx, meta, dL_dy, x, grid, dy_dx, bidx, batch_offsets, batch_data_size, loss_scale, max_level, grad_guard):
        ctx.set_materialize_grads(False)
        prefix, x, bidx = _prep(x, bidx)
        dL_dx, _ = _backend.lod_bwd(meta, _scaled(dL_dy.flatten(0, -2), loss_scale), x.flatten(0, -2), grid, dy_dx, bidx, batch_offsets,
                                    batch_data_size, max_level, True, False)
        if ctx.needs_input_grad[1] or ctx.needs_input_grad[3]:
            ctx.save_for_backward(dL_dy, x, grid, dy_dx.contiguous(), bidx, batch_offsets)
            _stash(ctx, meta, prefix, batch_data_size, loss_scale, max_level)
            ctx.grad_guard = grad_guard
        return None if dL_dx is None else _unscaled(dL_dx.unflatten(0, prefix), loss_scale)

    @staticmethod
    @once_differentiable
    def backward(ctx: FunctionCtx, dL_ddLdx):
        g_dLdy = g_x = g_grid = None
        if dL_ddLdx is not None:
            dL_dy, x, grid, dy_dx, bidx, batch_offsets = ctx.saved_tensors
            prefix, ls = x.shape[:-1], ctx.loss_scale
            g_dLdy, g_grid, g_x = _backend.lod_bwd_bwd_input(
                ctx.meta, dL_ddLdx.flatten(0, -2), _scaled(dL_dy.flatten(0, -2), ls), x.flatten(0, -2), grid, dy_dx,
                None if bidx is None else bidx.flatten(), batch_offsets, ctx.batch_data_size, ctx.max_level,
                ctx.needs_input_grad[1], ctx.needs_input_grad[3], False)
            g_dLdy = None if g_dLdy is None else g_dLdy.unflatten(0, prefix)
            g_grid = _unscaled(g_grid, ls)
            g_x = None if g_x is None else _unscaled(g_x.unflatten(0, prefix), ls)
            if ctx.grad_guard is not None and (g_grid is not None or g_dLdy is not None):
                ctx.grad_guard.custom_grad_clip_step(dL_ddLdx, dy_dx, g_grid, g_dLdy)
        return None, g_dLdy, g_x, g_grid, None, None, None, None, None, None, None


def _batch_data_size(input, input_batched):
    return prod(input.shape[1:-1]) if input_batched else 0


def _loss_scale(params):
    return 128.0 if params.dtype == torch.float16 else 1.0


def lotd_encoding(input, params, bidx=None, batch_offsets=None, input_batched=False, max_level=None, meta=None, n_input_dim=None,
                  lod_res=None, lod_n_feats=None, lod_types=None):
    """reference lotd.py:270-282"""
    if meta is None:
        meta = generate_meta(n_input_dim, lod_res, lod_n_feats, lod_types)
    if input_batched:
        bidx = None
    return LoTDFunction.apply(meta, input, params, bidx, batch_offsets, _batch_data_size(input, input_batched), _loss_scale(params), max_level)


def lotd_encoding_fwd_dydx(input, params, bidx=None, batch_offsets=None, input_batched=False, max_level=None,
                           need_dL_dinput: Optional[bool] = None, meta=None, n_input_dim=None, lod_res=None, lod_n_feats=None,
                           lod_types=None):
    """reference lotd.py:284-298"""
    if need_dL_dinput is None:
        need_dL_dinput = torch.is_grad_enabled() and input.requires_grad
    if meta is None:
        meta = generate_meta(n_input_dim, lod_res, lod_n_feats, lod_types)
    if input_batched:
        bidx = None
    output, dy_dx = LoTDFunctionFwdDydx.apply(meta, input, params, bidx, batch_offsets, _batch_data_size(input, input_batched),
                                              _loss_scale(params), max_level, need_dL_dinput)
    return output, dy_dx, meta


def lotd_encoding_bwd_dydx(meta, dL_dy, dy_dx, input, params, bidx=None, batch_offsets=None, input_batched=False, max_level=None):
    """reference lotd.py:300-309"""
    if input_batched:
        bidx = None
    return LoTDFunctionBwdDydx.apply(meta, dL_dy, input, params, dy_dx, bidx, batch_offsets, _batch_data_size(input, input_batched),
                                     _loss_scale(params), max_level, None)


def lotd_get_grid_index(meta, input, bidx=None, batch_offsets=None, input_batched=False, max_level=None) -> torch.LongTensor:
    """reference lotd.py:311-319"""
    if input_batched:
        bidx = None
    return _backend.lod_get_grid_index(meta, input, bidx, batch_offsets, _batch_data_size(input, input_batched), max_level)


class LoTD(nn.Module):
    """Parameter-free LoTD encoder module: owns the meta, receives the flattened params per call (reference lotd.py:321-502)."""

    def __init__(self, in_features: Literal[2, 3], lod_res: Union[List[int], List[List[int]]], lod_n_feats: Union[int, List[int]],
                 lod_types: Union[str, List[str]], hashmap_size: int = None, log2_hashmap_size: int = None, use_smooth_step=False,
                 use_profile=False, dtype=torch.half, device=None):
        super().__init__()
        assert dtype == torch.float or dtype == torch.float16, "dtype must be one of torch.float or torch.float16"
        self.params = dict(in_features=in_features, lod_res=lod_res, lod_n_feats=lod_n_feats, lod_types=lod_types,
                           hashmap_size=hashmap_size, log2_hashmap_size=log2_hashmap_size, use_smooth_step=use_smooth_step,
                           use_profile=use_profile, dtype=dtype, device=device)
        self.loss_scale = 128.0 if dtype == torch.float16 else 1.0
        self.dtype = dtype
        if log2_hashmap_size is not None:
            assert hashmap_size is None, "Do not specify `hashmap_size` when `log2_hashmap_size` is already specified."
            hashmap_size = 2 ** log2_hashmap_size
        self.meta = generate_meta(in_features, lod_res, lod_n_feats, lod_types, hashmap_size, use_smooth_step)

    in_features = property(lambda self: self.meta.n_dims_to_encode)
    out_features = property(lambda self: self.meta.n_encoded_dims)
    n_levels = property(lambda self: self.meta.n_levels)
    n_params = property(lambda self: self.meta.n_params)
    level_res_multidim = property(lambda self: self.meta.level_res_multidim)
    level_types_str = property(lambda self: self.meta.level_types_str)
    level_sizes = property(lambda self: self.meta.level_sizes)
    level_offsets = property(lambda self: self.meta.level_offsets)
    level_n_feats = property(lambda self: self.meta.level_n_feats)
    level_n_params = property(lambda self: self.meta.level_n_params)

    @property
    def level_res(self) -> Optional[List[int]]:
        r = np.array(self.meta.level_res_multidim)
        return r[:, 0].tolist() if (r == r[:, :1]).all() else None

    @property
    def level_types(self) -> List[LoDType]:
        return [LoDType(tp) for tp in self.meta.level_types]

    def _bds(self, input, bidx, input_batched):
        if input_batched:
            assert bidx is None, 'bidx is only taken care of when input is not batched.'
            return prod(input.shape[1:-1])
        return 0

    def forward(self, input: torch.Tensor, params: torch.Tensor, bidx: torch.Tensor = None, batch_offsets: torch.Tensor = None,
                input_batched=False, max_level: int = None) -> torch.Tensor:
        return LoTDFunction.apply(self.meta, input, params.to(self.dtype), bidx, batch_offsets, self._bds(input, bidx, input_batched),
                                  self.loss_scale, max_level)

    def forward_dydx(self, input: torch.Tensor, params: torch.Tensor, bidx: torch.Tensor = None, batch_offsets: torch.Tensor = None,
                     input_batched=False, max_level: int = None, need_dL_dinput: Optional[bool] = None) -> Tuple[torch.Tensor, torch.Tensor]:
        if need_dL_dinput is None:
            need_dL_dinput = torch.is_grad_enabled() and input.requires_grad
        return LoTDFunctionFwdDydx.apply(self.meta, input, params.to(self.dtype), bidx, batch_offsets,
                                         self._bds(input, bidx, input_batched), self.loss_scale, max_level, need_dL_dinput)

    def backward_dydx(self, dL_dy: torch.Tensor, dy_dx: torch.Tensor, input: torch.Tensor, params: torch.Tensor, bidx: torch.Tensor = None,
                      batch_offsets: torch.Tensor = None, input_batched=False, max_level: int = None, grad_guard=None) -> torch.Tensor:
        return LoTDFunctionBwdDydx.apply(self.meta, dL_dy, input, params.to(self.dtype), dy_dx, bidx, batch_offsets,
                                         self._bds(input, bidx, input_batched), self.loss_scale, max_level, grad_guard)

    def __getstate__(self):
        return self.params

    def __setstate__(self, state_dict):
        self.__init__(**state_dict)

    def extra_repr(self) -> str:
        ele = {torch.float32: 4, torch.float16: 2}[self.dtype]
        m = self.meta
        return (f"in_dim={m.n_dims_to_encode}, out_dim={m.n_encoded_dims}, num_levels={m.n_levels}, num_params={m.n_params}, "
                f"params_size={(m.n_params * ele) / (1024 ** 2):.3f} MiB, dtype={self.dtype}\n"
                f"lod_res={m.level_res}\nlod_n_feats={m.level_n_feats}\nlod_types={m.level_types_str}\nlod_n_params={m.level_n_params}")
