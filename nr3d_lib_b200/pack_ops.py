"""Host-side mirror of the reference's pack_ops operator interface (nr3d_lib/graphics/pack_ops/pack_ops.py:97-392,730-747):
same function names / argument meaning / gradients, bound to the B200 kernels via ``nr3d_lib_b200.bindings._pack_ops``.

A "pack" is a contiguous run of rows of a flat tensor; ``pack_infos`` is int64 ``[num_packs, 2] = (first index, length)``.
"""
from numbers import Number
from typing import Tuple, Union

import torch
from torch.autograd.function import once_differentiable

from .bindings import _pack_ops as _backend

__all__ = ['packed_sort_inplace', 'packed_sort', 'packed_searchsorted', 'packed_searchsorted_packed_vals', 'packed_invert_cdf',
           'packed_matmul', 'merge_two_packs_sorted_aligned', 'packed_sum', 'packed_mean', 'packed_cumprod', 'packed_cumsum', 'packed_diff', 'packed_backward_diff',
           'packed_alpha_to_vw', 'packed_volume_render_compression', 'packed_add', 'packed_sub', 'packed_mul', 'packed_div',
           'packed_gt', 'packed_geq', 'packed_lt', 'packed_leq', 'packed_eq', 'packed_neq', 'interleave_arange_simple',
           'interleave_arange', 'interleave_linstep', 'interleave_linspace', 'interleave_sample_step_wrt_depth_clamped',
           'interleave_sample_step_wrt_depth_in_packed_segments', 'get_pack_infos_from_boundary', 'get_pack_infos_from_first', 'get_pack_infos_from_n', 'get_pack_infos_from_batch',
           'mark_pack_boundaries', 'expand_pack_boundary']


# ---------------------------------------------------------------------------------------------------------------
# reductions / scans
# ---------------------------------------------------------------------------------------------------------------
class PackedSum(torch.autograd.Function):  # reference pack_ops.py:97-111
    @staticmethod
    def forward(ctx, feats, pack_infos):
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(pack_infos)
        return _backend.packed_sum(feats, pack_infos)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (pack_infos,) = ctx.saved_tensors
        grad = grad_output.repeat_interleave(pack_infos[..., 1], dim=0) if ctx.needs_input_grad[0] else None
        return grad, None


def packed_sum(feats: torch.Tensor, pack_infos: torch.LongTensor) -> torch.Tensor:
    if feats.requires_grad:
        return PackedSum.apply(feats.contiguous(), pack_infos)
    return _backend.packed_sum(feats.contiguous(), pack_infos)


def packed_mean(feats: torch.Tensor, pack_infos: torch.LongTensor) -> torch.Tensor:
    return packed_sum(feats, pack_infos) / (pack_infos[:, 1] + 1e-8)


class PackedCumprod(torch.autograd.Function):  # reference pack_ops.py:124-143
    @staticmethod
    def forward(ctx, feats, pack_infos, exclusive, reverse):
        prod = _backend.packed_cumprod(feats, pack_infos, exclusive, reverse)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(feats, pack_infos, prod)
            ctx.flags = (exclusive, reverse)
        return prod

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        grad_feats = None
        if ctx.needs_input_grad[0]:
            feats, pack_infos, prod = ctx.saved_tensors
            exclusive, reverse = ctx.flags
            out = _backend.packed_cumsum((prod * grad_output).contiguous(), pack_infos, exclusive, not reverse)
            grad_feats = out / feats  # approximate gradient, consistent with TensorFlow (reference comment)
            grad_feats[grad_feats.isnan()] = 0
        return grad_feats, None, None, None


def packed_cumprod(feats, pack_infos, exclusive: bool = False, reverse: bool = False) -> torch.Tensor:
    """Pack-wise cumulative product; ``exclusive`` = right-shifted with a leading 1 (documented semantics)."""
    if feats.requires_grad:
        return PackedCumprod.apply(feats.contiguous(), pack_infos, exclusive, reverse)
    return _backend.packed_cumprod(feats.contiguous(), pack_infos, exclusive, reverse)


class PackedCumsum(torch.autograd.Function):  # reference pack_ops.py:161-175
    @staticmethod
    def forward(ctx, feats, pack_infos, exclusive, reverse):
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(pack_infos)
            ctx.flags = (exclusive, reverse)
        return _backend.packed_cumsum(feats, pack_infos, exclusive, reverse)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        (pack_infos,) = ctx.saved_tensors
        exclusive, reverse = ctx.flags
        return _backend.packed_cumsum(grad_output.contiguous(), pack_infos, exclusive, not reverse), None, None, None


def packed_cumsum(feats, pack_infos, exclusive: bool = False, reverse: bool = False) -> torch.Tensor:
    """Pack-wise cumulative sum; ``exclusive`` = right-shifted with a leading 0."""
    if feats.requires_grad:
        return PackedCumsum.apply(feats.contiguous(), pack_infos, exclusive, reverse)
    return _backend.packed_cumsum(feats.contiguous(), pack_infos, exclusive, reverse)


# ---------------------------------------------------------------------------------------------------------------
# finite differences
# ---------------------------------------------------------------------------------------------------------------
def _ends(pack_infos):
    first, n = pack_infos[..., 0], pack_infos[..., 1]
    return first, n, first + n - 1


class PackedDiff(torch.autograd.Function):  # forward diff: d_i = f_{i+1} - f_i  (reference pack_ops.py:193-218)
    @staticmethod
    def forward(ctx, feats, pack_infos, pack_appends, pack_last_fill):
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(pack_infos)
            ctx.flags = (pack_appends is not None, pack_last_fill is not None)
        return _backend.packed_diff(feats, pack_infos, pack_appends, pack_last_fill)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        has_append, has_last_fill = ctx.flags
        (pack_infos,) = ctx.saved_tensors
        first, n, last = _ends(pack_infos)
        grad_feat = None
        if ctx.needs_input_grad[0]:
            grad_feat = -1 * _backend.packed_backward_diff(grad_output.contiguous(), pack_infos, None, grad_output[first].contiguous())
            if not has_append:
                second_last = grad_output[last - 1]
                grad_feat[last] = torch.where(n.view([-1] + [1] * (second_last.dim() - 1)) > 1, second_last, grad_output.new_tensor([0.]))
        grad_append = grad_output[last] if (has_append and ctx.needs_input_grad[2]) else None
        grad_last_fill = grad_output[last] if (has_last_fill and ctx.needs_input_grad[3]) else None
        return grad_feat, None, grad_append, grad_last_fill


def packed_diff(feats, pack_infos, pack_appends: torch.Tensor = None, pack_last_fill: torch.Tensor = None) -> torch.Tensor:
    if feats.requires_grad:
        return PackedDiff.apply(feats.contiguous(), pack_infos, pack_appends, pack_last_fill)
    return _backend.packed_diff(feats.contiguous(), pack_infos, pack_appends, pack_last_fill)


class PackedBackwardDiff(torch.autograd.Function):  # backward diff: d_i = f_i - f_{i-1}  (reference pack_ops.py:227-252)
    @staticmethod
    def forward(ctx, feats, pack_infos, pack_prepends, pack_first_fill):
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(pack_infos)
            ctx.flags = (pack_prepends is not None, pack_first_fill is not None)
        return _backend.packed_backward_diff(feats, pack_infos, pack_prepends, pack_first_fill)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_output):
        has_prepend, has_first_fill = ctx.flags
        (pack_infos,) = ctx.saved_tensors
        first, n, last = _ends(pack_infos)
        grad_feat = None
        if ctx.needs_input_grad[0]:
            grad_feat = -1 * _backend.packed_diff(grad_output.contiguous(), pack_infos, None, (-grad_output[last]).contiguous())
            if not has_prepend:
                second = grad_output[first + 1]
                grad_feat[first] = torch.where(n.view([-1] + [1] * (second.dim() - 1)) > 1, -second, grad_output.new_tensor([0.]))
        grad_prepend = -grad_output[first] if (has_prepend and ctx.needs_input_grad[2]) else None
        grad_first_fill = grad_output[first] if (has_first_fill and ctx.needs_input_grad[3]) else None
        return grad_feat, None, grad_prepend, grad_first_fill


def packed_backward_diff(feats, pack_infos, pack_prepends: torch.Tensor = None, pack_first_fill: torch.Tensor = None) -> torch.Tensor:
    if feats.requires_grad:
        return PackedBackwardDiff.apply(feats.contiguous(), pack_infos, pack_prepends, pack_first_fill)
    return _backend.packed_backward_diff(feats.contiguous(), pack_infos, pack_prepends, pack_first_fill)


# ---------------------------------------------------------------------------------------------------------------
# volume rendering
# ---------------------------------------------------------------------------------------------------------------
class PackedAlphaToVW(torch.autograd.Function):  # reference pack_ops.py:260-277
    @staticmethod
    def forward(ctx, alphas, pack_infos, early_stop_eps, alpha_thre):
        weights = _backend.packed_alpha_to_vw_forward(alphas, pack_infos, early_stop_eps, alpha_thre, False)[0]
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(alphas, pack_infos, weights)
            ctx.early_stop_eps, ctx.alpha_thre = early_stop_eps, alpha_thre
        return weights

    @staticmethod
    def backward(ctx, grad_weights):
        alphas, pack_infos, weights = ctx.saved_tensors
        grad_alphas = _backend.packed_alpha_to_vw_backward(weights, grad_weights.contiguous(), alphas, pack_infos, ctx.early_stop_eps,
                                                           ctx.alpha_thre)
        return grad_alphas, None, None, None


def packed_alpha_to_vw(alpha: torch.Tensor, pack_infos: torch.LongTensor, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0) -> torch.Tensor:
    """alpha -> front-to-back compositing weights w_j = alpha_j * prod_{k<j}(1 - alpha_k) per pack."""
    if alpha.requires_grad:
        return PackedAlphaToVW.apply(alpha.contiguous(), pack_infos, early_stop_eps, alpha_thre)
    return _backend.packed_alpha_to_vw_forward(alpha.contiguous(), pack_infos, early_stop_eps, alpha_thre, False)[0]


@torch.no_grad()
def packed_volume_render_compression(alpha, pack_infos, early_stop_eps: float = 1e-4, alpha_thre: float = 0.0):
    """reference pack_ops.py:285-291 -> (nidx_useful, compact_pack_infos, pidx)"""
    _, compact_pack_infos, compact_selector = _backend.packed_alpha_to_vw_forward(alpha.contiguous(), pack_infos, early_stop_eps, alpha_thre, True)
    pidx = compact_selector.nonzero().long()[..., 0]
    nidx_useful = (compact_pack_infos[:, 1] > 0).nonzero()[..., 0]
    return nidx_useful, compact_pack_infos[nidx_useful].long(), pidx


# ---------------------------------------------------------------------------------------------------------------
# arithmetic / logic against one operand per pack
# ---------------------------------------------------------------------------------------------------------------
class PackedAdd(torch.autograd.Function):  # reference pack_ops.py:293-314
    @staticmethod
    def forward(ctx, feats, other, pack_infos):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(pack_infos)
        return _backend.packed_add(feats, other, pack_infos)

    @staticmethod
    def backward(ctx, grad_out):
        if grad_out is None:
            return None, None, None
        (pack_infos,) = ctx.saved_tensors
        grad_in = grad_out if ctx.needs_input_grad[0] else None
        grad_other = PackedSum.apply(grad_out.contiguous(), pack_infos) if ctx.needs_input_grad[1] else None
        return grad_in, grad_other, None


class PackedSub(torch.autograd.Function):  # reference pack_ops.py:316-336
    @staticmethod
    def forward(ctx, feats, other, pack_infos):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(pack_infos)
        return _backend.packed_sub(feats, other, pack_infos)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        if grad_out is None:
            return None, None, None
        (pack_infos,) = ctx.saved_tensors
        grad_in = grad_out if ctx.needs_input_grad[0] else None
        grad_other = -1 * _backend.packed_sum(grad_out.contiguous(), pack_infos) if ctx.needs_input_grad[1] else None
        return grad_in, grad_other, None


class PackedMul(torch.autograd.Function):  # reference pack_ops.py:338-358
    @staticmethod
    def forward(ctx, feats, other, pack_infos):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(feats, other, pack_infos)
        return _backend.packed_mul(feats, other, pack_infos)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        if grad_out is None:
            return None, None, None
        feats, other, pack_infos = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_in = _backend.packed_mul(grad_out, other, pack_infos) if ctx.needs_input_grad[0] else None
        grad_other = _backend.packed_sum(grad_out * feats, pack_infos) if ctx.needs_input_grad[1] else None
        return grad_in, grad_other, None


class PackedDiv(torch.autograd.Function):  # reference pack_ops.py:360-382
    @staticmethod
    def forward(ctx, feats, other, pack_infos):
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(feats, other, pack_infos)
        return _backend.packed_div(feats, other, pack_infos)

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        if grad_out is None:
            return None, None, None
        feats, other, pack_infos = ctx.saved_tensors
        grad_out = grad_out.contiguous()
        grad_in = _backend.packed_div(grad_out, other, pack_infos) if ctx.needs_input_grad[0] else None
        grad_other = None
        if ctx.needs_input_grad[1]:  # d(f/o)/do = -f / o^2
            grad_other = _backend.packed_sum(_backend.packed_div(-grad_out * feats, other * other, pack_infos), pack_infos)
        return grad_in, grad_other, None


def packed_add(feats, other, pack_infos): return PackedAdd.apply(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
def packed_sub(feats, other, pack_infos): return PackedSub.apply(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
def packed_mul(feats, other, pack_infos): return PackedMul.apply(feats.contiguous(), other.contiguous(), pack_infos.contiguous())


def packed_div(feats, other, pack_infos):
    """Pack-wise division ``feats / other`` with ``other`` of shape [num_packs(, feat_dim)]."""
    return PackedDiv.apply(feats.contiguous(), other.contiguous(), pack_infos.contiguous())


def packed_gt(feats, other, pack_infos): return _backend.packed_gt(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
def packed_geq(feats, other, pack_infos): return _backend.packed_geq(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
def packed_lt(feats, other, pack_infos): return _backend.packed_lt(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
def packed_leq(feats, other, pack_infos): return _backend.packed_leq(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
def packed_eq(feats, other, pack_infos): return _backend.packed_eq(feats.contiguous(), other.contiguous(), pack_infos.contiguous())
def packed_neq(feats, other, pack_infos): return _backend.packed_neq(feats.contiguous(), other.contiguous(), pack_infos.contiguous())


# ---------------------------------------------------------------------------------------------------------------
# sort / search / inverse-CDF sampling / merge (reference pack_ops.py:74-95, 398-407, 550-570)
# ---------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def packed_sort_inplace(vals: torch.Tensor, pack_infos: torch.LongTensor, return_idx=True) -> torch.Tensor:
    return _backend.packed_sort_qsort(vals.contiguous(), pack_infos, return_idx)


def packed_sort(vals: torch.Tensor, pack_infos: torch.LongTensor):
    indices = packed_sort_inplace(vals.data.clone(), pack_infos, return_idx=True)
    return vals[indices], indices


@torch.no_grad()
def packed_searchsorted(bins: torch.Tensor, vals: torch.Tensor, pack_infos: torch.LongTensor) -> torch.Tensor:
    """Search a batch (vals [num_packs, n]) in a sorted pack (bins)."""
    return _backend.packed_searchsorted(bins.contiguous(), vals.contiguous(), pack_infos)


@torch.no_grad()
def packed_searchsorted_packed_vals(bins, pack_infos, vals, u_pack_infos) -> torch.Tensor:
    """Search a pack (vals, u_pack_infos) in a sorted pack (bins, pack_infos)."""
    return _backend.packed_searchsorted_packed_vals(bins.contiguous(), pack_infos.contiguous(), vals.contiguous(), u_pack_infos)


@torch.no_grad()
def packed_invert_cdf(bins, cdfs, u_vals, pack_infos) -> Tuple[torch.Tensor, torch.Tensor]:
    return _backend.packed_invert_cdf(bins.contiguous(), cdfs.contiguous(), u_vals.contiguous(), pack_infos)


def packed_matmul(feats: torch.Tensor, other: torch.Tensor, pack_infos: torch.LongTensor) -> torch.Tensor:
    """Pack-wise left multiplication ``other[pack] @ feat``; differentiable torch formulation as in the reference (pack_ops.py:407)."""
    return (feats.unsqueeze(-2) * torch.repeat_interleave(other, pack_infos[:, 1], dim=0)).sum(-1)


def merge_two_packs_sorted_aligned(vals_a, pack_infos_a, vals_b, pack_infos_b, b_sorted=True, return_val=False):
    """Merge two aligned sorted packs (vals_b may be unsorted): positions of a / b in the merged pack (reference pack_ops.py:550-570)."""
    pidx_a, pidx_b, pack_infos = _backend.try_merge_two_packs_sorted_aligned(vals_a.contiguous(), pack_infos_a, vals_b.contiguous(), pack_infos_b, b_sorted)
    if return_val:
        val = vals_a.new_empty([vals_a.numel() + vals_b.numel()])
        val[pidx_a], val[pidx_b] = vals_a, vals_b
        return val, pack_infos
    return pidx_a, pidx_b, pack_infos


# ---------------------------------------------------------------------------------------------------------------
# generators
# ---------------------------------------------------------------------------------------------------------------
@torch.no_grad()
def interleave_arange_simple(stop: torch.Tensor, return_idx: bool = True):
    ret = _backend.interleave_arange(stop.contiguous(), return_idx)
    return ret if return_idx else ret[0]


@torch.no_grad()
def interleave_linstep(start: torch.Tensor, num_steps: torch.Tensor, step_size: Union[torch.Tensor, Number], return_idx: bool = True):
    ret = _backend.interleave_linstep(start.contiguous(), num_steps.contiguous(),
                                      step_size.contiguous() if isinstance(step_size, torch.Tensor) else step_size, return_idx)
    return ret if return_idx else ret[0]


@torch.no_grad()
def interleave_arange(start: torch.Tensor, stop: torch.Tensor, step_size: Union[torch.Tensor, Number], return_idx: bool = True):
    num_steps = stop.subtract(start).div(step_size).ceil().long()
    return interleave_linstep(start, num_steps, step_size, return_idx)


@torch.no_grad()
def interleave_linspace(start: torch.Tensor, stop: torch.Tensor, num_steps: Union[torch.Tensor, Number], return_idx: bool = True):
    step_size = (stop - start) / (num_steps - 1)
    if not isinstance(num_steps, torch.Tensor):
        num_steps = torch.full(start.shape, num_steps, device=start.device, dtype=torch.long)
    return interleave_linstep(start, num_steps, step_size, return_idx=return_idx)


@torch.no_grad()
def interleave_sample_step_wrt_depth_clamped(near, far, max_steps: int = 512, dt_gamma: float = 0.01, min_step_size: float = 0.01,
                                             max_step_size: float = 1.0, step_size_factor: float = 1.0, perturb=False):
    """reference pack_ops.py:463-485 -> (t_samples, deltas, ridx, pack_infos)"""
    t_samples, deltas, ridx, pack_infos = _backend.interleave_sample_step_wrt_depth_clamped(
        near.contiguous(), far.contiguous(), max_steps, dt_gamma * step_size_factor, min_step_size * step_size_factor,
        max_step_size * step_size_factor)
    if perturb:
        noise = torch.rand_like(deltas)
        t_samples = torch.addcmul(t_samples, noise, deltas)
        last = pack_infos[..., 0] + pack_infos[..., 1] - 1
        deltas = t_samples.diff(append=t_samples.new_empty([1])).index_put_((last,), deltas[last])
    return t_samples, deltas, ridx, pack_infos


@torch.no_grad()
def interleave_sample_step_wrt_depth_in_packed_segments(near, far, entry: torch.Tensor, exit: torch.Tensor, seg_pack_infos: torch.Tensor,
                                                        max_steps: int = 512, dt_gamma: float = 0.01, min_step_size: float = 0.01,
                                                        max_step_size: float = 1e10, step_size_factor: float = 1.0, perturb=False):
    """reference pack_ops.py:476-499 -> (t_samples, deltas, ridx, ray_pack_infos, sidx, out_seg_pack_infos)"""
    num_rays = seg_pack_infos.shape[0]
    near = entry.new_full([num_rays], near) if not isinstance(near, torch.Tensor) else near
    far = entry.new_full([num_rays], far) if not isinstance(far, torch.Tensor) else far
    t_samples, deltas, sidx, ridx, ray_pack_infos = _backend.interleave_sample_step_wrt_depth_in_packed_segments(
        near.contiguous(), far.contiguous(), entry.contiguous(), exit.contiguous(), seg_pack_infos.contiguous(), max_steps,
        dt_gamma * step_size_factor, min_step_size * step_size_factor, max_step_size * step_size_factor)
    out_seg_pack_infos = get_pack_infos_from_boundary(mark_pack_boundaries(sidx))
    if perturb:
        noise = torch.rand_like(deltas)
        t_samples = torch.addcmul(t_samples, noise, deltas)
        last = ray_pack_infos[..., 0] + ray_pack_infos[..., 1] - 1
        deltas = t_samples.diff(append=t_samples.new_empty([1])).index_put_((last,), deltas[last])
    return t_samples, deltas, ridx, ray_pack_infos, sidx, out_seg_pack_infos


# ---------------------------------------------------------------------------------------------------------------
# pack-info helpers (reference pack_ops.py:730-747)
# ---------------------------------------------------------------------------------------------------------------
def mark_pack_boundaries(pack_ids: torch.Tensor) -> torch.Tensor:
    return _backend.mark_pack_boundaries_cuda(pack_ids.contiguous()).bool()


@torch.no_grad()
def expand_pack_boundary(pack_boundary: torch.Tensor, num_samples: int):
    big = torch.zeros(pack_boundary.shape[0] * num_samples, device=pack_boundary.device, dtype=torch.bool)
    big[pack_boundary.nonzero().long() * num_samples] = 1
    return big


@torch.no_grad()
def get_pack_infos_from_first(first_inds: torch.Tensor, numel: int):
    return torch.stack([first_inds, first_inds.diff(append=first_inds.new_tensor([numel]))], 1)


@torch.no_grad()
def get_pack_infos_from_boundary(boundary: torch.Tensor):
    return get_pack_infos_from_first(boundary.nonzero().long()[..., 0], boundary.numel())


@torch.no_grad()
def get_pack_infos_from_n(n_per_pack: torch.Tensor):
    return torch.stack([n_per_pack.cumsum(0) - n_per_pack, n_per_pack], 1)


@torch.no_grad()
def get_pack_infos_from_batch(n_batches: int, batch_data_size: int, device=None):
    return torch.stack([torch.arange(0, n_batches * batch_data_size, batch_data_size, device=device, dtype=torch.long),
                        torch.full([n_batches], batch_data_size, device=device, dtype=torch.long)], 1)
