"""Drop-in replacement for ``nr3d_lib.bindings._pack_ops`` (reference: csrc/pack_ops/pack_ops.cpp:20-58).

Implemented ops (hot path): interleave_arange, interleave_linstep,
interleave_sample_step_wrt_depth_clamped, packed_{add,sub,mul,div,gt,geq,lt,leq,eq,neq}, packed_sum, packed_diff,
packed_backward_diff, packed_cumsum, packed_cumprod, packed_alpha_to_vw_forward/backward, mark_pack_boundaries_cuda.
Hierarchical-sampling ops: packed_searchsorted, packed_searchsorted_packed_vals,
packed_invert_cdf, try_merge_two_packs_sorted_aligned, packed_sort_qsort / packed_sort_thrust, packed_matmul.
Also: interleave_sample_step_wrt_depth_in_packed_segments, interleave_sample_step_wrt_depth_clamp_deprecated and
octree_mark_consecutive_segments -- every name the reference module exports is implemented.
"""
import ctypes
from typing import Optional, Tuple

import torch

from .. import _lib

# exclusive packed_cumprod: documented semantics by default; set True to reproduce the reference CUDA output
# (all zeros).
EXCLUSIVE_CUMPROD_BUG_COMPAT = False


def _check_feats(fn, feats, pack_infos, check_size=False):
    if feats.dim() not in (1, 2):
        raise RuntimeError(f"{fn}: Expected 1 to 2 dimensions, but got {feats.dim()} for argument 'feats'")
    _check_pack_infos(fn, pack_infos)
    if not feats.is_contiguous():
        raise RuntimeError(f"{fn}: Expected contiguous tensor for argument 'feats'")
    dev = _lib.require_cuda(feats, pack_infos, who=fn)
    C = 1 if feats.dim() == 1 else feats.shape[1]
    return dev, pack_infos.shape[0], C


def _check_pack_infos(fn, pack_infos):
    if pack_infos.dim() != 2 or pack_infos.shape[1] != 2:
        raise RuntimeError(f"{fn}: Expected `pack_infos` of shape [num_packs, 2], got {tuple(pack_infos.shape)}")
    if pack_infos.dtype != torch.int64:
        raise RuntimeError(f"{fn}: Expected `pack_infos` to have scalar type Long, got {pack_infos.dtype}")
    if not pack_infos.is_contiguous():
        raise RuntimeError(f"{fn}: Expected contiguous tensor for argument 'pack_infos'")


def _opt_like(fn, name, t, feats, P, C):
    if t is None:
        return None
    if t.dtype != feats.dtype or not t.is_contiguous() or t.device != feats.device:
        raise RuntimeError(f"{fn}: `{name}` must be a contiguous tensor with the dtype/device of `feats`")
    want = (P,) if feats.dim() == 1 else (P, C)
    if tuple(t.shape) != want:
        raise RuntimeError(f"{fn}: `{name}` has shape {tuple(t.shape)}, expected {want}")
    return t


def _total_from_device(total: torch.Tensor) -> int:
    return int(total.item())  # the single host read needed to size the outputs


def packed_sum(feats: torch.Tensor, pack_infos: torch.Tensor) -> torch.Tensor:
    """== packed_sum (pack_ops_cuda.cu:826-861)."""
    dev, P, C = _check_feats("packed_sum", feats, pack_infos)
    with torch.cuda.device(dev):
        out = torch.empty([P] if feats.dim() == 1 else [P, C], dtype=feats.dtype, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_sum(_lib.dtype_code(feats.dtype), P, C, feats.shape[0], feats.data_ptr(), pack_infos.data_ptr(),
                                                out.data_ptr(), _lib.stream_of(dev)))
    return out


def _scan(fn_name, cfun, feats, pack_infos, exclusive, reverse, extra=()):
    dev, P, C = _check_feats(fn_name, feats, pack_infos)
    with torch.cuda.device(dev):
        out = torch.zeros_like(feats)
        _lib.check(cfun(_lib.dtype_code(feats.dtype), P, C, feats.shape[0], feats.data_ptr(), pack_infos.data_ptr(), int(bool(exclusive)),
                        int(bool(reverse)), *extra, out.data_ptr(), _lib.stream_of(dev)))
    return out


def packed_cumsum(feats: torch.Tensor, pack_infos: torch.Tensor, exclusive: bool, reverse: bool) -> torch.Tensor:
    """== packed_cumsum (pack_ops_cuda.cu:1048-1095)."""
    return _scan("packed_cumsum", _lib.get_lib().nr3d_pack_cumsum, feats, pack_infos, exclusive, reverse)


def packed_cumprod(feats: torch.Tensor, pack_infos: torch.Tensor, exclusive: bool, reverse: bool) -> torch.Tensor:
    """== packed_cumprod (pack_ops_cuda.cu:931-978); exclusive follows the documented semantics (see module flag)."""
    return _scan("packed_cumprod", _lib.get_lib().nr3d_pack_cumprod, feats, pack_infos, exclusive, reverse,
                 extra=(int(EXCLUSIVE_CUMPROD_BUG_COMPAT),))


def _diff(fn_name, cfun, feats, pack_infos, edge, fill, edge_name, fill_name):
    dev, P, C = _check_feats(fn_name, feats, pack_infos)
    if edge is not None and fill is not None:
        raise RuntimeError("You should only specify AT MOST one of [appends, prepends, last_fill, first_fill]")
    edge = _opt_like(fn_name, edge_name, edge, feats, P, C)
    fill = _opt_like(fn_name, fill_name, fill, feats, P, C)
    with torch.cuda.device(dev):
        out = torch.zeros_like(feats)
        _lib.check(cfun(_lib.dtype_code(feats.dtype), P, C, feats.data_ptr(), pack_infos.data_ptr(), _lib.ptr(edge), _lib.ptr(fill),
                        out.data_ptr(), _lib.stream_of(dev)))
    return out


def packed_diff(feats, pack_infos, pack_appends: Optional[torch.Tensor] = None, pack_last_fill: Optional[torch.Tensor] = None):
    """== packed_diff (pack_ops_cuda.cu:1186-1259)."""
    return _diff("packed_diff", _lib.get_lib().nr3d_pack_diff, feats, pack_infos, pack_appends, pack_last_fill,
                 "pack_appends", "pack_last_fill")


def packed_backward_diff(feats, pack_infos, pack_prepends: Optional[torch.Tensor] = None, pack_first_fill: Optional[torch.Tensor] = None):
    """== packed_backward_diff (pack_ops_cuda.cu:1261-1334)."""
    return _diff("packed_backward_diff", _lib.get_lib().nr3d_pack_backward_diff, feats, pack_infos, pack_prepends, pack_first_fill,
                 "pack_prepends", "pack_first_fill")


def _binary(op, name, feats, other, pack_infos):
    dev, P, C = _check_feats(name, feats, pack_infos)
    if other.dim() != feats.dim() or other.shape[0] != P or (feats.dim() == 2 and other.shape[1] != C):
        raise RuntimeError(f"{name}: `other` has shape {tuple(other.shape)}, expected [{P}{', %d' % C if feats.dim() == 2 else ''}]")
    if other.dtype != feats.dtype or not other.is_contiguous():
        raise RuntimeError(f"{name}: `other` must be contiguous and share the dtype of `feats`")
    _lib.require_cuda(feats, other, who=name)
    with torch.cuda.device(dev):
        out = torch.zeros(feats.shape, dtype=feats.dtype if op < 4 else torch.bool, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_binary(op, _lib.dtype_code(feats.dtype), P, C, feats.data_ptr(), other.data_ptr(),
                                                   pack_infos.data_ptr(), out.data_ptr(), _lib.stream_of(dev)))
    return out


def packed_add(feats, other, pack_infos): return _binary(0, "packed_add", feats, other, pack_infos)
def packed_sub(feats, other, pack_infos): return _binary(1, "packed_sub", feats, other, pack_infos)
def packed_mul(feats, other, pack_infos): return _binary(2, "packed_mul", feats, other, pack_infos)
def packed_div(feats, other, pack_infos): return _binary(3, "packed_div", feats, other, pack_infos)
def packed_gt(feats, other, pack_infos): return _binary(5, "packed_gt", feats, other, pack_infos)
def packed_geq(feats, other, pack_infos): return _binary(6, "packed_geq", feats, other, pack_infos)
def packed_lt(feats, other, pack_infos): return _binary(7, "packed_lt", feats, other, pack_infos)
def packed_leq(feats, other, pack_infos): return _binary(8, "packed_leq", feats, other, pack_infos)
def packed_eq(feats, other, pack_infos): return _binary(9, "packed_eq", feats, other, pack_infos)
def packed_neq(feats, other, pack_infos): return _binary(10, "packed_neq", feats, other, pack_infos)


def _pack_infos_from_counts(counts: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """counts int64 [P] -> (pack_infos int64 [P,2], total int64 [1]) entirely on device."""
    dev = counts.device
    P = counts.shape[0]
    lib = _lib.get_lib()
    pack_infos = torch.empty([P, 2], dtype=torch.int64, device=dev)
    total = torch.zeros([1], dtype=torch.int64, device=dev)
    nbytes = ctypes.c_uint64(0)
    _lib.check(lib.nr3d_pack_infos_from_counts(P, None, None, None, None, ctypes.byref(nbytes), None))
    ws = torch.empty([max(8, nbytes.value)], dtype=torch.uint8, device=dev)
    nbytes = ctypes.c_uint64(ws.numel())
    _lib.check(lib.nr3d_pack_infos_from_counts(P, counts.data_ptr(), pack_infos.data_ptr(), total.data_ptr(), ws.data_ptr(),
                                               ctypes.byref(nbytes), _lib.stream_of(dev)))
    return pack_infos, total


def packed_alpha_to_vw_forward(alphas: torch.Tensor, pack_infos: torch.Tensor, early_stop_eps: float, alpha_thre: float,
                               compression: bool):
    """== packed_alpha_to_vw_forward (pack_ops_cuda.cu:1850-1911) -> (weights, compact_pack_info, compact_selector)."""
    fn = "packed_alpha_to_vw_forward"
    if alphas.dim() != 1:
        raise RuntimeError(f"{fn}: Expected 1-dimensional tensor for argument 'alphas'")
    if alphas.dtype not in (torch.float16, torch.float32, torch.float64):
        raise RuntimeError(f"{fn}: Expected 'alphas' to have scalar type Half, Float or Double")
    dev, P, _ = _check_feats(fn, alphas, pack_infos)
    lib = _lib.get_lib()
    weights = compact_pack_info = compact_selector = None
    with torch.cuda.device(dev):
        st = _lib.stream_of(dev)
        if compression:
            num_steps = torch.zeros([P], dtype=torch.int64, device=dev)
            compact_selector = torch.zeros([alphas.shape[0]], dtype=torch.bool, device=dev)
            _lib.check(lib.nr3d_pack_alpha_to_vw_fwd(_lib.dtype_code(alphas.dtype), P, alphas.shape[0], alphas.data_ptr(), pack_infos.data_ptr(),
                                                     float(early_stop_eps), float(alpha_thre), None, num_steps.data_ptr(),
                                                     compact_selector.data_ptr(), st))
            # the reference builds this with cumsum(at::kInt) + stack => an int32 [P,2] tensor (pack_ops_cuda.cu:1874-1875)
            compact_pack_info = _pack_infos_from_counts(num_steps)[0].to(torch.int32)
        else:
            weights = torch.zeros_like(alphas)
            _lib.check(lib.nr3d_pack_alpha_to_vw_fwd(_lib.dtype_code(alphas.dtype), P, alphas.shape[0], alphas.data_ptr(), pack_infos.data_ptr(),
                                                     float(early_stop_eps), float(alpha_thre), weights.data_ptr(), None, None, st))
    return weights, compact_pack_info, compact_selector


def packed_alpha_to_vw_backward(weights: torch.Tensor, grad_weights: torch.Tensor, alphas: torch.Tensor, pack_infos: torch.Tensor,
                                early_stop_eps: float, alpha_thre: float) -> torch.Tensor:
    """== packed_alpha_to_vw_backward (pack_ops_cuda.cu:1914-1958)."""
    fn = "packed_alpha_to_vw_backward"
    for name, t in (("weights", weights), ("grad_weights", grad_weights), ("alphas", alphas)):
        if t.dim() != 1 or not t.is_contiguous():
            raise RuntimeError(f"{fn}: Expected contiguous 1-dimensional tensor for argument '{name}'")
    if not (weights.dtype == grad_weights.dtype == alphas.dtype) or weights.dtype not in (torch.float16, torch.float32, torch.float64):
        raise RuntimeError(f"{fn}: weights / grad_weights / alphas must share a Half, Float or Double dtype")
    if not (weights.shape == grad_weights.shape == alphas.shape):
        raise RuntimeError(f"{fn}: weights / grad_weights / alphas must have the same size")
    _check_pack_infos(fn, pack_infos)
    dev = _lib.require_cuda(weights, grad_weights, alphas, pack_infos, who=fn)
    P = pack_infos.shape[0]
    with torch.cuda.device(dev):
        grad_alphas = torch.zeros_like(alphas)
        _lib.check(_lib.get_lib().nr3d_pack_alpha_to_vw_bwd(
            _lib.dtype_code(weights.dtype), P, alphas.shape[0], weights.data_ptr(), grad_weights.data_ptr(), alphas.data_ptr(), pack_infos.data_ptr(),
            float(early_stop_eps), float(alpha_thre), grad_alphas.data_ptr(), _lib.stream_of(dev)))
    return grad_alphas


def interleave_arange(stop: torch.Tensor, return_idx: bool):
    """== interleave_arange (pack_ops_cuda.cu:85-118): concatenated arange(stop[p]) (+ pack ids)."""
    if stop.dim() != 1 or stop.dtype != torch.int64 or not stop.is_contiguous():
        raise RuntimeError("interleave_arange: `stop` must be a contiguous 1-D int64 tensor")
    dev = _lib.require_cuda(stop, who="interleave_arange")
    with torch.cuda.device(dev):
        pack_infos, total = _pack_infos_from_counts(stop)
        num = _total_from_device(total)
        out = torch.empty([num], dtype=torch.int64, device=dev)
        nidx = torch.empty([num], dtype=torch.int64, device=dev) if return_idx else None
        _lib.check(_lib.get_lib().nr3d_pack_interleave_linstep(_lib.I64, stop.shape[0], pack_infos.data_ptr(), None, None, 0.0, 1.0,
                                                               out.data_ptr(), _lib.ptr(nidx), _lib.stream_of(dev)))
    return out, nidx


def interleave_linstep(start: torch.Tensor, num_steps: torch.Tensor, step_size, return_idx: bool):
    """== interleave_linstep, three overloads (pack_ops_cuda.cu:120-218)."""
    if start.dim() != 1 or num_steps.dim() != 1 or start.shape != num_steps.shape:
        raise RuntimeError("interleave_linstep: `start` and `num_steps` must be 1-D tensors of the same size")
    if num_steps.dtype != torch.int64:
        raise RuntimeError("interleave_linstep: `num_steps` must have scalar type Long")
    steps_t = None
    step_s = 0.0
    if isinstance(step_size, torch.Tensor):
        if step_size.dim() != 1 or step_size.dtype != start.dtype or not step_size.is_contiguous():
            raise RuntimeError("interleave_linstep: tensor `step_size` must be 1-D, contiguous and share the dtype of `start`")
        steps_t = step_size
    else:
        step_s = float(step_size)
    start, num_steps = start.contiguous(), num_steps.contiguous()
    dev = _lib.require_cuda(start, num_steps, steps_t, who="interleave_linstep")
    with torch.cuda.device(dev):
        pack_infos, total = _pack_infos_from_counts(num_steps)
        num = _total_from_device(total)
        out = torch.empty([num], dtype=start.dtype, device=dev)
        nidx = torch.empty([num], dtype=torch.int64, device=dev) if return_idx else None
        _lib.check(_lib.get_lib().nr3d_pack_interleave_linstep(_lib.dtype_code(start.dtype), start.shape[0], pack_infos.data_ptr(),
                                                               start.data_ptr(), _lib.ptr(steps_t), 0.0, step_s, out.data_ptr(),
                                                               _lib.ptr(nidx), _lib.stream_of(dev)))
    return out, nidx


def interleave_sample_step_wrt_depth_clamped(near: torch.Tensor, far: torch.Tensor, max_steps: int, dt_gamma: float,
                                             min_step_size: float, max_step_size: float):
    """== interleave_sample_step_wrt_depth_clamped (pack_ops_cuda.cu:547-604) -> (t_samples, deltas, ridx, pack_infos)."""
    fn = "interleave_sample_step_wrt_depth_clamped"
    if near.dim() != 1 or far.dim() != 1 or near.shape != far.shape or near.dtype != far.dtype:
        raise RuntimeError(f"{fn}: `near` / `far` must be 1-D tensors of the same size and dtype")
    if not (near.is_contiguous() and far.is_contiguous()):
        raise RuntimeError(f"{fn}: Expected contiguous tensors")
    dev = _lib.require_cuda(near, far, who=fn)
    lib = _lib.get_lib()
    P = near.shape[0]
    code = _lib.dtype_code(near.dtype)
    with torch.cuda.device(dev):
        st = _lib.stream_of(dev)
        n_per_pack = torch.empty([P], dtype=torch.int64, device=dev)
        _lib.check(lib.nr3d_pack_sample_step_count(code, P, near.data_ptr(), far.data_ptr(), int(max_steps), float(dt_gamma),
                                                   float(min_step_size), float(max_step_size), n_per_pack.data_ptr(), st))
        pack_infos, total = _pack_infos_from_counts(n_per_pack)
        num = _total_from_device(total)
        t_samples = torch.empty([num], dtype=near.dtype, device=dev)
        deltas = torch.empty([num], dtype=near.dtype, device=dev)
        nidx = torch.empty([num], dtype=torch.int64, device=dev)
        _lib.check(lib.nr3d_pack_sample_step_fill(code, P, near.data_ptr(), pack_infos.data_ptr(), float(dt_gamma),
                                                  float(min_step_size), float(max_step_size), t_samples.data_ptr(), deltas.data_ptr(),
                                                  nidx.data_ptr(), st))
    return t_samples, deltas, nidx, pack_infos


def mark_pack_boundaries_cuda(pack_ids: torch.Tensor) -> torch.Tensor:
    """== mark_pack_boundaries_cuda (pack_ops_cuda.cu:2784-2805): int32 [S]."""
    if pack_ids.dim() != 1 or not pack_ids.is_contiguous():
        raise RuntimeError("mark_pack_boundaries_cuda: Expected contiguous 1-dimensional tensor")
    if pack_ids.dtype not in (torch.uint8, torch.int8, torch.int16, torch.int32, torch.int64):
        raise RuntimeError("mark_pack_boundaries_cuda: Expected an integral scalar type (Byte, Char, Short, Int, Long)")
    dev = _lib.require_cuda(pack_ids, who="mark_pack_boundaries_cuda")
    with torch.cuda.device(dev):
        out = torch.zeros([pack_ids.shape[0]], dtype=torch.int32, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_mark_boundaries(_lib.dtype_code(pack_ids.dtype), pack_ids.shape[0], pack_ids.data_ptr(),
                                                            out.data_ptr(), _lib.stream_of(dev)))
    return out


# ------------------------------------------------------------------------------------------------------------------
# hierarchical-sampling ops
# ------------------------------------------------------------------------------------------------------------------
def _check_1d_same(fn, *named):
    ref = named[0][1]
    for name, t in named:
        if t.dim() != 1 or not t.is_contiguous():
            raise RuntimeError(f"{fn}: Expected contiguous 1-dimensional tensor for argument '{name}'")
        if t.dtype != ref.dtype:
            raise RuntimeError(f"{fn}: '{name}' must have the dtype of '{named[0][0]}'")


def packed_searchsorted(bins: torch.Tensor, vals: torch.Tensor, pack_infos: torch.Tensor) -> torch.Tensor:
    """== packed_searchsorted (pack_ops_cuda.cu:1409-1455): vals [num_packs, num_to_search] searched in the sorted packs of `bins`."""
    fn = "packed_searchsorted"
    _check_1d_same(fn, ("bins", bins))
    _check_pack_infos(fn, pack_infos)
    if vals.dim() != 2 or not vals.is_contiguous() or vals.dtype != bins.dtype or vals.shape[0] != pack_infos.shape[0]:
        raise RuntimeError(f"{fn}: `vals` must be a contiguous [num_packs, num_to_search] tensor with the dtype of `bins`")
    dev = _lib.require_cuda(bins, vals, pack_infos, who=fn)
    with torch.cuda.device(dev):
        pidx = torch.full(vals.shape, -1, dtype=torch.int64, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_searchsorted(_lib.dtype_code(bins.dtype), pack_infos.shape[0], bins.data_ptr(), pack_infos.data_ptr(),
                                                         vals.data_ptr(), vals.shape[1], None, pidx.data_ptr(), _lib.stream_of(dev)))
    return pidx


def packed_searchsorted_packed_vals(bins: torch.Tensor, pack_infos: torch.Tensor, vals: torch.Tensor, val_pack_infos: torch.Tensor) -> torch.Tensor:
    """== packed_searchsorted_packed_vals (pack_ops_cuda.cu:1457-1503): packed queries, one query pack per bin pack."""
    fn = "packed_searchsorted_packed_vals"
    _check_1d_same(fn, ("bins", bins), ("vals", vals))
    _check_pack_infos(fn, pack_infos)
    _check_pack_infos(fn, val_pack_infos)
    if val_pack_infos.shape[0] != pack_infos.shape[0]:
        raise RuntimeError(f"{fn}: `val_pack_infos` must have one row per pack")
    dev = _lib.require_cuda(bins, vals, pack_infos, val_pack_infos, who=fn)
    with torch.cuda.device(dev):
        pidx = torch.full(vals.shape, -1, dtype=torch.int64, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_searchsorted(_lib.dtype_code(bins.dtype), pack_infos.shape[0], bins.data_ptr(), pack_infos.data_ptr(),
                                                         vals.data_ptr(), 0, val_pack_infos.data_ptr(), pidx.data_ptr(), _lib.stream_of(dev)))
    return pidx


def packed_invert_cdf(bins: torch.Tensor, cdfs: torch.Tensor, u_vals: torch.Tensor, pack_infos: torch.Tensor):
    """== packed_invert_cdf (pack_ops_cuda.cu:1683-1733) -> (samples, bin_idx), both [num_packs, num_to_sample]."""
    fn = "packed_invert_cdf"
    _check_1d_same(fn, ("bins", bins), ("cdfs", cdfs))
    _check_pack_infos(fn, pack_infos)
    if bins.shape != cdfs.shape:
        raise RuntimeError(f"{fn}: `bins` and `cdfs` must have the same size")
    if u_vals.dim() != 2 or not u_vals.is_contiguous() or u_vals.dtype != bins.dtype or u_vals.shape[0] != pack_infos.shape[0]:
        raise RuntimeError(f"{fn}: `u_vals` must be a contiguous [num_packs, num_to_sample] tensor with the dtype of `bins`")
    dev = _lib.require_cuda(bins, cdfs, u_vals, pack_infos, who=fn)
    with torch.cuda.device(dev):
        bin_idx = torch.full(u_vals.shape, -1, dtype=torch.int64, device=dev)
        samples = torch.zeros_like(u_vals)
        _lib.check(_lib.get_lib().nr3d_pack_invert_cdf(_lib.dtype_code(bins.dtype), pack_infos.shape[0], bins.data_ptr(), cdfs.data_ptr(),
                                                       pack_infos.data_ptr(), u_vals.data_ptr(), u_vals.shape[1], samples.data_ptr(),
                                                       bin_idx.data_ptr(), _lib.stream_of(dev)))
    return samples, bin_idx


def try_merge_two_packs_sorted_aligned(vals_a: torch.Tensor, pack_infos_a: torch.Tensor, vals_b: torch.Tensor, pack_infos_b: torch.Tensor,
                                       b_sorted: bool):
    """== try_merge_two_packs_sorted_aligned (pack_ops_cuda.cu:1573-1631) -> (pidx_a, pidx_b, merged pack_infos)."""
    fn = "try_merge_two_packs_sorted_aligned"
    _check_1d_same(fn, ("vals_a", vals_a), ("vals_b", vals_b))
    _check_pack_infos(fn, pack_infos_a)
    _check_pack_infos(fn, pack_infos_b)
    if pack_infos_a.shape[0] != pack_infos_b.shape[0]:
        raise RuntimeError(f"{fn}: the two packs must be aligned (same number of packs)")
    dev = _lib.require_cuda(vals_a, vals_b, pack_infos_a, pack_infos_b, who=fn)
    with torch.cuda.device(dev):
        n = (pack_infos_a[:, 1] + pack_infos_b[:, 1]).contiguous()
        pack_infos, _ = _pack_infos_from_counts(n)
        pidx_a = torch.zeros([vals_a.shape[0]], dtype=torch.int64, device=dev)
        pidx_b = torch.zeros([vals_b.shape[0]], dtype=torch.int64, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_merge_sorted_aligned(
            _lib.dtype_code(vals_a.dtype), pack_infos_a.shape[0], vals_a.data_ptr(), pack_infos_a.data_ptr(), vals_b.data_ptr(),
            pack_infos_b.data_ptr(), pack_infos.data_ptr(), pidx_a.data_ptr(), pidx_b.data_ptr(), _lib.stream_of(dev)))
    return pidx_a, pidx_b, pack_infos


def packed_sort_qsort(vals: torch.Tensor, pack_infos: torch.Tensor, return_idx: bool):
    """== packed_sort_qsort (pack_ops_cuda.cu:2709-2763): ascending sort of every pack IN PLACE; returns the permuted global
    indices when `return_idx` (equal keys may come out in a different order than the reference's quicksort)."""
    fn = "packed_sort_qsort"
    _check_1d_same(fn, ("vals", vals))
    _check_pack_infos(fn, pack_infos)
    dev = _lib.require_cuda(vals, pack_infos, who=fn)
    with torch.cuda.device(dev):
        idx = torch.arange(vals.shape[0], dtype=torch.int64, device=dev) if return_idx else None
        _lib.check(_lib.get_lib().nr3d_pack_sort(_lib.dtype_code(vals.dtype), pack_infos.shape[0], vals.data_ptr(), pack_infos.data_ptr(),
                                                 _lib.ptr(idx), _lib.stream_of(dev)))
    return idx


packed_sort_thrust = packed_sort_qsort   # same contract (pack_ops_cuda.cu:2556-2621)


def packed_matmul(feats: torch.Tensor, other: torch.Tensor, pack_infos: torch.Tensor) -> torch.Tensor:
    """== packed_matmul (pack_ops_cuda.cu:2251-2540, Matmul branch): out = other[pack] @ feat, feats [S, C], other [P, C_out, C]."""
    fn = "packed_matmul"
    _check_pack_infos(fn, pack_infos)
    if feats.dim() != 2 or other.dim() != 3 or other.shape[2] != feats.shape[1] or other.shape[0] != pack_infos.shape[0]:
        raise RuntimeError(f"{fn}: expected feats [num_feats, C] and other [num_packs, C_out, C]")
    if not (feats.is_contiguous() and other.is_contiguous()) or feats.dtype != other.dtype:
        raise RuntimeError(f"{fn}: feats / other must be contiguous and share a dtype")
    dev = _lib.require_cuda(feats, other, pack_infos, who=fn)
    with torch.cuda.device(dev):
        out = torch.zeros([feats.shape[0], other.shape[1]], dtype=feats.dtype, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_matmul(_lib.dtype_code(feats.dtype), pack_infos.shape[0], feats.shape[1], other.shape[1],
                                                   feats.data_ptr(), other.data_ptr(), pack_infos.data_ptr(), out.data_ptr(), _lib.stream_of(dev)))
    return out


def interleave_sample_step_wrt_depth_clamp_deprecated(near: torch.Tensor, far: torch.Tensor, max_steps: int, dt_gamma: float,
                                                      min_step_size: float, max_step_size: float):
    """== interleave_sample_step_wrt_depth_clamp_deprecated (pack_ops_cuda.cu:226-396): the buffered two-storage variant of the
    depth sampler; its compacted outputs equal `interleave_sample_step_wrt_depth_clamped`'s, so both share one implementation."""
    return interleave_sample_step_wrt_depth_clamped(near, far, max_steps, dt_gamma, min_step_size, max_step_size)


def interleave_sample_step_wrt_depth_in_packed_segments(near: torch.Tensor, far: torch.Tensor, entry: torch.Tensor, exit: torch.Tensor,
                                                        seg_pack_infos: torch.Tensor, max_steps: int, dt_gamma: float, min_step_size: float,
                                                        max_step_size: float):
    """== interleave_sample_step_wrt_depth_in_packed_segments (pack_ops_cuda.cu:723-795)
    -> (t_samples, deltas, sidx, nidx, pack_infos)."""
    fn = "interleave_sample_step_wrt_depth_in_packed_segments"
    for name, t_ in (("near", near), ("far", far), ("entry", entry), ("exit", exit)):
        if t_.dim() != 1:
            raise RuntimeError(f"{fn}: Expected 1-dimensional tensor for argument '{name}'")
        if not t_.is_contiguous():
            raise RuntimeError(f"{fn}: Expected contiguous tensor for argument '{name}'")
        if t_.dtype != near.dtype:
            raise RuntimeError(f"{fn}: Expected `{name}` to have the same scalar type as `near` ({near.dtype}), got {t_.dtype}")
    if near.dtype not in (torch.float32, torch.float64):
        raise RuntimeError(f"{fn}: supported scalar types are Float and Double (Half: not built), got {near.dtype}")
    _check_pack_infos(fn, seg_pack_infos)
    if near.shape != far.shape or entry.shape != exit.shape or seg_pack_infos.shape[0] != near.shape[0]:
        raise RuntimeError(f"{fn}: size mismatch between near/far [{near.shape[0]}], entry/exit [{entry.shape[0]}] and seg_pack_infos")
    dev = _lib.require_cuda(near, far, entry, exit, seg_pack_infos, who=fn)
    lib = _lib.get_lib()
    P = near.shape[0]
    code = _lib.dtype_code(near.dtype)
    with torch.cuda.device(dev):
        st = _lib.stream_of(dev)
        n_per_pack = torch.empty([P], dtype=torch.int64, device=dev)
        _lib.check(lib.nr3d_pack_seg_sample_count(code, P, near.data_ptr(), far.data_ptr(), _lib.ptr(entry), _lib.ptr(exit),
                                                  seg_pack_infos.data_ptr(), int(max_steps), float(dt_gamma), float(min_step_size),
                                                  float(max_step_size), n_per_pack.data_ptr(), st))
        pack_infos, total = _pack_infos_from_counts(n_per_pack)
        num = _total_from_device(total)
        t_samples = torch.empty([num], dtype=near.dtype, device=dev)
        deltas = torch.empty([num], dtype=near.dtype, device=dev)
        nidx = torch.empty([num], dtype=torch.int64, device=dev)
        sidx = torch.empty([num], dtype=torch.int64, device=dev)
        _lib.check(lib.nr3d_pack_seg_sample_fill(code, P, near.data_ptr(), far.data_ptr(), _lib.ptr(entry), _lib.ptr(exit),
                                                 seg_pack_infos.data_ptr(), pack_infos.data_ptr(), float(dt_gamma), float(min_step_size),
                                                 float(max_step_size), _lib.ptr(t_samples), _lib.ptr(deltas), _lib.ptr(nidx), _lib.ptr(sidx), st))
    return t_samples, deltas, sidx, nidx, pack_infos


# octree_mark_consecutive_segments: the reference walks `point_indices` from element 0 for every pack (pack_ops_cuda.cu:2832-2833),
# which is only right for the first pack.  Default reproduces it; set True to walk each pack's own nuggets.
OCTREE_SEGMENTS_OFFSET_FIX = False


def octree_mark_consecutive_segments(pidx: torch.Tensor, pack_infos: torch.Tensor, point_hierarchies: torch.Tensor):
    """== octree_mark_consecutive_segments (pack_ops_cuda.cu:2843-2885) -> (mark_start, mark_end), bool [num_nuggets]."""
    fn = "octree_mark_consecutive_segments"
    if pidx.dim() != 1 or pidx.dtype != torch.int32 or not pidx.is_contiguous():
        raise RuntimeError(f"{fn}: Expected contiguous 1-dimensional Int tensor for argument 'pidx'")
    _check_pack_infos(fn, pack_infos)
    if point_hierarchies.dtype != torch.int16 or not point_hierarchies.is_contiguous():
        raise RuntimeError(f"{fn}: Expected contiguous Short tensor for argument 'point_hierarchies'")
    dev = _lib.require_cuda(pidx, pack_infos, point_hierarchies, who=fn)
    with torch.cuda.device(dev):
        mark_start = torch.zeros([pidx.shape[0]], dtype=torch.bool, device=dev)
        mark_end = torch.zeros([pidx.shape[0]], dtype=torch.bool, device=dev)
        _lib.check(_lib.get_lib().nr3d_pack_mark_consecutive_segments(pack_infos.shape[0], pack_infos.data_ptr(), _lib.ptr(pidx),
                                                                      _lib.ptr(point_hierarchies), int(OCTREE_SEGMENTS_OFFSET_FIX),
                                                                      _lib.ptr(mark_start), _lib.ptr(mark_end), _lib.stream_of(dev)))
    return mark_start, mark_end
