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
    """Stand-in for the density head of the decoder: softplus of the (scaled) feature sum (torch composition; the
    pipeline itself uses the one-pass `density_alpha` below, this one is what the tests compare it with)."""
    return F.softplus(features.sum(-1) * gain)


class _DensityAlpha(torch.autograd.Function):
    """sigma = softplus(gain * sum_c h), alpha = 1 - exp(-sigma * deltas) in ONE pass over the [S, C] features each way
    (csrc/pipeline_ops.cu) instead of sum / mul / softplus / mul / exp / rsub and their six backward kernels."""

    @staticmethod
    def forward(ctx, h, deltas, gain):
        from . import _lib
        dev = _lib.require_cuda(h, deltas, who="density_alpha")
        if h.dtype != torch.float32 or h.dim() != 2:
            raise RuntimeError("density_alpha: features must be float32 [S, C]")
        if h.stride(1) != 1:
            h = h.contiguous()
        deltas = deltas.contiguous()
        S, C = h.shape
        with torch.cuda.device(dev):
            sigma = torch.empty(S, dtype=torch.float32, device=dev)
            alpha = torch.empty(S, dtype=torch.float32, device=dev)
            _lib.check(_lib.get_lib().nr3d_density_alpha_fwd(S, C, h.data_ptr(), h.stride(0), deltas.data_ptr(), float(gain), sigma.data_ptr(),
                                                             alpha.data_ptr(), _lib.stream_of(dev)))
        ctx.save_for_backward(sigma, alpha, deltas)
        ctx.gain, ctx.C = float(gain), C
        ctx.mark_non_differentiable(sigma)
        return alpha, sigma

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_alpha, _d_sigma):
        from . import _lib
        sigma, alpha, deltas = ctx.saved_tensors
        if d_alpha is None:
            return None, None, None
        d_alpha = d_alpha.contiguous()
        S, dev = sigma.shape[0], sigma.device
        with torch.cuda.device(dev):
            d_h = torch.empty(S, ctx.C, dtype=torch.float32, device=dev)
            _lib.check(_lib.get_lib().nr3d_density_alpha_bwd(S, ctx.C, d_alpha.data_ptr(), None, sigma.data_ptr(), alpha.data_ptr(),
                                                             deltas.data_ptr(), ctx.gain, d_h.data_ptr(), _lib.stream_of(dev)))
        return d_h, None, None


def density_alpha(features: torch.Tensor, deltas: torch.Tensor, gain: float = 20.0):
    """(alpha, sigma) of the stand-in density head; differentiable w.r.t. `features` through alpha."""
    return _DensityAlpha.apply(features, deltas, gain)


class _EncodeDensityAlpha(torch.autograd.Function):
    """encode + density head in ONE kernel each way (csrc/lotd_fast.cu, HEAD = true): alpha, sigma = head(LoTD(x)) without the [S, n_enc]
    features or their gradient ever reaching HBM.  Same values as `density_alpha(encoder(x, params), deltas, gain)` up to the summation
    order of the features (tests/test_pipeline_gpu.py).  Differentiable w.r.t. `params` through alpha."""

    @staticmethod
    def forward(ctx, meta, x01, params, deltas, gain, max_level, from_ndc=False):
        import ctypes
        from . import _lib
        from .bindings import _lotd
        dev = _lib.require_cuda(x01, params, deltas, who="encode_density_alpha")
        # LoTDFunction clamps the points (reference lotd.py:211) and LoTDEncoding maps [-1, 1] -> [0, 1] (lotd_encoding.py:162): both happen
        # inside the point sort here, so no pass over [S, 3] is spent on them
        x = x01.detach().contiguous()
        cmap = (0.5, 0.5, True) if from_ndc else (1.0, 0.0, True)
        deltas = deltas.detach().float().contiguous().flatten()
        N = x.shape[0]
        ml = meta.n_levels if max_level is None else int(max_level)
        with torch.cuda.device(dev):
            sigma = torch.empty(N, dtype=torch.float32, device=dev)
            alpha = torch.empty(N, dtype=torch.float32, device=dev)
            if N:
                xs, scenes = _lotd._sorted_points(x, expect_new=True, coord_map=cmap)
                _lib.check(_lib.get_lib().nr3d_lotd_density_head_fwd_sorted(
                    ctypes.byref(meta._c), _lib.dtype_code(params.dtype), N, xs.data_ptr(), _lib.ptr(scenes), 1, params.data_ptr(), ml,
                    deltas.data_ptr(), float(gain), sigma.data_ptr(), alpha.data_ptr(), _lib.stream_of(dev)))
        ctx.save_for_backward(x, sigma, alpha, deltas)
        ctx.meta, ctx.gain, ctx.ml, ctx.pshape, ctx.pdtype, ctx.cmap = meta, float(gain), ml, params.shape, params.dtype, cmap
        ctx.records = _lotd._records_token(x) if N else None
        ctx.mark_non_differentiable(sigma)
        return alpha, sigma

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, d_alpha, _d_sigma):
        import ctypes
        from . import _lib
        from .bindings import _lotd
        x, sigma, alpha, deltas = ctx.saved_tensors
        if d_alpha is None or not ctx.needs_input_grad[2]:
            return None, None, None, None, None, None, None
        d_alpha = d_alpha.float().contiguous()
        dev, N = x.device, x.shape[0]
        with torch.cuda.device(dev):
            g = torch.zeros(ctx.pshape, dtype=ctx.pdtype, device=dev)
            if N:
                # the forward's records: for certain when no other sort ran on this stream since (x is saved for backward, so autograd has
                # checked that nobody wrote to it) -- otherwise the fingerprint pass decides and re-sorts if it must
                cur = _lotd._records_if_untouched(ctx.records, dev)
                xs, scenes = cur if cur is not None else _lotd._sorted_points(x, coord_map=ctx.cmap)
                _lib.check(_lib.get_lib().nr3d_lotd_density_head_bwd_sorted(
                    ctypes.byref(ctx.meta._c), _lib.dtype_code(ctx.pdtype), N, xs.data_ptr(), _lib.ptr(scenes), 1, d_alpha.data_ptr(), sigma.data_ptr(),
                    alpha.data_ptr(), deltas.data_ptr(), ctx.gain, ctx.ml, g.data_ptr(), _lib.stream_of(dev)))
        return None, None, g, None, None, None, None


def encode_density_alpha(encoder: LoTD, x01: torch.Tensor, params: torch.Tensor, deltas: torch.Tensor, gain: float = 20.0, max_level=None,
                         from_ndc: bool = False):
    """(alpha, sigma) of the stand-in density head applied to the LoTD features of `x01`; one fused kernel each way when the encoder is
    eligible for the cell-sorted fast path (Dense/Hash, D = 3, single scene), the two-kernel composition otherwise.
    `from_ndc`: the points are given in [-1, 1]^3 (ray samples) and mapped to the unit cube on the fly."""
    from .bindings import _lotd
    meta = encoder.meta
    p = params.to(encoder.dtype)
    if _lotd._sorted_eligible(meta, x01, p, None, None, 0) and x01.dim() == 2 and p.shape[0] == meta.n_params:
        return _EncodeDensityAlpha.apply(meta, x01, p, deltas, gain, max_level, from_ndc)
    if from_ndc:
        x01 = x01 * 0.5 + 0.5
    return density_alpha(encoder(x01, params, max_level=max_level).float(), deltas, gain)


class _PackedWeightedSums(torch.autograd.Function):
    """(acc, depth) = (packed_sum(w), packed_sum(w * t)) in one pass each way (csrc/pack_staged.cu); bit-identical to the composition."""

    @staticmethod
    def forward(ctx, w, t, pack_infos):
        from . import _lib
        dev = _lib.require_cuda(w, t, pack_infos, who="packed_weighted_sums")
        if w.dtype != torch.float32 or t.dtype != torch.float32 or w.dim() != 1 or w.shape != t.shape:
            raise RuntimeError("packed_weighted_sums: weights / depths must be float32 [S] tensors of the same size")
        if pack_infos.dim() != 2 or pack_infos.shape[1] != 2 or pack_infos.dtype != torch.int64:
            raise RuntimeError("packed_weighted_sums: pack_infos must be int64 [P, 2]")
        w, t, pack_infos = w.contiguous(), t.detach().contiguous(), pack_infos.contiguous()
        P, S = pack_infos.shape[0], w.shape[0]
        with torch.cuda.device(dev):
            out = torch.empty([2, P], dtype=torch.float32, device=dev)
            _lib.check(_lib.get_lib().nr3d_pack_weighted_sums_fwd(P, S, w.data_ptr(), t.data_ptr(), pack_infos.data_ptr(), out[0].data_ptr(),
                                                                  out[1].data_ptr(), _lib.stream_of(dev)))
        ctx.save_for_backward(t, pack_infos)
        return out[0], out[1]

    @staticmethod
    @torch.autograd.function.once_differentiable
    def backward(ctx, g_acc, g_depth):
        from . import _lib
        t, pack_infos = ctx.saved_tensors
        if g_acc is None and g_depth is None:
            return None, None, None
        dev, P, S = t.device, pack_infos.shape[0], t.shape[0]
        ga = None if g_acc is None else g_acc.float().contiguous()
        gd = None if g_depth is None else g_depth.float().contiguous()
        with torch.cuda.device(dev):
            gw = torch.zeros([S], dtype=torch.float32, device=dev)
            _lib.check(_lib.get_lib().nr3d_pack_weighted_sums_bwd(P, S, t.data_ptr(), pack_infos.data_ptr(), _lib.ptr(ga), _lib.ptr(gd), gw.data_ptr(),
                                                                  _lib.stream_of(dev)))
        return gw, None, None


def packed_weighted_sums(weights: torch.Tensor, depths: torch.Tensor, pack_infos: torch.Tensor):
    """(accumulated opacity, expected depth) of every pack: == (packed_sum(w, pi), packed_sum(w * depths, pi)), differentiable w.r.t. `weights`."""
    return _PackedWeightedSums.apply(weights, depths, pack_infos)


def march_encode_composite(encoder: LoTD, params: torch.Tensor, occ_grid: torch.Tensor, rays_o: torch.Tensor, rays_d: torch.Tensor,
                           near: torch.Tensor, far: torch.Tensor, *, step_size: float = 0.01, max_steps: int = 512, gain: float = 20.0,
                           early_stop_eps: float = 1e-4, alpha_thre: float = 0.0, fuse_head: bool = True) -> RenderOut:
    """One differentiable render step for rays already intersected with the [-1,1]^3 box (near / far given).  `fuse_head=False` keeps the
    [S, n_enc] features in HBM between the encoder and the density head (the composition the reference's callers run)."""
    ret = occgrid_raymarch(occ_grid, rays_o, rays_d, near, far, step_size=step_size, max_steps=max_steps)
    if ret.num_hit_rays == 0:
        return RenderOut(ret, None, None, None)
    if fuse_head:
        # [-1,1] -> [0,1] (LoTDEncoding.forward, lotd_encoding.py:162) and the clamp happen inside the point sort
        alpha, _sigma = encode_density_alpha(encoder, ret.samples, params, ret.deltas, gain, from_ndc=True)
    else:
        x01 = ret.samples * 0.5 + 0.5
        h = encoder(x01, params)                       # [S, n_enc]
        alpha, _sigma = density_alpha(h.float(), ret.deltas, gain)   # softplus head + (1 - exp(-sigma * delta)), nerf_ray_query.py:182
    w = packed_alpha_to_vw(alpha, ret.pack_infos, early_stop_eps, alpha_thre)
    if fuse_head:
        acc, depth = packed_weighted_sums(w, ret.depth_samples.reshape(-1), ret.pack_infos)     # one pass each way, same values as below
    else:
        depth = packed_sum(w * ret.depth_samples, ret.pack_infos)
        acc = packed_sum(w, ret.pack_infos)
    return RenderOut(ret, depth, acc, w)


class HostFedLoTDStep:
    """LoTD fwd + bwd(dL/dparam) steps whose inputs and results live in HOST memory.

    Every step copies its points host->device and its parameter gradients device->host, like a trainer whose sampler and
    optimizer sit on the host.  Three CUDA streams and double buffers keep both copy engines and the SMs busy at once: while
    the kernels of step i run, the points of step i+1 are already crossing PCIe and the gradients of step i-1 are on their way
    back.  No copy is skipped or shortened -- they are only taken off the critical path.

        pipe = HostFedLoTDStep(meta, params, n_points, device, grad_of_y=lambda y: y * 1e-4)
        pipe.prefetch(x_host[0])
        for i in range(steps):
            pipe.step(x_host[i + 1] if i + 1 < steps else None, grad_host[i % 2])
        pipe.drain()                    # the compute stream waits for the last device->host copy

    `x_host` / `grad_host` must be pinned (allocate them after `dist.bind_to_gpu_numa`, so that they live on the GPU's NUMA node).
    `grad_of_y` stands for whatever turns the step's features into dL/dy on the device (decoder + loss); with world > 1 the gradients are
    summed with ONE reduce-scatter (`dist.GradReducer(mode="scatter")`): every rank ends with its own contiguous 1/world slice of the
    summed table and writes exactly that slice of `grad_host`, so the job returns the result once and NVLink carries half the bytes of an
    all-reduce.  Consecutive steps must bring different points (a trainer always does): the forward sorts them, the backward reuses the
    records after the on-device fingerprint check (bindings/_lotd.py).
    """

    def __init__(self, meta, params: torch.Tensor, n_points: int, device, grad_of_y, world: int = 1, rank: int = 0):
        from .bindings import _lotd
        self._lotd, self.meta, self.params, self.world, self.rank, self.grad_of_y = _lotd, meta, params, world, rank, grad_of_y
        self.dev = torch.device(device)
        self.compute = torch.cuda.current_stream(self.dev)
        self.h2d, self.d2h = torch.cuda.Stream(self.dev), torch.cuda.Stream(self.dev)
        self.xbuf = [torch.empty(n_points, 3, dtype=torch.float32, device=self.dev) for _ in range(2)]
        self.x_ready = [torch.cuda.Event(), torch.cuda.Event()]     # H2D of buffer b finished
        self.x_free = [torch.cuda.Event(), torch.cuda.Event()]      # kernels reading buffer b finished
        self.inflight = []                                          # (event, grad tensor) of pending device->host copies
        self.slot = 0
        from . import dist as ndist
        self.reducer = ndist.GradReducer(meta, world, self.dev, mode="scatter")
        for e in self.x_free:
            e.record(self.compute)

    def prefetch(self, x_host: torch.Tensor):
        b = self.slot
        self.h2d.wait_event(self.x_free[b])
        with torch.cuda.stream(self.h2d):
            self.xbuf[b].copy_(x_host, non_blocking=True)
            self.x_ready[b].record(self.h2d)

    def step(self, next_x_host: Optional[torch.Tensor], grad_host: torch.Tensor):
        from . import dist as ndist
        b = self.slot
        self.slot ^= 1
        if next_x_host is not None:
            self.prefetch(next_x_host)                              # goes into the other buffer, overlaps the kernels below
        self.compute.wait_event(self.x_ready[b])
        x = self.xbuf[b]
        y, _ = self._lotd.lod_fwd(self.meta, x, self.params, need_input_grad=False)
        gy = self.grad_of_y(y)
        _, g = self._lotd.lod_bwd(self.meta, gy, x, self.params, None, need_input_grad=False, need_param_grad=True)
        self.x_free[b].record(self.compute)
        lo, hi = 0, g.shape[0]
        out = g
        if self.world > 1:
            # reduce-scatter: this rank now holds the summed gradient of its own 1/world slice and returns exactly that slice to the host
            out = self.reducer.reduce(g)
            lo, hi = self.reducer.slice_range(g.shape[0], self.rank)
        done = torch.cuda.Event()
        done.record(self.compute)
        self.d2h.wait_event(done)
        with torch.cuda.stream(self.d2h):
            grad_host[lo:hi].copy_(out[: hi - lo], non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.d2h)
        out.record_stream(self.d2h)
        self.inflight = [(e, t) for (e, t) in self.inflight if not e.query()] + [(copied, out)]
        return out

    def drain(self):
        for e, _ in self.inflight:
            self.compute.wait_event(e)
        self.inflight = []
