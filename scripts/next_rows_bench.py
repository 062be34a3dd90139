#!/usr/bin/env python
"""Timing of the widened rows on one B200: this build vs the reference's own CUDA build
(oracle/_ref/_pack_ops.so, n2) / vs the reference's torch composition with scatter_reduce_ standing in for the missing
torch_scatter (n1).  CUDA-event timed, 5 warm-up + 20 timed calls, median.  Usage:  python scripts/next_rows_bench.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import load_ref  # noqa: E402

dev = torch.device("cuda:0")


def timeit(fn, warm=5, it=20):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(it):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def row(name, ours, ref, unit_bytes=None):
    s = f"{name:46s} ours {ours * 1e3:9.1f} us"
    if ref is not None:
        s += f"   reference {ref * 1e3:9.1f} us   x{ref / ours:6.2f}"
    if unit_bytes:
        s += f"   {unit_bytes / ours / 1e6:8.1f} GB/s algorithmic"
    print(s, flush=True)


def bench_n2():
    from nr3d_lib_b200.bindings import _pack_ops as ours
    ref = load_ref("_pack_ops")
    g = torch.Generator(device=dev); g.manual_seed(0)
    P, L, NQ = 2 ** 18, 64, 32
    S = P * L
    pi = torch.stack([torch.arange(P, device=dev) * L, torch.full([P], L, device=dev)], 1).long().contiguous()
    bins = (torch.rand(P, L, device=dev, generator=g).sort(-1).values + torch.arange(P, device=dev)[:, None]).reshape(-1).contiguous()
    w = torch.rand(P, L, device=dev, generator=g) + 0.01
    cdfs = (w.cumsum(-1) / w.sum(-1, keepdim=True)).reshape(-1).contiguous()
    u = torch.rand(P, NQ, device=dev, generator=g)
    vq = u + torch.arange(P, device=dev)[:, None]
    pib = torch.stack([torch.arange(P, device=dev) * NQ, torch.full([P], NQ, device=dev)], 1).long().contiguous()
    vb = vq.sort(-1).values.reshape(-1).contiguous()
    unsorted = torch.randn(S, device=dev, generator=g)
    print(f"# n2: {P} packs x {L} samples ({S / 1e6:.1f} M), {NQ} queries per pack")
    for name, call, nbytes in (
        ("packed_searchsorted", lambda be: be.packed_searchsorted(bins, vq, pi), P * NQ * 12 + S * 4),
        ("packed_searchsorted_packed_vals", lambda be: be.packed_searchsorted_packed_vals(bins, pi, vb, pib), P * NQ * 12 + S * 4),
        ("packed_invert_cdf", lambda be: be.packed_invert_cdf(bins, cdfs, u, pi), P * NQ * 16 + S * 8),
        ("try_merge_two_packs_sorted_aligned", lambda be: be.try_merge_two_packs_sorted_aligned(bins, pi, vb, pib, True), (S + P * NQ) * 12),
        ("packed_sort_qsort (in place + idx)", lambda be: be.packed_sort_qsort(unsorted.clone(), pi, True), S * 16),
    ):
        t_o = timeit(lambda: call(ours))
        t_r = timeit(lambda: call(ref)) if ref is not None else None
        row(name, t_o, t_r, nbytes)
    # segment sampler
    nseg = 4
    spi = torch.stack([torch.arange(P, device=dev) * nseg, torch.full([P], nseg, device=dev)], 1).long().contiguous()
    cuts = (torch.rand(P, 2 * nseg, device=dev, generator=g) * 6 + 0.05).sort(-1).values
    entry, exit_ = cuts[:, 0::2].reshape(-1).contiguous(), cuts[:, 1::2].reshape(-1).contiguous()
    near = torch.rand(P, device=dev, generator=g)
    far = near + 5.0
    args = (near, far, entry, exit_, spi, 256, 0.01, 0.01, 0.2)
    n_out = ours.interleave_sample_step_wrt_depth_in_packed_segments(*args)[0].shape[0]
    t_o = timeit(lambda: ours.interleave_sample_step_wrt_depth_in_packed_segments(*args))
    t_r = timeit(lambda: ref.interleave_sample_step_wrt_depth_in_packed_segments(*args)) if ref is not None else None
    row(f"sample_step_in_packed_segments ({n_out / 1e6:.1f} M out)", t_o, t_r, n_out * 24)


def bench_n1():
    from nr3d_lib_b200 import occgrid as G
    R, N = 128, 2 ** 20
    g = torch.Generator(device=dev); g.manual_seed(0)
    grid0 = torch.rand(R, R, R, device=dev, generator=g)
    pts = torch.rand(N, 3, device=dev, generator=g) * 2 - 1
    vals = torch.rand(N, device=dev, generator=g)
    res = torch.tensor([R, R, R], device=dev)

    def reference_composition(grid):      # utils.py:93-110 + binarize, scatter_max -> scatter_reduce_(amax, include_self=True)
        gidx = ((pts / 2. + 0.5) * res).long().clamp(res.new_tensor([0]), res - 1)
        ravel = (gidx * gidx.new_tensor([R * R, R, 1])).sum(-1)
        new = (0.95 * grid.flatten()).scatter_reduce_(0, ravel, vals, reduce="amax", include_self=True)
        grid.index_put_(tuple(gidx.t()), new[ravel])
        return grid > 0.5

    occ = torch.zeros(R, R, R, dtype=torch.bool, device=dev)
    ga, gb = grid0.clone(), grid0.clone()
    G.update_and_binarize_(ga, occ, pts=pts, occ_val=vals, ema_decay=0.95, occ_threshold=0.5)
    occ_ref = reference_composition(gb)
    same = torch.equal(ga, gb) and torch.equal(occ, occ_ref)
    print(f"# n1: {R}^3 value grid, {N} samples, EMA update + binarize; identical to the torch composition: {same}")
    t_o = timeit(lambda: G.update_and_binarize_(ga, occ, pts=pts, occ_val=vals, ema_decay=0.95, occ_threshold=0.5))
    t_r = timeit(lambda: reference_composition(gb))
    row("update_occ_val_grid_ + binarize", t_o, t_r, N * 16 + R ** 3 * 9)
    t_o = timeit(lambda: G.update_and_binarize_(ga, occ, pts=pts, occ_val=vals, ema_decay=0.95, occ_threshold=0.5, consider_mean=True))
    row("  ... with mean-relative threshold", t_o, None, N * 16 + R ** 3 * 13)
    vox = occ.nonzero().long()
    t_o = timeit(lambda: G.sample_pts_in_voxels(vox, 2 ** 18, res))

    def ref_sample():
        vidx = torch.randint(vox.shape[0], [2 ** 18], device=dev)
        off = torch.rand([2 ** 18, 3], device=dev)
        return ((vox[vidx] + off) / res.float()) * 2 - 1
    row("sample_pts_in_voxels (2^18 pts, incl. RNG)", t_o, timeit(ref_sample), 2 ** 18 * 56)
    t_o = timeit(lambda: G.query_occ_grid(occ, pts))
    t_r = timeit(lambda: occ[tuple(((pts / 2. + 0.5) * res).long().clamp(res.new_tensor([0]), res - 1).movedim(-1, 0))])
    row("query (2^20 pts)", t_o, t_r, N * 13)


if __name__ == "__main__":
    print(f"# {torch.cuda.get_device_name(0)}")
    bench_n1()
    bench_n2()
