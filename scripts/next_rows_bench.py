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


def bench_n3():
    from nr3d_lib_b200.bindings import _lotd
    from nr3d_lib_b200.fused import FusedDensityDecoder
    res = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
    meta = _lotd.LoDMeta(3, res, [2] * 16, ["Dense" if r ** 3 <= 2 ** 19 else "Hash" for r in res], 2 ** 19)
    meta.c_sort_points = True
    N = 4 * 2 ** 20
    g = torch.Generator(device=dev); g.manual_seed(0)
    x = torch.rand(N, 3, device=dev, generator=g).clamp(1e-6, 1 - 1e-6)
    params = (torch.rand(meta.n_params, device=dev, generator=g) * 2 - 1) * 1e-1
    w1, b1 = torch.randn(64, 32, device=dev, generator=g) * 0.3, torch.randn(64, device=dev, generator=g) * 0.1
    w2, b2 = torch.randn(16, 64, device=dev, generator=g) * 0.2, torch.randn(16, device=dev, generator=g) * 0.1
    dec = FusedDensityDecoder(meta, w1, b1, w2, b2, activation="exp")
    w1h, w2h = w1.half(), w2.half()

    def unfused_fp32():      # the reference's composition: features to HBM, two GEMM launches, activation
        _lotd.clear_sort_cache()
        h, _ = _lotd.lod_fwd(meta, x, params, need_input_grad=False)
        return (torch.relu(h @ w1.t() + b1) @ w2.t() + b2)[:, 0].exp()

    def unfused_fp16_mlp():  # same with a half-precision MLP (what a tcnn-style decoder does)
        _lotd.clear_sort_cache()
        h, _ = _lotd.lod_fwd(meta, x, params, need_input_grad=False)
        return (torch.relu(h.half() @ w1h.t() + b1.half()) @ w2h.t() + b2.half())[:, 0].float().exp()

    def fused():
        _lotd.clear_sort_cache()
        return dec.query_density(x, params)[0]

    def encode_only():
        _lotd.clear_sort_cache()
        return _lotd.lod_fwd(meta, x, params, need_input_grad=False)[0]
    err = ((fused() - unfused_fp32()).abs().max() / unfused_fp32().abs().max()).item()
    print(f"# n3: query_density on {N} points, 16-level NGP LoTD + MLP 32-64-16, exp; fused vs fp32 composition max rel err {err:.2e}")
    t_f = timeit(fused)
    row("fused encode + decoder (tcgen05)", t_f, timeit(unfused_fp32), N * (12 + 1024 + 4))
    row("  vs unfused with fp16 MLP", t_f, timeit(unfused_fp16_mlp))
    row("  (encode alone, features to HBM)", timeit(encode_only), None, N * (12 + 1024 + 128))


def bench_n3_train():
    """Training step through encoder + density decoder: forward, loss, backward to the LoTD tables and the decoder weights."""
    from nr3d_lib_b200.bindings import _lotd
    from nr3d_lib_b200.fused import fused_density
    from nr3d_lib_b200.lotd import LoTDFunction
    res = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
    meta = _lotd.LoDMeta(3, res, [2] * 16, ["Dense" if r ** 3 <= 2 ** 19 else "Hash" for r in res], 2 ** 19)
    N = 4 * 2 ** 20
    g = torch.Generator(device=dev); g.manual_seed(0)
    xs = [torch.rand(N, 3, device=dev, generator=g).clamp(1e-6, 1 - 1e-6) for _ in range(2)]   # alternating point sets: every step sorts
    params = ((torch.rand(meta.n_params, device=dev, generator=g) * 2 - 1) * 1e-1).requires_grad_(True)
    w1 = (torch.randn(64, 32, device=dev, generator=g) * 0.3).requires_grad_(True)
    b1 = (torch.randn(64, device=dev, generator=g) * 0.1).requires_grad_(True)
    w2 = (torch.randn(1, 64, device=dev, generator=g) * 0.05).requires_grad_(True)
    b2 = (torch.randn(1, device=dev, generator=g) * 0.1).requires_grad_(True)
    target = torch.rand(N, device=dev, generator=g)
    leaves = [params, w1, b1, w2, b2]
    k = [0]

    def zero():
        for t in leaves:
            t.grad = None
        k[0] ^= 1
        return xs[k[0]]

    def step_fused():
        x = zero()
        sigma, _ = fused_density(x, params, w1, b1, w2, b2, meta, activation="softplus")
        ((sigma - target) ** 2).mean().backward()

    def step_unfused(mlp_dtype):
        x = zero()
        h = LoTDFunction.apply(meta, x, params, None, None, 0, 1.0, None).to(mlp_dtype)
        out = torch.relu(h @ w1.to(mlp_dtype).t() + b1.to(mlp_dtype)) @ w2.to(mlp_dtype).t() + b2.to(mlp_dtype)
        sigma = torch.nn.functional.softplus(out[:, 0].float())
        ((sigma - target) ** 2).mean().backward()

    step_fused(); gf = [t.grad.clone() for t in leaves]
    k[0] ^= 1
    step_unfused(torch.float32); gu = [t.grad.clone() for t in leaves]
    errs = [((a - b).abs().max() / b.abs().max()).item() for a, b in zip(gf, gu)]
    print(f"# n3 training step on {N} points (16-level NGP LoTD + MLP 32-64-1, softplus, MSE): fused vs fp32 autograd max rel err of "
          f"dparams / dW1 / db1 / dW2 / db2 = " + " / ".join(f"{e:.1e}" for e in errs))
    t_f = timeit(step_fused, warm=3, it=10)
    row("fused fwd + bwd (tcgen05, 2 kernels + sort)", t_f, timeit(lambda: step_unfused(torch.float32), warm=3, it=10))
    row("  vs unfused with bf16 MLP (cuBLAS)", t_f, timeit(lambda: step_unfused(torch.bfloat16), warm=3, it=10))

    def enc_only():
        x = zero()
        y, _ = _lotd.lod_fwd(meta, x, params.detach(), need_input_grad=False)
        _lotd.lod_bwd(meta, y, x, params.detach(), None, need_input_grad=False, need_param_grad=True)
    row("  (encoder fwd + bwd alone, features in HBM)", timeit(enc_only, warm=3, it=10), None)


if __name__ == "__main__":
    print(f"# {torch.cuda.get_device_name(0)}")
    if len(sys.argv) > 1 and sys.argv[1] == "n3":
        bench_n3()
        bench_n3_train()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "n3train":
        bench_n3_train()
        sys.exit(0)
    bench_n1()
    bench_n2()
    bench_n3()
    bench_n3_train()
