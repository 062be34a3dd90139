"""Core pack ops and the marcher against the reference's CUDA build (oracle/_ref) on the M2 shapes: 262144 rays x ~115 samples
(30 Mi samples, what one M2 chunk composites), CUDA events, median of 10.  Reports ms and the achieved fraction of the HBM bandwidth
for the bytes each op has to move.

    python scripts/pack_bench.py > profiles/r2_pack_march_bench.txt
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nr3d_lib_b200.bindings import _occ_grid, _pack_ops  # noqa: E402
from tests.util import load_ref  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    dev = torch.device("cuda:0")
    peak = 6553.9
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    g = torch.Generator(device=dev).manual_seed(0)
    P = 262144
    n = torch.randint(32, 200, (P,), device=dev, generator=g)
    pi = torch.stack([torch.cumsum(n, 0) - n, n], 1).contiguous()
    S = int(n.sum())
    alphas = (torch.rand(S, device=dev, generator=g) ** 2 * 0.2).contiguous()
    gw = torch.randn(S, device=dev, generator=g)
    ref_p, ref_o = load_ref("_pack_ops"), load_ref("_occ_grid")
    rows = []

    def row(name, nbytes, mine, ref):
        tm = timeit(mine)
        tr = timeit(ref) if ref is not None else float("nan")
        rows.append((name, tm, tr, nbytes / tm / 1e6, nbytes / tm / 1e6 / peak))

    w = _pack_ops.packed_alpha_to_vw_forward(alphas, pi, 1e-4, 0.0, False)[0]
    row("packed_alpha_to_vw_forward (weights)", S * 8 + P * 16,
        lambda: _pack_ops.packed_alpha_to_vw_forward(alphas, pi, 1e-4, 0.0, False),
        (lambda: ref_p.packed_alpha_to_vw_forward(alphas, pi, 1e-4, 0.0, False)) if ref_p else None)
    row("packed_alpha_to_vw_backward", S * 16 + P * 16,
        lambda: _pack_ops.packed_alpha_to_vw_backward(w, gw, alphas, pi, 1e-4, 0.0),
        (lambda: ref_p.packed_alpha_to_vw_backward(w, gw, alphas, pi, 1e-4, 0.0)) if ref_p else None)
    row("packed_sum (1 channel)", S * 4 + P * 20,
        lambda: _pack_ops.packed_sum(w, pi), (lambda: ref_p.packed_sum(w, pi)) if ref_p else None)
    row("packed_cumsum (1 channel)", S * 8 + P * 16,
        lambda: _pack_ops.packed_cumsum(w, pi, False, False), (lambda: ref_p.packed_cumsum(w, pi, False, False)) if ref_p else None)
    row("packed_cumprod (1 channel, inclusive)", S * 8 + P * 16,
        lambda: _pack_ops.packed_cumprod(alphas, pi, False, False), (lambda: ref_p.packed_cumprod(alphas, pi, False, False)) if ref_p else None)
    # marcher: 262144 rays through a random 128^3 grid, step 0.01 (the M2 chunk)
    from scripts.m2_bench import make_rays
    o, d, near, far = make_rays(P, dev, 7)
    grid = torch.rand(128, 128, 128, device=dev, generator=g) > 0.5
    roi = torch.tensor([-1.0, -1.0, -1.0, 1.0, 1.0, 1.0], device=dev)
    CT = _occ_grid.ContractionType.AABB
    r = _occ_grid.ray_marching(o, d, near, far, roi, grid, CT, 0.01, 1e10, 0.0, 512, True)
    Sm = int(r[1].shape[0])
    row(f"ray_marching ({Sm / P:.0f} samples/ray), record + compact", Sm * 16 + P * 40,
        lambda: _occ_grid.ray_marching(o, d, near, far, roi, grid, CT, 0.01, 1e10, 0.0, 512, True),
        (lambda: ref_o.ray_marching(o, d, near, far, roi, grid, ref_o.ContractionType.AABB, 0.01, 1e10, 0.0, 512, True)) if ref_o else None)
    budget, _occ_grid.MARCH_SCRATCH_BYTES = _occ_grid.MARCH_SCRATCH_BYTES, 0
    row("ray_marching, count + fill (two passes)", Sm * 16 + P * 40,
        lambda: _occ_grid.ray_marching(o, d, near, far, roi, grid, CT, 0.01, 1e10, 0.0, 512, True), None)
    _occ_grid.MARCH_SCRATCH_BYTES = budget
    print(f"# {P} packs / rays, {S} samples ({S / P:.0f} per pack), fp32, B200; peak = {peak:.0f} GB/s (MEASURED_PEAKS.json)")
    print(f"{'op':58s} {'ours ms':>9s} {'ref build ms':>13s} {'speed-up':>9s} {'GB/s':>8s} {'of HBM':>7s}")
    for name, tm, tr, gbs, frac in rows:
        print(f"{name:58s} {tm:9.3f} {tr:13.3f} {tr / tm:9.2f} {gbs:8.0f} {frac:7.2f}")


if __name__ == "__main__":
    main()
