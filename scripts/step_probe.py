"""Per-phase device times of the headline step (16-level NGP LoTD, 4 Mi uniform points) for one build of the library:
sort (verify + hist + scan + scatter), forward, backward -- CUDA events on the launch stream, median of `iters` after warm-up, two
alternating point sets so that every forward really re-sorts and every backward hits the verified records.

    python scripts/step_probe.py [--half] [--iters 20]            # NR3D_B200_LIB selects an A/B build (scripts/ab_bench.py)
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ngp_cfg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--half", action="store_true")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--points", type=int, default=4 * 1024 * 1024)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    from nr3d_lib_b200 import _lib
    from nr3d_lib_b200.bindings import _lotd
    dev = torch.device("cuda:0")
    meta = _lotd.LoDMeta(*ngp_cfg())
    N = args.points
    torch.manual_seed(42)
    xs = [torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6) for _ in range(2)]
    dt = torch.float16 if args.half else torch.float32
    params = ((torch.rand(meta.n_params, device=dev) * 2 - 1) * 1e-4).to(dt)
    dL_dy = (torch.randn(N, meta.n_encoded_dims, device=dev) * 1e-4).to(dt)
    st = torch.cuda.current_stream(dev)

    def timed(fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st)
        torch.cuda.synchronize()
        return a.elapsed_time(b)
    res = {"sort": [], "fwd": [], "bwd": [], "step": []}
    for it in range(args.iters + 3):
        x = xs[it & 1]
        t_sort = timed(lambda: _lotd._sorted_points(x))                       # new points: verify says "different", full sort
        t_fwd = timed(lambda: _lotd.lod_fwd(meta, x, params, need_input_grad=False))   # verify says "same": records reused
        t_bwd = timed(lambda: _lotd.lod_bwd(meta, dL_dy, x, params, None, need_input_grad=False, need_param_grad=True))
        x2 = xs[(it + 1) & 1]

        def step():
            _lotd.lod_fwd(meta, x2, params, need_input_grad=False)
            _lotd.lod_bwd(meta, dL_dy, x2, params, None, need_input_grad=False, need_param_grad=True)
        t_step = timed(step)
        if it >= 3:
            for k, v in (("sort", t_sort), ("fwd", t_fwd), ("bwd", t_bwd), ("step", t_step)):
                res[k].append(v)
    out = {k: float(np.median(v)) for k, v in res.items()}
    # checksums of one forward / backward: the same for every build of the library up to fp32 summation order
    y, _ = _lotd.lod_fwd(meta, xs[0], params, need_input_grad=False)
    _, g = _lotd.lod_bwd(meta, dL_dy, xs[0], params, None, need_input_grad=False, need_param_grad=True)
    out["checksum_y"], out["checksum_grad"] = float(y.double().abs().sum()), float(g.double().abs().sum())
    # the gradient table of the first build that runs is kept; later builds (A/B variants in the same gpurun call) report their distance to it
    ref_file = os.path.join(os.environ.get("NR3D_AB_REF_DIR", "/tmp"), f"ab_ref_grad_{'f16' if args.half else 'f32'}_{N}.pt")
    if os.path.exists(ref_file):
        ref = torch.load(ref_file, map_location=dev)
        out["grad_max_abs_diff_vs_first"] = float((g.double() - ref.double()).abs().max())
        out["grad_max_abs"] = float(ref.double().abs().max())
    else:
        torch.save(g, ref_file)
    out.update(tag=args.tag or os.path.basename(_lib.LIB_PATH), half=args.half, msamples_per_s=N / out["step"] / 1e3,
               note="fwd = lod_fwd of new points (sort + gather), bwd = lod_bwd (fingerprint check + scatter), step = fwd + bwd back to back")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
