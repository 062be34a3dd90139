"""NeuS-style call pattern on the headline encoder (16-level NGP LoTD, 4 Mi points, fp32): forward with dy/dx, first-order backward
(dL/dx + dL/dparam) and the second-order backward (dL_ddLdy + d(dL/dx)/dparam), fast path vs generic kernels vs the reference build.

    python scripts/nablas_bench.py [N]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import ngp_cfg  # noqa: E402
from nr3d_lib_b200.bindings import _lotd as mine  # noqa: E402
from scripts.quick_bench import timeit  # noqa: E402
from tests.util import load_ref  # noqa: E402


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
    dev = torch.device("cuda:0")
    torch.manual_seed(42)
    x = torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6)
    x_other = torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6)    # sorted between timed forwards so that every forward re-sorts
    rows = {}
    for name, be, sort in (("fast path", mine, True), ("generic", mine, False), ("reference build", load_ref("_lotd"), False)):
        if be is None:
            continue
        meta = be.LoDMeta(*ngp_cfg())
        if be is mine:
            meta.c_sort_points = sort
        params = (torch.rand(meta.n_params, device=dev) * 2 - 1) * 1e-4
        dL_dy = torch.randn(N, meta.n_encoded_dims, device=dev) * 1e-4
        ddx = torch.randn(N, 3, device=dev)
        y, dydx = be.lod_fwd(meta, x, params, need_input_grad=True)

        def fwd():
            be.lod_fwd(meta, x, params, need_input_grad=True)
        if sort:   # a new batch of points every step: fingerprint + sort are part of the forward (the other point set is sorted untimed)
            def pre():
                mine._sorted_points(x_other)
        else:
            pre = None
        t = {"fwd+dydx": timeit(fwd, pre=pre),
             "bwd (dx, dparam)": timeit(lambda: be.lod_bwd(meta, dL_dy, x, params, dydx, need_input_grad=True, need_param_grad=True)),
             "bwd_bwd (ddLdy, dparam)": timeit(lambda: be.lod_bwd_bwd_input(meta, ddx, dL_dy, x, params, dydx, need_dLdinput_ddLdoutput=True,
                                                                           need_dLdinput_dparams=True, need_dLdinput_dinput=False))}
        t["step"] = sum(t.values())
        rows[name] = t
        print(f"{name:16s} N={N}: " + " | ".join(f"{k} {v:7.3f} ms" for k, v in t.items()) + f" | {N / t['step'] / 1e3:8.1f} Msamples/s", flush=True)
        del y, dydx
    if "reference build" in rows:
        print("fast path vs reference build: " + " | ".join(f"{k} x{rows['reference build'][k] / rows['fast path'][k]:.2f}" for k in rows["fast path"]))


if __name__ == "__main__":
    main()
