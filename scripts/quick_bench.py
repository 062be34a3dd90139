"""Quick kernel timing (mine vs the reference CUDA build) on the C2 config: 16-level NGP LoTD, 4M points."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests.util import load_ref
from nr3d_lib_b200.bindings import _lotd as mine

def ngp_cfg(min_res=16, n_levels=16, scale=1.382, log2_T=19, F=2):
    res = (min_res * scale ** np.arange(n_levels)).astype(int).tolist()
    types = ["Dense" if r ** 3 <= 2 ** log2_T else "Hash" for r in res]
    return (3, res, [F] * n_levels, types, 2 ** log2_T, False)

def timeit(fn, iters=10, warm=3, pre=None):
    """median device time of fn(); `pre` runs untimed before every call (e.g. to sort another point set so that fn re-sorts)."""
    for _ in range(warm):
        if pre: pre()
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if pre: pre()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return float(np.median(ts))

def main():
    dev = torch.device("cuda:0")
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 4 * 1024 * 1024
    args = ngp_cfg()
    torch.manual_seed(42)
    x = torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6)
    for pd in (torch.float32, torch.float16):
        for name, be in (("mine", mine), ("ref", load_ref("_lotd"))):
            if be is None: continue
            meta = be.LoDMeta(*args)
            params = ((torch.rand(meta.n_params, device=dev) * 2 - 1) * 1e-4).to(pd)
            dL_dy = (torch.randn(N, meta.n_encoded_dims, device=dev) * 1e-4).to(pd)
            t_f = timeit(lambda: be.lod_fwd(meta, x, params, need_input_grad=False))
            t_fd = timeit(lambda: be.lod_fwd(meta, x, params, need_input_grad=True))
            y, dydx = be.lod_fwd(meta, x, params, need_input_grad=True)
            t_b = timeit(lambda: be.lod_bwd(meta, dL_dy, x, params, None, need_input_grad=False, need_param_grad=True))
            t_bx = timeit(lambda: be.lod_bwd(meta, dL_dy, x, params, dydx, need_input_grad=True, need_param_grad=False))
            gy = dL_dy.t().contiguous().t()  # feature-major dL_dy
            t_b2 = timeit(lambda: be.lod_bwd(meta, gy, x, params, None, need_input_grad=False, need_param_grad=True))
            print(f"{name:5s} {str(pd):14s} N={N} fwd {t_f:7.3f} ms | fwd+dydx {t_fd:7.3f} | bwd_param {t_b:7.3f} (feat-major dLdy {t_b2:7.3f}) | bwd_dx {t_bx:7.3f} | "
                  f"fwd+bwd {N / (t_f + t_b) / 1e3:8.1f} Msamples/s", flush=True)
            del y, dydx

if __name__ == '__main__':
    main()
