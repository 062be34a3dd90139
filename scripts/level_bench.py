"""Per-level cost of the LoTD kernels: one single-level meta per NGP level (4M uniform points)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from scripts.quick_bench import timeit, ngp_cfg  # noqa
from nr3d_lib_b200.bindings import _lotd as mine
from tests.util import load_ref
dev = torch.device("cuda:0")
N = 4 * 1024 * 1024
torch.manual_seed(0)
x = torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6)
xs = x[torch.argsort((x[:, 2] * 256).floor() * 65536 + (x[:, 1] * 256).floor() * 256 + (x[:, 0] * 256).floor())].contiguous()
_, res, _, types, T, _ = ngp_cfg()
ref = load_ref("_lotd")
for r, tp in zip(res, types):
    row = f"res {r:5d} {tp:6s}"
    for name, be, xx in (("mine", mine, x), ("mine-sorted", mine, xs), ("ref", ref, x)):
        if be is None: continue
        meta = be.LoDMeta(3, [r], [2], [tp], T, False)
        p = torch.randn(meta.n_params, device=dev) * 1e-2
        g = torch.randn(N, 2, device=dev)
        tf = timeit(lambda: be.lod_fwd(meta, xx, p, need_input_grad=False), iters=5, warm=2)
        tb = timeit(lambda: be.lod_bwd(meta, g, xx, p, None, need_input_grad=False, need_param_grad=True), iters=5, warm=2)
        row += f" | {name}: fwd {tf*1e3:7.1f} us bwd {tb*1e3:7.1f} us"
    print(row, flush=True)
