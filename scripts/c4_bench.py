"""BASELINE.json configs[3]: batched LoTD, B = 8 scenes, mixed Dense / VM / CP levels with mixed feature widths,
2 Mi points with per-point batch indices (1 % skipped), first- and second-order passes.  Ours vs the reference CUDA build (if present).

    python scripts/c4_bench.py [N]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from nr3d_lib_b200.bindings import _lotd as mine  # noqa: E402
from scripts.quick_bench import timeit  # noqa: E402
from tests.util import load_ref  # noqa: E402

CFG = (3, [8, 16, 32, 64, 128, 256], [4, 4, 4, 4, 2, 2], ["Dense", "Dense", "VM", "VM", "CP", "CP"], None, False)


def main():
    N = int(sys.argv[1]) if len(sys.argv) > 1 else 2 * 1024 * 1024
    dev = torch.device("cuda:0")
    B = 8
    torch.manual_seed(42)
    x = torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6)
    bi = torch.randint(0, B, (N,), device=dev)
    bi[torch.rand(N, device=dev) < 0.01] = -1
    rows = {}
    for name, be in (("mine", mine), ("ref", load_ref("_lotd"))):
        if be is None:
            continue
        meta = be.LoDMeta(*CFG)
        E = meta.n_encoded_dims
        g = torch.Generator(device=dev).manual_seed(1)
        params = torch.randn(B * meta.n_params, device=dev, generator=g) * 1e-2
        dL_dy = torch.randn(N, E, device=dev, generator=g) * 1e-2
        ddx = torch.randn(N, 3, device=dev, generator=g)
        kw = dict(batch_inds=bi, batch_offsets=None, batch_data_size=None, max_level=None)
        y, dydx = be.lod_fwd(meta, x, params, need_input_grad=True, **kw)
        t = {}
        t["fwd"] = timeit(lambda: be.lod_fwd(meta, x, params, need_input_grad=False, **kw))
        t["fwd+dydx"] = timeit(lambda: be.lod_fwd(meta, x, params, need_input_grad=True, **kw))
        t["bwd dparam"] = timeit(lambda: be.lod_bwd(meta, dL_dy, x, params, None, need_input_grad=False, need_param_grad=True, **kw))
        t["bwd dx"] = timeit(lambda: be.lod_bwd(meta, dL_dy, x, params, dydx, need_input_grad=True, need_param_grad=False, **kw))
        t["bwd_bwd (ddLdy, dparam)"] = timeit(lambda: be.lod_bwd_bwd_input(meta, ddx, dL_dy, x, params, dydx, need_dLdinput_ddLdoutput=True,
                                                                          need_dLdinput_dparams=True, need_dLdinput_dinput=False, **kw))
        t["bwd_bwd (dx)"] = timeit(lambda: be.lod_bwd_bwd_input(meta, ddx, dL_dy, x, params, dydx, need_dLdinput_ddLdoutput=False,
                                                               need_dLdinput_dparams=False, need_dLdinput_dinput=True, **kw))
        rows[name] = t
        print(f"{name:5s} N={N} B={B} n_params/scene={meta.n_params} n_enc={E}: " + " | ".join(f"{k} {v:7.3f} ms" for k, v in t.items()), flush=True)
    if "ref" in rows:
        print("speed-up vs reference build: " + " | ".join(f"{k} x{rows['ref'][k] / rows['mine'][k]:.2f}" for k in rows["mine"]))
    step = sum(rows["mine"][k] for k in ("fwd+dydx", "bwd dparam", "bwd dx", "bwd_bwd (ddLdy, dparam)"))
    print(json.dumps({"config": "C4", "N": N, "ms_step_mine": step, "Msamples_per_s": N / step / 1e3,
                      "ms_step_ref": sum(rows["ref"][k] for k in ("fwd+dydx", "bwd dparam", "bwd dx", "bwd_bwd (ddLdy, dparam)")) if "ref" in rows else None}))


if __name__ == "__main__":
    main()
