"""Which build of the reference's generic LoTD kernels returns a dy/dx that passes finite differences of its own forward?

Round 1 found that the reference's GENERIC kernels (csrc/lotd/include/lotd/lotd_encoding.h:31-111, `#pragma unroll 1` over grad_dim with a
dynamically indexed local array), built with nvcc 12.9 -O3 for sm_100, return a dy/dx for the n-linear level types at D >= 3 that contradicts
central differences of the same build's forward.  This script loads ONE build per process (pybind type registration is process-global),
runs the identity on the `mixed`, `cuboid_vm`, `batched` and `d4` configurations and prints one JSON line.  Variants come from
`python oracle/build_ref.py --variant {O1,O0,G}` (only the generic-kernel TUs compile_split_*.cu get the extra ptxas flag).

    python scripts/ref_variant_check.py            # stock -O3 build (oracle/_ref/_lotd.so)
    python scripts/ref_variant_check.py O1         # oracle/_ref/_lotd__O1.so
"""
import importlib.util
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import LOTD_CONFIGS, lotd_inputs, meta_args, rel_err  # noqa: E402


def load(variant):
    path = os.path.join(ROOT, "oracle", "_ref", "_lotd" + (("__" + variant) if variant else "") + ".so")
    spec = importlib.util.spec_from_file_location("_lotd", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def central_diff(backend, meta, x, params, h, kw):
    cols = []
    for d in range(x.shape[1]):
        e = torch.zeros_like(x)
        e[:, d] = h
        yp, _ = backend.lod_fwd(meta, (x + e).contiguous(), params, need_input_grad=False, **kw)
        ym, _ = backend.lod_fwd(meta, (x - e).contiguous(), params, need_input_grad=False, **kw)
        cols.append((yp.double() - ym.double()) / ((x + e)[:, d:d + 1].double() - (x - e)[:, d:d + 1].double()))
    return torch.stack(cols, -1)


def main():
    variant = sys.argv[1] if len(sys.argv) > 1 else ""
    ref = load(variant)
    from nr3d_lib_b200.bindings import _lotd as mine
    dev = torch.device("cuda:0")
    out = {"variant": variant or "stock -O3"}
    for name in ("mixed", "cuboid_vm", "batched", "d4"):
        cfg = LOTD_CONFIGS[name]
        m_ref, m_mine = ref.LoDMeta(*meta_args(cfg)), mine.LoDMeta(*meta_args(cfg))
        inp = lotd_inputs(cfg, m_mine.n_params, N=4000, seed=17)
        h = 1.0e-4
        x = inp["x"].clamp(0.01, 0.99)
        keep = torch.ones(x.shape[0], dtype=torch.bool)
        for R in m_mine.level_res_multidim:
            s = torch.tensor([r - 2 for r in R], dtype=torch.float64)
            keep &= (torch.floor((x.double() + 2 * h) * s + 0.5) == torch.floor((x.double() - 2 * h) * s + 0.5)).all(-1)
        bi = inp["batch_inds"]
        kw = {}
        if bi is not None:
            kw = dict(batch_inds=bi[keep].to(dev).contiguous())
        x = x[keep].to(dev).contiguous()
        params = inp["params"].to(dev)
        N, E, D = x.shape[0], m_mine.n_encoded_dims, m_mine.n_dims_to_encode
        _, dy_r = ref.lod_fwd(m_ref, x, params, need_input_grad=True, **kw)
        fd_r = central_diff(ref, m_ref, x, params, h, kw)
        _, dy_m = mine.lod_fwd(m_mine, x, params, need_input_grad=True, **kw)
        dy_r = dy_r.reshape(N, E, D).double().cpu()
        dy_m = dy_m.reshape(N, E, D).double().cpu()
        # second order: d(dL/dx)/dparam and dL_ddLdy of the two builds against each other
        ddx = inp["dL_ddLdx"][keep].to(dev).contiguous()
        dL_dy = inp["dL_dy"][keep].to(dev).contiguous()
        _, dyx_r = ref.lod_fwd(m_ref, x, params, need_input_grad=True, **kw)
        _, dyx_m = mine.lod_fwd(m_mine, x, params, need_input_grad=True, **kw)
        a = ref.lod_bwd_bwd_input(m_ref, ddx, dL_dy, x, params, dyx_r, need_dLdinput_ddLdoutput=True, need_dLdinput_dparams=True, need_dLdinput_dinput=True, **kw)
        b = mine.lod_bwd_bwd_input(m_mine, ddx, dL_dy, x, params, dyx_m, need_dLdinput_ddLdoutput=True, need_dLdinput_dparams=True, need_dLdinput_dinput=True, **kw)
        out[name] = dict(ref_dydx_vs_own_fd=rel_err(dy_r, fd_r.cpu()), ours_vs_ref_dydx=rel_err(dy_m, dy_r),
                         ours_vs_ref_ddLdy=rel_err(b[0].double().cpu(), a[0].double().cpu()),
                         ours_vs_ref_dparam2=rel_err(b[1].double().cpu(), a[1].double().cpu()),
                         ours_vs_ref_dx2=rel_err(b[2].double().cpu(), a[2].double().cpu()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
