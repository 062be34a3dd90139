"""Subprocess helper of the GPU tests: run LoTD jobs on ONE named build of the reference's own CUDA extension and save every output.

    python tests/ref_worker.py <variant> '<json list of jobs>' <out.npz>

Only one build of the reference's `_lotd` can live in a process (pybind registers its C++ types globally), so the tests -- which load the
stock -O3 build for the hash-only kernels -- call this script for the `-G` checker build of the generic kernels (oracle/build_ref.py
--variant G; why: scripts/ref_variant_check.py).  Inputs are regenerated from the seeded generators of tests/util.py, so nothing but the job
description crosses the process boundary.  TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests.util import C4_ARGS, LOTD_CONFIGS, c4_inputs, load_ref, lotd_inputs, meta_args  # noqa: E402


def main():
    variant, jobs, out = sys.argv[1], json.loads(sys.argv[2]), sys.argv[3]
    ref = load_ref("_lotd", variant=variant)
    assert ref is not None, f"reference build variant {variant!r} not found"
    dev = torch.device("cuda:0")
    res = {}
    for j, job in enumerate(jobs):
        meta = ref.LoDMeta(*(C4_ARGS if job["name"] == "c4" else meta_args(LOTD_CONFIGS[job["name"]])))
        pdtype = torch.float16 if job.get("dtype", "f32") == "f16" else torch.float32
        N = int(job["N"])
        if job["name"] == "c4":
            inp = c4_inputs(N, meta.n_params, seed=int(job["seed"]))
        else:
            inp = lotd_inputs(LOTD_CONFIGS[job["name"]], meta.n_params, N=N, seed=int(job["seed"]), batch_mode=job.get("batch_mode", "inds"))
        x, params = inp["x"].to(dev), inp["params"].to(dev).to(pdtype)
        dL_dy, ddx = inp["dL_dy"].to(dev).to(pdtype), inp["dL_ddLdx"].to(dev)
        bi = None if inp.get("batch_inds") is None else inp["batch_inds"].to(dev)
        kw = dict(batch_inds=bi, batch_offsets=None, batch_data_size=inp.get("batch_data_size") or None, max_level=None)
        y, dy_dx = ref.lod_fwd(meta, x, params, need_input_grad=True, **kw)
        dL_dx, dL_dparam = ref.lod_bwd(meta, dL_dy, x, params, dy_dx, need_input_grad=True, need_param_grad=True, **kw)
        a, b, c = ref.lod_bwd_bwd_input(meta, ddx, dL_dy, x, params, dy_dx, need_dLdinput_ddLdoutput=True, need_dLdinput_dparams=True,
                                        need_dLdinput_dinput=True, **kw)
        E, D = meta.n_encoded_dims, meta.n_dims_to_encode
        outs = dict(y=y, dy_dx=dy_dx.reshape(N, E, D), dL_dx=dL_dx, dL_dparam=dL_dparam, dL_ddLdy=a, dL_dparam2=b, dL_dx2=c)
        want = job.get("keys")
        for k, v in outs.items():
            if want is None or k in want:
                res[f"{j}/{k}"] = v.detach().float().cpu().numpy()
    np.savez(out, **res)


if __name__ == "__main__":
    main()
