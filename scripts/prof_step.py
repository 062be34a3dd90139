"""Short profiling target: a few fwd+bwd steps of the headline config (16-level NGP LoTD, 4 Mi points, fp32 tables) for ncu.

    ncu --set full --clock-control none --import-source on -k regex:lotd_pair -s 4 -c 2 -o gpurun_out/prof python scripts/prof_step.py [steps] [points] [half]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from bench import ngp_cfg  # noqa: E402
from nr3d_lib_b200.bindings import _lotd  # noqa: E402


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4 * 1024 * 1024
    half = len(sys.argv) > 3 and sys.argv[3] == "half"
    dev = torch.device("cuda:0")
    meta = _lotd.LoDMeta(*ngp_cfg())
    torch.manual_seed(42)
    xs = [torch.rand(N, 3, device=dev).clamp(1e-6, 1 - 1e-6) for _ in range(2)]        # alternating point sets: every forward sorts
    dt = torch.float16 if half else torch.float32
    params = ((torch.rand(meta.n_params, device=dev) * 2 - 1) * 1e-4).to(dt)
    dL_dy = (torch.randn(N, meta.n_encoded_dims, device=dev) * 1e-4).to(dt)
    for k in range(steps):
        _lotd.lod_fwd(meta, xs[k & 1], params, need_input_grad=False)
        _lotd.lod_bwd(meta, dL_dy, xs[k & 1], params, None, need_input_grad=False, need_param_grad=True)
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
