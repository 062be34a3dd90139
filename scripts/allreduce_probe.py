"""Time the step's one collective on its own: all-reduce (sum) of the 12 131 648-float dL/dparams (48.5 MB fp32), NCCL defaults vs
whatever NCCL_* environment the launcher sets, plus reduce-scatter and the torch symmetric-memory all-reduces when this torch has them.

    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/allreduce_probe.py
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nr3d_lib_b200 import dist as ndist  # noqa: E402


def main():
    rank, world, local = ndist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n = 12131648
    g = torch.randn(n, device=dev)
    out = {"world": world, "env": {k: v for k, v in os.environ.items() if k.startswith("NCCL_")}}

    def timeit(fn, iters=20):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(dev)
        dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return ndist.max_over_ranks(a.elapsed_time(b) / iters, dev)

    out["all_reduce_ms"] = timeit(lambda: dist.all_reduce(g))
    per = n // world
    slice_out = torch.empty(per, device=dev)
    out["reduce_scatter_ms"] = timeit(lambda: dist.reduce_scatter_tensor(slice_out, g[: per * world]))
    half = g.half()
    out["all_reduce_fp16_ms"] = timeit(lambda: dist.all_reduce(half))
    try:
        import torch.distributed._symmetric_memory as symm
        t = symm.empty(n, dtype=torch.float32, device=dev)
        t.copy_(g)
        hdl = symm.rendezvous(t, dist.group.WORLD.group_name)
        for name in ("multimem_all_reduce_", "two_shot_all_reduce_", "one_shot_all_reduce"):
            op = getattr(torch.ops.symm_mem, name, None)
            if op is None:
                out["symm_" + name] = "absent"
                continue
            try:
                out["symm_" + name + "_ms"] = timeit(lambda: op(t, "sum", dist.group.WORLD.group_name))
            except Exception as e:
                out["symm_" + name] = "failed: " + str(e)[:160]
    except Exception as e:
        out["symm"] = "unavailable: " + str(e)[:200]
    if rank == 0:
        print(json.dumps(out))
    ndist.shutdown()


if __name__ == "__main__":
    main()
