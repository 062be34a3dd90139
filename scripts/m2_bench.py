"""M2 of BASELINE.json (configs[2] / configs[4]): occupancy march (128^3 grid) + 16-level NGP LoTD + packed alpha-composite,
forward + backward to the LoTD parameters, 1024^2 rays per GPU in chunks; reports rays/s and samples/ray.

    python scripts/m2_bench.py [--rays 1048576] [--chunk 262144] [--grid random|shell] [--steps 3]
    python -m torch.distributed.run --nproc-per-node N ... scripts/m2_bench.py        # rays sharded, one all-reduce per step
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import ngp_cfg  # noqa: E402
from nr3d_lib_b200 import dist as ndist  # noqa: E402
from nr3d_lib_b200.lotd import LoTD  # noqa: E402
from nr3d_lib_b200.pipeline import march_encode_composite  # noqa: E402


def make_rays(n, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    o = torch.randn(n, 3, device=dev, generator=g)
    o = 4.0 * o / o.norm(dim=-1, keepdim=True)
    tgt = torch.rand(n, 3, device=dev, generator=g) - 0.5
    d = tgt - o
    d = d / d.norm(dim=-1, keepdim=True)
    t1, t2 = (-1.0 - o) / d, (1.0 - o) / d
    near = torch.minimum(t1, t2).amax(-1).clamp_min(0.0)
    far = torch.maximum(t1, t2).amin(-1)
    far = torch.where(far <= near, near, far)
    return o.contiguous(), d.contiguous(), near.contiguous(), far.contiguous()


def run_m2(dev, rank, world, rays=1024 * 1024, chunk=262144, grid_kind="random", steps=3, warmup=1, sort_points=True, fuse_head=True):
    """Times `steps` M2 steps (after `warmup`) on this rank's ray shard; returns the result dict (max over ranks inside)."""
    _, res, feats, types, T, _ = ngp_cfg()
    enc = LoTD(3, res, feats, types, hashmap_size=T, dtype=torch.float)
    enc.meta.c_sort_points = sort_points
    g = torch.Generator(device=dev).manual_seed(42)
    params = ((torch.rand(enc.n_params, device=dev, generator=g) * 2 - 1) * 1e-2).requires_grad_(True)
    R = 128
    if grid_kind == "random":
        grid = torch.rand(R, R, R, device=dev, generator=g) > 0.5
    else:
        ax = torch.linspace(-1, 1, R, device=dev)
        pts = torch.stack(torch.meshgrid(ax, ax, ax, indexing="ij"), -1)
        grid = (pts.norm(dim=-1) - 0.6).abs() < 0.05
    ray_set = make_rays(rays, dev, 1000 + rank)
    n_samples = 0

    def step():
        nonlocal n_samples
        params.grad = None
        n_samples = 0
        for b in range(0, rays, chunk):
            o, d, near, far = (r[b:b + chunk] for r in ray_set)
            out = march_encode_composite(enc, params, grid, o, d, near, far, step_size=0.01, max_steps=512, gain=2.0, fuse_head=fuse_head)
            if out.depth is None:
                continue
            n_samples += out.weights.numel()
            ((out.depth ** 2).sum() + out.acc.sum()).backward()
        if params.grad is not None:
            ndist.allreduce_param_grads(params.grad, world)

    for _ in range(warmup):
        step()
    ndist.barrier()
    torch.cuda.synchronize(dev)
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    evs[0].record()
    for k in range(steps):
        step()
        evs[k + 1].record()
    torch.cuda.synchronize(dev)
    e0, e1 = evs[0], evs[-1]
    step_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(steps)]      # this rank's steps one by one (spread = how noisy the box is)
    ms = ndist.max_over_ranks(e0.elapsed_time(e1), dev) / steps
    ndist.barrier()
    total_samples = ndist.sum_over_ranks(float(n_samples), dev)
    # Roofline of the whole step against the measured HBM copy bandwidth.  Algorithmic bytes per sample (fp32, fused density head; the 48.5 MB
    # table itself is L2 resident, its corner traffic is counted like SURVEY.md 8d counts it for M1):
    #   march 16 (t_start, t_end, ray id, voxel id written)            march_samples 28 (3 reads, position + delta written)
    #   point sort 48 (x twice, rank, 16-byte record)                   encode + head forward 1052 (record, 1024 corner bytes, delta, sigma, alpha)
    #   composite forward 8, per-pack sums forward 8                    per-pack sums backward 8, composite backward 16
    #   encode + head backward 1080 (x fingerprint 12, record 16, 4 scalars, 1024 scatter bytes), table zero-init / reduce 3
    bytes_per_sample = 16 + 28 + 48 + 1052 + 8 + 8 + 8 + 16 + 1080 + 3
    peak = 6553.9
    try:
        peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    achieved = total_samples / world * bytes_per_sample / (ms * 1e-3) / 1e9      # per GPU
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None,
                "algorithmic_bytes_per_sample": bytes_per_sample,
                "note": "whole M2 step per GPU; 93 % of the bytes are LoTD corner gathers / scatters that hit the L2-resident table, so -- as for M1 -- "
                        "the binding resource is the L1 line rate (forward) and the L2 reduction rate (backward), not DRAM"}
    return {"roofline": roofline, "metric": "full march+encode+composite fwd+bwd Mrays/s", "value": world * rays / ms / 1e3, "unit": "Mrays/s",
            "n_gpus": world, "ms_per_step": ms, "step_ms": [round(t, 3) for t in step_ms], "rays_per_gpu": rays, "samples_per_ray": total_samples / (world * rays),
            "Msamples_per_s": total_samples / ms / 1e3, "grid": grid_kind, "chunk": chunk, "steps": steps, "warmup": warmup,
            "sort_points": sort_points, "fuse_head": bool(fuse_head and sort_points), "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 2 ** 30,
            "workload": ("configs[4] (4096^2 rays over 8 GPUs)" if world * rays == 4096 * 4096 else "configs[2] shape (1024^2 rays per GPU)") +
                        ": occ_grid 128^3 march (step 0.01, <=512 steps) + 16L NGP LoTD + softplus density head + "
                        "packed alpha-composite + per-ray sums, forward + backward to the LoTD parameters, rays sharded over the GPUs, "
                        "one NCCL all-reduce of dL/dparams per step"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rays", type=int, default=1024 * 1024)
    ap.add_argument("--chunk", type=int, default=262144)
    ap.add_argument("--grid", default="random", choices=["random", "shell"])
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=1)
    ap.add_argument("--no-sort", action="store_true")
    ap.add_argument("--no-fuse-head", action="store_true", help="keep the [S, 32] features in HBM between the encoder and the density head")
    args = ap.parse_args()
    rank, world, local = ndist.init_from_env("nccl")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if os.environ.get("NR3D_M2_TWO_LEVEL_OFF"):      # A/B: force the one-level point sort for every size
        from nr3d_lib_b200 import _lib
        _lib.check(_lib.get_lib().nr3d_lotd_sort_set_two_level_min(1 << 40))
    out = run_m2(dev, rank, world, args.rays, args.chunk, args.grid, args.steps, args.warmup, not args.no_sort, not args.no_fuse_head)
    ndist.shutdown()
    if rank == 0:
        print(json.dumps(out))


if __name__ == "__main__":
    main()
