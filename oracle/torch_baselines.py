"""The reference's PURE-PYTORCH paths, timed on the host cores as reported baselines (BASELINE.json north_star / configs[0]) -- TEST AND
BENCH INFRASTRUCTURE ONLY (bench.py's cpu_baseline leg; nothing under nr3d_lib_b200/ imports this).

  * `param_interpolate`  nr3d_lib/models/grid_encodings/lotd/lotd_helpers.py:274-346 -- a Dense LoTD level evaluated with F.grid_sample
    (align_corners=True, coordinates scaled by (R-2)/(R-1), zyx flip).  configs[0]: 1 level, R = 32, F = 4, 4096 random points, fwd + bwd.
  * `ray_alpha_to_vw`    nr3d_lib/graphics/nerf/nerf_utils.py:98-110 -- front-to-back compositing weights of a DENSE [rays, samples] batch:
    roll + torch.cumprod.

The functions are taken from the reference's own files when a copy is reachable (/root/reference here, the verbatim staging copy
oracle/_ref/pyref on the GPU box; `kind: "reference"`); otherwise the restatements below run (`kind: "port"`).  tests/test_oracle_cpu.py
checks the restatements against the reference files and the Dense path against the float64 LoTD oracle.
"""
import importlib.util
import os
import sys
import time
import types

import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
_ROOTS = ("/root/reference/nr3d_lib", os.path.join(HERE, "_ref", "pyref", "nr3d_lib"))


# ---- restatements (used only when the reference files are not reachable) ----
def param_interpolate_port(param: torch.Tensor, rel_x: torch.Tensor, res: int) -> torch.Tensor:
    """3-D case of lotd_helpers.py:327-346: param [B,R,R,R,M], rel_x [B,...,3] in [-1,1] -> [B,...,M]."""
    data_shape = rel_x.shape[1:-1]
    B, N = param.shape[0], int(torch.tensor(data_shape).prod()) if len(data_shape) else 1
    scale = (res - 2.0) / (res - 1.0)
    g = (rel_x * scale)[..., [2, 1, 0]].view(B, 1, 1, N, 3)          # LoTD tables are z fastest, grid_sample wants x fastest
    ret = F.grid_sample(param.permute(0, 4, 1, 2, 3), g, align_corners=True, padding_mode="zeros")
    return ret.view(B, -1, N).permute(0, 2, 1).unflatten(1, data_shape)


def ray_alpha_to_vw_port(alpha: torch.Tensor) -> torch.Tensor:
    """nerf_utils.py:98-110."""
    shifted = torch.roll((1 + 1e-10) - alpha, 1, dims=-1)
    shifted[..., 0] = 1
    return alpha * torch.cumprod(shifted, dim=-1)


def _load_reference_functions():
    """(param_interpolate, ray_alpha_to_vw, where) from the reference's own files, or (None, None, None).  The two files import a long
    chain of unrelated modules at the top; those names are provided as empty stand-ins (the two functions only use torch)."""
    for root in _ROOTS:
        helpers = os.path.join(root, "models", "grid_encodings", "lotd", "lotd_helpers.py")
        nerf = os.path.join(root, "graphics", "nerf", "nerf_utils.py")
        if not (os.path.isfile(helpers) and os.path.isfile(nerf)):
            continue
        saved = dict(sys.modules)
        try:
            def stub(name, **attrs):
                m = types.ModuleType(name)
                m.__path__ = []
                for k, v in attrs.items():
                    setattr(m, k, v)
                sys.modules[name] = m
            dummy = lambda *a, **k: None   # noqa: E731
            stub("nr3d_lib")
            stub("nr3d_lib.utils", tensor_statistics=dummy)
            stub("nr3d_lib.models")
            stub("nr3d_lib.models.utils", clip_norm_=dummy)
            stub("nr3d_lib.models.grid_encodings")
            stub("nr3d_lib.models.grid_encodings.utils", gridsample1d_by2d=dummy)
            stub("nr3d_lib.models.grid_encodings.lotd")
            stub("nr3d_lib.models.grid_encodings.lotd.lotd", LoDType=type("LoDType", (), {}))
            stub("nr3d_lib.graphics")
            stub("nr3d_lib.graphics.pack_ops", packed_cumsum=dummy, packed_cumprod=dummy, packed_alpha_to_vw=dummy, packed_volume_render_compression=dummy)
            mods = []
            for name, path in (("_nr3d_ref_lotd_helpers", helpers), ("_nr3d_ref_nerf_utils", nerf)):
                spec = importlib.util.spec_from_file_location(name, path)
                mod = importlib.util.module_from_spec(spec)
                spec.loader.exec_module(mod)
                mods.append(mod)
            return mods[0].param_interpolate, mods[1].ray_alpha_to_vw, root
        except Exception:
            continue
        finally:
            for k in list(sys.modules):
                if k not in saved:
                    del sys.modules[k]
            sys.modules.update(saved)
    return None, None, None


def functions():
    pi, vw, where = _load_reference_functions()
    if pi is None:
        return param_interpolate_port, ray_alpha_to_vw_port, "port", "restatement in oracle/torch_baselines.py"
    return pi, vw, "reference", where


def _cores():
    return max(1, len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1))


def time_config0(iters=20, warmup=3, seed=42):
    """BASELINE.json configs[0]: 1-level Dense LoTD 32^3, F = 4, 4096 random points, forward + backward through F.grid_sample on the CPU."""
    pi, _, kind, where = functions()
    cores = _cores()
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    R, M, N = 32, 4, 4096
    param = (torch.randn(1, R, R, R, M, generator=g) * 1e-2).requires_grad_(True)
    x = torch.rand(1, N, 3, generator=g) * 2 - 1
    w = torch.randn(1, N, M, generator=g)
    ts = []
    for it in range(warmup + iters):
        param.grad = None
        t0 = time.perf_counter()
        y = pi(param, x, R)
        (y * w).sum().backward()
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = sorted(ts)[len(ts) // 2]
    return dict(value=N / sec / 1e6, unit="Msamples/s", ms=sec * 1e3, cores=cores, kind=kind, source=where,
                sample=f"configs[0]: 1-level Dense 32^3 F=4, {N} points, fwd+bwd via F.grid_sample (lotd_helpers.py:327-346), median of {iters}")


def time_composite(rays=65536, samples=128, iters=5, warmup=1, seed=42):
    """Dense-batch compositing weights (roll + cumprod, nerf_utils.py:98-110) + weighted depth sum, forward + backward, on the CPU."""
    _, vw, kind, where = functions()
    cores = _cores()
    torch.set_num_threads(cores)
    g = torch.Generator().manual_seed(seed)
    alpha = (torch.rand(rays, samples, generator=g) ** 2 * 0.2).requires_grad_(True)
    depth = torch.rand(rays, samples, generator=g)
    ts = []
    for it in range(warmup + iters):
        alpha.grad = None
        t0 = time.perf_counter()
        wts = vw(alpha)
        (wts * depth).sum(-1).sum().backward()
        if it >= warmup:
            ts.append(time.perf_counter() - t0)
    sec = sorted(ts)[len(ts) // 2]
    return dict(value=rays / sec / 1e6, unit="Mrays/s", ms=sec * 1e3, samples_per_s=rays * samples / sec, cores=cores, kind=kind, source=where,
                sample=f"{rays} rays x {samples} samples dense batch, ray_alpha_to_vw (roll + torch.cumprod, nerf_utils.py:98-110) + depth sum, fwd+bwd, median of {iters}")
