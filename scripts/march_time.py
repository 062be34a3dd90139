import os, sys
sys.path.insert(0, os.getcwd())
import torch
from scripts.m2_bench import make_rays
from scripts.quick_bench import timeit
from nr3d_lib_b200.bindings import _occ_grid as og
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(42)
grid = torch.rand(128, 128, 128, device=dev, generator=g) > 0.5
o, d, near, far = make_rays(262144, dev, 1000)
roi = torch.tensor([-1., -1, -1, 1, 1, 1], device=dev)
f = lambda: og.ray_marching(o, d, near, far, roi, grid, og.ContractionType.AABB, 0.01, 1e10, 0.0, 512, True)
r = f(); print("samples", r[1].shape[0], "ray_marching (count + scan + fill):", timeit(f), "ms")
