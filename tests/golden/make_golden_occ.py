#!/usr/bin/env python
"""Golden vectors for the occupancy value-grid maintenance: runs the REFERENCE'S OWN
nr3d_lib/models/accelerations/occgrid/utils.py, imported from /root/reference, on CPU in the build container:

    python tests/golden/make_golden_occ.py            # writes tests/golden/occ_update.npz

The file only needs torch plus three imports that are absent here and never reached by the functions under test:
`torch_scatter.scatter_max` (not installed; replaced by torch's scatter_reduce_(amax, include_self=True), which is its
documented behaviour when `out` is given), `nr3d_lib.models.annealers` and `nr3d_lib.maths` (used by sdf_to_occ_val only).
torch.rand / torch.randint are wrapped to record the random tensors the reference draws, so that the kernels can be
checked on identical offsets.  Nothing from this repository's kernels or oracle takes part in producing the numbers.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests.util import occ_inputs  # noqa: E402


def load_reference_utils():
    def scatter_max(src, index, dim=-1, out=None):
        out.scatter_reduce_(dim if dim >= 0 else src.dim() + dim, index, src, reduce="amax", include_self=True)
        return out, None
    ts = types.ModuleType("torch_scatter"); ts.scatter_max = scatter_max
    sys.modules["torch_scatter"] = ts
    for name, attrs in (("nr3d_lib", {}), ("nr3d_lib.models", {}), ("nr3d_lib.models.annealers", {"get_anneal_val": None}),
                        ("nr3d_lib.maths", {"normalized_logistic_density": None})):
        m = types.ModuleType(name); m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules.setdefault(name, m)
    path = "/root/reference/nr3d_lib/models/accelerations/occgrid/utils.py"
    spec = importlib.util.spec_from_file_location("ref_occgrid_utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    U = load_reference_utils()
    d = occ_inputs()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    out = dict(d)
    g = t(d["grid"]).clone(); U.update_occ_val_grid_(g, t(d["pts"]), t(d["vals"]), ema_decay=float(d["ema"])); out["upd_pts"] = g.numpy()
    g = t(d["grid"]).clone(); U.update_occ_val_grid_idx_(g, t(d["gidx"]), t(d["vals"]), ema_decay=float(d["ema"])); out["upd_idx"] = g.numpy()
    g = t(d["grid"]).clone(); U.update_occ_val_grid_(g, t(d["pts"]), t(d["vals"]), ema_decay=1.0); out["upd_pts_ema1"] = g.numpy()
    g = t(d["bgrid"]).clone(); U.update_batched_occ_val_grid_(g, t(d["pts"]), t(d["bidx"]), t(d["vals"]), ema_decay=float(d["ema"])); out["upd_b_pts"] = g.numpy()
    g = t(d["bgrid"]).clone(); U.update_batched_occ_val_grid_idx_(g, None, t(d["bgidx"]), t(d["bvals"]), ema_decay=float(d["ema"])); out["upd_b_idx"] = g.numpy()
    g = t(d["bgrid"]).clone(); U.update_batched_occ_val_grid_(g, t(d["bpts"]), None, t(d["bvals"]), ema_decay=float(d["ema"])); out["upd_b_pts_nobidx"] = g.numpy()
    out["bin_const"] = U.binarize(t(out["upd_pts"]), float(d["thre"])).numpy()
    out["bin_mean"] = U.binarize(t(out["upd_pts"]), float(d["thre"]), consider_mean=True).numpy()
    out["bin_mean_thr"] = np.float32((t(out["upd_pts"]).mean() - 1e-5).clamp_max_(float(d["thre"])).item())
    # sample_pts_in_voxels: record the random draws
    rec = {}
    real_rand, real_randint = torch.rand, torch.randint
    def rand(*a, **k):
        r = real_rand(*a, **k); rec["offsets"] = r; return r
    def randint(*a, **k):
        r = real_randint(*a, **k); rec["vidx"] = r; return r
    torch.rand, torch.randint = rand, randint
    try:
        torch.manual_seed(7)
        res = torch.tensor(d["grid"].shape)
        vox = t(d["vox"])
        pts, vidx = U.sample_pts_in_voxels(vox, int(1.5 * vox.shape[0]), res)                 # sparse branch: random voxel per point
        out.update(smp_sparse_pts=pts.numpy(), smp_sparse_vidx=vidx.numpy(), smp_sparse_off=rec["offsets"].numpy())
        pts, vidx = U.sample_pts_in_voxels(vox, int(3.3 * vox.shape[0]), res)                 # dense branch: n_per_vox per voxel
        out.update(smp_dense_pts=pts.numpy(), smp_dense_vidx=vidx.numpy(), smp_dense_off=rec["offsets"].numpy())
    finally:
        torch.rand, torch.randint = real_rand, real_randint
    path = os.path.join(ROOT, "tests", "golden", "occ_update.npz")
    np.savez_compressed(path, **out)
    print("[golden] occ_update: " + ", ".join(f"{k}{tuple(np.shape(v))}" for k, v in out.items()))


if __name__ == "__main__":
    main()
