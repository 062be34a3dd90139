"""Occupancy value-grid maintenance.  CPU part: the numpy oracle against golden vectors produced by the
reference's own utils.py (tests/golden/make_golden_occ.py).  GPU part: the B200 kernels against the golden vectors and the
oracle -- value grids and cell indices bit-exact."""
import numpy as np
import pytest
import torch

from oracle import occ_oracle as OO
from tests.util import golden, occ_inputs


def _oracle_all(d):
    ema = float(d["ema"])
    o = {"upd_pts": OO.update_pts(d["grid"].copy(), d["pts"], d["vals"], ema),
         "upd_idx": OO.update_idx(d["grid"].copy(), d["gidx"], d["vals"], ema),
         "upd_pts_ema1": OO.update_pts(d["grid"].copy(), d["pts"], d["vals"], 1.0),
         "upd_b_pts": OO.update_pts(d["bgrid"].copy(), d["pts"], d["vals"], ema, bidx=d["bidx"]),
         "upd_b_idx": OO.update_idx(d["bgrid"].copy(), d["bgidx"], d["bvals"], ema),
         "upd_b_pts_nobidx": OO.update_idx(d["bgrid"].copy(), OO.cells_of(d["bpts"], d["bgrid"].shape[1:]), d["bvals"], ema)}
    o["bin_const"] = OO.binarize(o["upd_pts"], d["thre"])
    o["bin_mean"] = OO.binarize(o["upd_pts"], d["thre"], consider_mean=True)
    return o


def _check_updates(got, want):
    for k in ("upd_pts", "upd_idx", "upd_pts_ema1", "upd_b_pts", "upd_b_idx", "upd_b_pts_nobidx", "bin_const"):
        assert np.array_equal(got[k], want[k]), k


def _check_bin_mean(got_mask, grid, g):
    """mean-relative threshold: the fp32 mean may differ by an ulp between summation orders -> only cells within 1e-6 of the
    threshold may differ from the reference."""
    diff = got_mask != g["bin_mean"]
    assert np.all(np.abs(grid[diff] - g["bin_mean_thr"]) < 1e-6)


def test_oracle_vs_golden():
    g = golden("occ_update")
    assert g is not None, "tests/golden/occ_update.npz missing"
    o = _oracle_all(g)
    _check_updates(o, g)
    _check_bin_mean(o["bin_mean"], g["upd_pts"], g)
    res = g["grid"].shape
    p, v = OO.sample_pts(g["vox"], res, g["smp_sparse_off"], g["smp_sparse_vidx"])
    assert np.array_equal(p, g["smp_sparse_pts"]) and np.array_equal(v, g["smp_sparse_vidx"])
    p, v = OO.sample_pts(g["vox"], res, g["smp_dense_off"])
    assert np.array_equal(p, g["smp_dense_pts"]) and np.array_equal(v, g["smp_dense_vidx"])
    # sampled points fall back into their voxel
    assert np.array_equal(OO.cells_of(g["smp_dense_pts"], res), g["vox"][g["smp_dense_vidx"]])


def test_oracle_semantics():
    d = occ_inputs(seed=4)
    before = d["grid"].copy()
    after = OO.update_pts(d["grid"].copy(), d["pts"], d["vals"], 0.5)
    cells = OO.cells_of(d["pts"], before.shape)
    touched = np.zeros(before.shape, bool)
    touched[cells[:, 0], cells[:, 1], cells[:, 2]] = True
    assert np.array_equal(after[~touched], before[~touched])                  # untouched cells are NOT decayed
    assert np.all(after[touched] >= np.float32(0.5) * before[touched])         # touched: at least the decayed old value
    again = OO.update_pts(after.copy(), d["pts"], d["vals"], 1.0)
    assert np.array_equal(again, after)                                        # idempotent without decay


# ------------------------------------------------------------------------------------------------------------------
def _gpu_all(d, dev):
    from nr3d_lib_b200 import occgrid as G
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ema = float(d["ema"])
    o = {}
    g = t(d["grid"]); G.update_occ_val_grid_(g, t(d["pts"]), t(d["vals"]), ema_decay=ema); o["upd_pts"] = g
    g = t(d["grid"]); G.update_occ_val_grid_idx_(g, t(d["gidx"]), t(d["vals"]), ema_decay=ema); o["upd_idx"] = g
    g = t(d["grid"]); G.update_occ_val_grid_(g, t(d["pts"]), t(d["vals"]), ema_decay=1.0); o["upd_pts_ema1"] = g
    g = t(d["bgrid"]); G.update_batched_occ_val_grid_(g, t(d["pts"]), t(d["bidx"]), t(d["vals"]), ema_decay=ema); o["upd_b_pts"] = g
    g = t(d["bgrid"]); G.update_batched_occ_val_grid_idx_(g, None, t(d["bgidx"]), t(d["bvals"]), ema_decay=ema); o["upd_b_idx"] = g
    g = t(d["bgrid"]); G.update_batched_occ_val_grid_(g, t(d["bpts"]), None, t(d["bvals"]), ema_decay=ema); o["upd_b_pts_nobidx"] = g
    o["bin_const"] = G.binarize(o["upd_pts"], float(d["thre"]))
    o["bin_mean"] = G.binarize(o["upd_pts"], float(d["thre"]), consider_mean=True)
    return {k: v.cpu().numpy() for k, v in o.items()}


@pytest.mark.gpu
def test_gpu_vs_golden(dev):
    from nr3d_lib_b200 import occgrid as G
    g = golden("occ_update")
    got = _gpu_all(g, dev)
    _check_updates(got, g)
    _check_bin_mean(got["bin_mean"], g["upd_pts"], g)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = g["grid"].shape
    p, v = G.sample_pts_from_offsets(t(g["vox"]), res, t(g["smp_sparse_off"]), t(g["smp_sparse_vidx"]))
    assert np.array_equal(p.cpu().numpy(), g["smp_sparse_pts"]) and np.array_equal(v.cpu().numpy(), g["smp_sparse_vidx"])
    p, v = G.sample_pts_from_offsets(t(g["vox"]), res, t(g["smp_dense_off"]))
    assert np.array_equal(p.cpu().numpy(), g["smp_dense_pts"]) and np.array_equal(v.cpu().numpy(), g["smp_dense_vidx"])


@pytest.mark.gpu
@pytest.mark.parametrize("res,N,B,seed", [((8, 8, 8), 500, 2, 1), ((33, 17, 9), 20000, 4, 2), ((64, 64, 64), 200000, 2, 3)])
def test_gpu_vs_oracle(res, N, B, seed, dev):
    d = occ_inputs(res=res, N=N, B=B, seed=seed)
    got, want = _gpu_all(d, dev), _oracle_all(d)
    _check_updates(got, want)
    thr = min(np.float32(np.float32(want["upd_pts"].astype(np.float64).mean()) - np.float32(1e-5)), d["thre"])
    diff = got["bin_mean"] != want["bin_mean"]
    assert np.all(np.abs(want["upd_pts"][diff] - thr) < 1e-6)


@pytest.mark.gpu
def test_gpu_fused_update_query_and_class(dev):
    from nr3d_lib_b200 import occgrid as G
    d = occ_inputs(res=(32, 32, 32), N=50000, B=2, seed=6)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for consider_mean in (False, True):
        grid, occ = t(d["grid"]), torch.zeros(d["grid"].shape, dtype=torch.bool, device=dev)
        G.update_and_binarize_(grid, occ, pts=t(d["pts"]), occ_val=t(d["vals"]), ema_decay=0.9, occ_threshold=0.55, consider_mean=consider_mean,
                               extra=(t(d["gidx"][:100]), None, t(d["vals"][:100] + 1.0)))
        want = OO.update_idx(d["grid"].copy(), np.concatenate([OO.cells_of(d["pts"], d["grid"].shape), d["gidx"][:100]]),
                             np.concatenate([d["vals"], d["vals"][:100] + 1.0]), 0.9)
        assert np.array_equal(grid.cpu().numpy(), want)
        wocc = OO.binarize(want, 0.55, consider_mean)
        diff = occ.cpu().numpy() != wocc
        assert diff.sum() == 0 or consider_mean
        # the scratch is clean again: a second update with no new information and ema 1 is a no-op
        G.update_occ_val_grid_(grid, t(d["pts"][:1]), t(np.array([-1e30], np.float32)), ema_decay=1.0)
        assert np.array_equal(grid.cpu().numpy(), want)
        q = G.query_occ_grid(occ, t(d["pts"]))
        assert np.array_equal(q.cpu().numpy(), OO.query(occ.cpu().numpy(), d["pts"]))
    bocc = t(d["bgrid"]) > 0.5
    q = G.query_occ_grid(bocc, t(d["pts"]), t(d["bidx"]))
    assert np.array_equal(q.cpu().numpy(), OO.query(bocc.cpu().numpy(), d["pts"], d["bidx"]))
    # the maintainer class on the reference's smoke scenario (occgrid/unit_test.py:7-35): sphere SDF of radius 0.5
    torch.manual_seed(0)
    occ = G.OccGridEma([32, 32, 32], occ_val_fn=lambda sdf: 1.0 - sdf.abs(), occ_thre=0.9, ema_decay=0.95, n_steps_between_update=4,
                       n_steps_warmup=8, should_collect_samples=True, device=dev)
    sdf = lambda x: x.norm(dim=-1) - 0.5
    occ.init_from_net(sdf, num_steps=4, num_pts=2 ** 16)
    for it in range(1, 17):
        p = torch.rand([4096, 3], device=dev) * 2 - 1
        occ.collect_samples(p, sdf(p))
        occ.step(it, sdf, num_steps=2, num_pts=2 ** 15)
    frac = occ.occ_grid.float().mean().item()
    assert 0.02 < frac < 0.5                                       # a thin shell around r = 0.5
    centers = occ.sample_pts_in_occupied(2000)
    assert (centers.norm(dim=-1) - 0.5).abs().max().item() < 0.25
    assert occ.query(centers).all()
    with pytest.raises(RuntimeError):
        G.update_occ_val_grid_(torch.zeros(4, 4, 4), torch.zeros(1, 3), torch.zeros(1))    # CPU tensors: no fallback
