"""LoTD over a forest of blocks -- SURVEY.md section 8f row n4, reference csrc/lotd/include/lotd/lotd_forest.h.

Pinning: the reference's forest FORWARD cannot run at this commit (`output` / `dy_dx` are never allocated on the forest branch of
lod_fwd_common, lotd_torch_api.cu:300-362); its backward passes do run, so the goldens (tests/golden/forest_lotd_*.npz, made by
tests/golden/make_golden.py:make_forest_lotd from the reference's own CUDA build) hold dL/dparam, d(dL/dx)/dparam and d(dL/dx)/dx.
dL/dparam fixes the corner weights, indices and the cross-block remap of the forward (y is linear in each table entry with exactly
those weights); y and dy_dx themselves are checked against the float64 oracle, the adjoint identity and finite differences.
"""
import numpy as np
import pytest
import torch

from tests.util import FOREST_LOTD_CONFIGS, forest_lotd_inputs, golden, load_ref, meta_args, rel_err

LINEAR_TYPES = ("Dense", "Hash")


def _oracle_forest(f, continuity=True):
    from oracle import lotd_oracle as O
    return O.OracleForest(f["octree"], f["exsum"], f["block_ks"], int(f["level"]), int(f["level_poffset"]), continuity)


def _golden_inputs(g):
    t = lambda k: torch.from_numpy(g[k])
    return dict(x=t("x"), params=t("params"), dL_dy=t("dL_dy"), dL_ddLdx=t("dL_ddLdx"), batch_inds=t("batch_inds"),
                forest=dict(octree=g["octree"], exsum=g["exsum"], block_ks=g["block_ks"], level=int(g["level"]), level_poffset=int(g["level_poffset"])))


def _param_masks(om, cfg, n_blocks):
    """Entries of the second-order dL/dparam that the reference build can check: ALL of them -- unlike the reference's generic
    single-block kernels (DESIGN.md "reference build defect"), its forest kernels agree with the oracle on every level type
    (measured on B200: <= 2.5e-7 relative on Dense / VM / NPlaneMul / CP / Hash)."""
    return torch.ones(om.n_params * n_blocks, dtype=torch.bool)


GOLDENS = [("mixed", "f32"), ("mixed", "f16"), ("smooth", "f32"), ("hash_f4", "f32")]


@pytest.mark.parametrize("name,tag", GOLDENS)
def test_forest_lotd_oracle_vs_golden(name, tag):
    from oracle import lotd_oracle as O
    g = golden(f"forest_lotd_{name}_{tag}")
    if g is None:
        pytest.skip("golden fixture not generated yet")
    cfg = FOREST_LOTD_CONFIGS[name]
    om = O.OracleMeta(*meta_args(cfg))
    inp = _golden_inputs(g)
    F = _oracle_forest(inp["forest"])
    kw = dict(batch_inds=inp["batch_inds"], forest=F)
    tol = 2e-5 if tag == "f32" else 3e-2
    _, gp = O.bwd(om, inp["dL_dy"].float(), inp["x"], inp["params"].float(), **kw)
    assert rel_err(gp, g["dL_dparam"].astype(np.float32)) < tol
    _, gp_ml = O.bwd(om, inp["dL_dy"].float(), inp["x"], inp["params"].float(), max_level=1, **kw)
    assert rel_err(gp_ml, g["dL_dparam_maxlevel1"].astype(np.float32)) < tol
    _, gp2, gx2 = O.bwd_bwd_input(om, inp["dL_ddLdx"], inp["dL_dy"].float(), inp["x"], inp["params"].float(), **kw)
    assert rel_err(gx2, g["dL_dx2"]) < (1e-4 if tag == "f32" else 3e-2)
    ok2 = _param_masks(om, cfg, g["block_ks"].shape[0])
    if ok2.any():
        assert rel_err(gp2[ok2], torch.from_numpy(g["dL_dparam2"].astype(np.float32))[ok2]) < tol


def test_forest_oracle_identities():
    """Structure of the oracle itself: adjoint identity on the linear level types, continuity across a shared face, zero
    contribution of absent neighbours / disabled continuity, and block-index -1 rows."""
    from oracle import lotd_oracle as O
    cfg = FOREST_LOTD_CONFIGS["hash_f4"]
    om = O.OracleMeta(*meta_args(cfg))
    inp = forest_lotd_inputs(cfg, om.n_params, N=300, seed=3)
    F = _oracle_forest(inp["forest"])
    kw = dict(batch_inds=inp["batch_inds"], forest=F)
    y = O.encode(om, inp["x"], inp["params"], **kw)
    _, gp = O.bwd(om, inp["dL_dy"], inp["x"], inp["params"], **kw)
    assert abs(float((y * inp["dL_dy"].double()).sum() - (inp["params"].double() * gp).sum())) < 1e-9      # <y, g> == <params, dL/dparam>
    assert float(y[inp["batch_inds"] < 0].abs().max()) == 0.0
    y_nc = O.encode(om, inp["x"], inp["params"], batch_inds=inp["batch_inds"], forest=_oracle_forest(inp["forest"], continuity=False))
    interior = ((inp["x"] > 0.2) & (inp["x"] < 0.8)).all(dim=1) & (inp["batch_inds"] >= 0)
    assert interior.any() and torch.equal(y[interior], y_nc[interior]) and not torch.equal(y, y_nc)
    ks = inp["forest"]["block_ks"].astype(np.int64)
    lut = {tuple(k): i for i, k in enumerate(ks.tolist())}
    pair = next(((i, lut[(k[0] + 1, k[1], k[2])]) for i, k in enumerate(ks.tolist()) if (k[0] + 1, k[1], k[2]) in lut), None)
    assert pair is not None
    yz = torch.rand(40, 2, generator=torch.Generator().manual_seed(0))
    xa, xb = torch.cat([torch.full((40, 1), 1.0 - 1e-7), yz], 1), torch.cat([torch.full((40, 1), 1e-7), yz], 1)
    ya = O.encode(om, xa, inp["params"], batch_inds=torch.full((40,), pair[0]), forest=F)
    yb = O.encode(om, xb, inp["params"], batch_inds=torch.full((40,), pair[1]), forest=F)
    assert rel_err(ya, yb) < 1e-5                                                                            # the field is continuous across the face


# ------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------
def _forest_meta(cls, f, dev, continuity=True):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    m = cls()
    m.octree, m.exsum, m.block_ks = t(f["octree"]), t(f["exsum"]), t(f["block_ks"])
    m.n_trees, m.level, m.level_poffset = int(f["block_ks"].shape[0]), int(f["level"]), int(f["level_poffset"])
    m.resolution = [1 << int(f["level"])] * 3
    m.world_origin, m.world_block_size = [0.0, 0.0, 0.0], [1.0, 1.0, 1.0]
    m.continuity_enabled = continuity
    return m


def _run_all(be, metas, inp, dev, pdtype, fwd=True, dydx=None):
    x, params, dL_dy = inp["x"].to(dev), inp["params"].to(dev).to(pdtype), inp["dL_dy"].to(dev).to(pdtype)
    ddx, bi = inp["dL_ddLdx"].to(dev), inp["batch_inds"].to(dev)
    out = {}
    if fwd:
        out["y"], dydx = be.lod_fwd(metas, x, params, bi, None, None, None, True)
        y0, none = be.lod_fwd(metas, x, params, bi, None, None, None, False)
        assert none is None and torch.equal(y0, out["y"])
        out["dy_dx"] = dydx.view(x.shape[0], -1, 3)
        out["y_maxlevel1"], _ = be.lod_fwd(metas, x, params, bi, None, None, 1, False)
    out["dL_dx"], out["dL_dparam"] = be.lod_bwd(metas, dL_dy, x, params, dydx, bi, None, None, None, dydx is not None, True)
    out["dL_ddLdy"], out["dL_dparam2"], out["dL_dx2"] = be.lod_bwd_bwd_input(metas, ddx, dL_dy, x, params, dydx, bi, None, None, None,
                                                                             dydx is not None, True, True)
    _, out["dL_dparam_maxlevel1"] = be.lod_bwd(metas, dL_dy, x, params, None, bi, None, None, 1, False, True)
    return out


def _oracle_all(om, inp, F, half):
    from oracle import lotd_oracle as O
    p = inp["params"].half().float() if half else inp["params"]
    gy = inp["dL_dy"].half().float() if half else inp["dL_dy"]
    kw = dict(batch_inds=inp["batch_inds"], forest=F)
    y, dydx = O.fwd_dydx(om, inp["x"], p, **kw)
    gx, gp = O.bwd(om, gy, inp["x"], p, **kw)
    a, b, c = O.bwd_bwd_input(om, inp["dL_ddLdx"], gy, inp["x"], p, **kw)
    return dict(y=y, dy_dx=dydx, dL_dx=gx, dL_dparam=gp, dL_ddLdy=a, dL_dparam2=b, dL_dx2=c,
                y_maxlevel1=O.encode(om, inp["x"], p, max_level=1, **kw), dL_dparam_maxlevel1=O.bwd(om, gy, inp["x"], p, max_level=1, **kw)[1])


TOL = {torch.float32: dict(y=1e-5, dy_dx=1e-5, dL_dx=1e-5, dL_dparam=2e-5, dL_ddLdy=1e-5, dL_dparam2=2e-5, dL_dx2=2e-5, y_maxlevel1=1e-5,
                           dL_dparam_maxlevel1=2e-5),
       torch.float16: dict(y=2e-3, dy_dx=2e-3, dL_dx=2e-3, dL_dparam=3e-2, dL_ddLdy=2e-3, dL_dparam2=3e-2, dL_dx2=2e-3, y_maxlevel1=2e-3,
                           dL_dparam_maxlevel1=3e-2)}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(FOREST_LOTD_CONFIGS))
@pytest.mark.parametrize("pdtype", [torch.float32, torch.float16])
def test_forest_lotd_vs_oracle(name, pdtype, dev):
    from nr3d_lib_b200.bindings import _lotd, _occ_grid
    from oracle import lotd_oracle as O
    cfg = FOREST_LOTD_CONFIGS[name]
    meta = _lotd.LoDMeta(*meta_args(cfg))
    om = O.OracleMeta(*meta_args(cfg))
    inp = forest_lotd_inputs(cfg, meta.n_params, N=2000, seed=7)
    got = _run_all(_lotd, (meta, _forest_meta(_occ_grid.ForestMeta, inp["forest"], dev)), inp, dev, pdtype)
    want = _oracle_all(om, inp, _oracle_forest(inp["forest"]), pdtype == torch.float16)
    bad = [(k, rel_err(got[k].float().cpu(), w)) for k, w in want.items() if not rel_err(got[k].float().cpu(), w) <= TOL[pdtype][k]]
    assert not bad, bad
    assert got["y"].dtype == pdtype and got["y"].is_contiguous() and got["dy_dx"].dtype == torch.float32
    if pdtype == torch.float32 and all(t in LINEAR_TYPES for t in cfg["types"]):      # adjoint identity through the CUDA path alone
        lhs = (got["y"].double() * inp["dL_dy"].to(dev).double()).sum()
        rhs = (inp["params"].to(dev).double() * got["dL_dparam"].double()).sum()
        assert abs(float(lhs - rhs)) < 1e-4 * max(1.0, abs(float(lhs)))


@pytest.mark.gpu
@pytest.mark.parametrize("name,tag", GOLDENS)
def test_forest_lotd_vs_golden_and_reference_build(name, tag, dev):
    from nr3d_lib_b200.bindings import _lotd, _occ_grid
    g = golden(f"forest_lotd_{name}_{tag}")
    if g is None:
        pytest.skip("golden fixture not generated yet")
    cfg = FOREST_LOTD_CONFIGS[name]
    pdtype = torch.float32 if tag == "f32" else torch.float16
    meta = _lotd.LoDMeta(*meta_args(cfg))
    inp = _golden_inputs(g)
    got = _run_all(_lotd, (meta, _forest_meta(_occ_grid.ForestMeta, inp["forest"], dev)), inp, dev, pdtype, fwd=False)
    ok2 = _param_masks(meta, cfg, g["block_ks"].shape[0])
    tol = TOL[pdtype]
    assert rel_err(got["dL_dparam"].float().cpu(), g["dL_dparam"].astype(np.float32)) < tol["dL_dparam"]
    assert rel_err(got["dL_dparam_maxlevel1"].float().cpu(), g["dL_dparam_maxlevel1"].astype(np.float32)) < tol["dL_dparam"]
    assert rel_err(got["dL_dx2"].cpu(), g["dL_dx2"]) < (1e-4 if tag == "f32" else 3e-2)
    if ok2.any():
        assert rel_err(got["dL_dparam2"].float().cpu()[ok2], torch.from_numpy(g["dL_dparam2"].astype(np.float32))[ok2]) < tol["dL_dparam2"]
    ref, fm = load_ref("_lotd"), load_ref("_forest")
    if ref is not None and fm is not None:      # live A/B on fresh, larger inputs
        inp2 = forest_lotd_inputs(cfg, meta.n_params, N=20000, seed=11)
        mine = _run_all(_lotd, (meta, _forest_meta(_occ_grid.ForestMeta, inp2["forest"], dev)), inp2, dev, pdtype, fwd=False)
        theirs = _run_all(ref, (ref.LoDMeta(*meta_args(cfg)), _forest_meta(fm.ForestMeta, inp2["forest"], dev)), inp2, dev, pdtype, fwd=False)
        assert rel_err(mine["dL_dparam"].float().cpu(), theirs["dL_dparam"].float().cpu()) < tol["dL_dparam"]
        assert rel_err(mine["dL_dx2"].cpu(), theirs["dL_dx2"].cpu()) < (1e-4 if tag == "f32" else 3e-2)


@pytest.mark.gpu
def test_forest_lotd_dydx_finite_differences_and_edges(dev):
    from nr3d_lib_b200.bindings import _lotd, _occ_grid
    cfg = FOREST_LOTD_CONFIGS["smooth"]
    meta = _lotd.LoDMeta(*meta_args(cfg))
    inp = forest_lotd_inputs(cfg, meta.n_params, N=1500, seed=9)
    fmeta = _forest_meta(_occ_grid.ForestMeta, inp["forest"], dev)
    x, params, bi = inp["x"].to(dev), inp["params"].to(dev), inp["batch_inds"].to(dev)
    y, dydx = _lotd.lod_fwd((meta, fmeta), x, params, bi, None, None, None, True)
    dydx = dydx.view(x.shape[0], -1, 3)
    eps = 1e-3
    safe = ((x > 2 * eps) & (x < 1 - 2 * eps)).all(dim=1) & (bi >= 0)
    for R in cfg["res"]:                                                               # x +- eps must stay inside one cell of every level
        u = x * R + 0.5
        safe &= (((u - u.floor()) > 0.1) & ((u - u.floor()) < 0.9)).all(dim=1)
    assert int(safe.sum()) > 100
    for d in range(3):
        e = torch.zeros(1, 3, device=dev); e[0, d] = eps
        yp, _ = _lotd.lod_fwd((meta, fmeta), (x + e).clamp(0, 1), params, bi, None, None, None, False)
        ym, _ = _lotd.lod_fwd((meta, fmeta), (x - e).clamp(0, 1), params, bi, None, None, None, False)
        fd = (yp - ym) / (2 * eps)
        # smoothstep is cubic: central differences carry an O(eps^2 * scale^3) term, so compare on a coarse tolerance
        assert rel_err(dydx[safe][..., d], fd[safe]) < 5e-2
    # continuity switch: interior points identical, face points differ
    y_nc, _ = _lotd.lod_fwd((meta, _forest_meta(_occ_grid.ForestMeta, inp["forest"], dev, continuity=False)), x, params, bi, None, None, None, False)
    interior = ((x > 0.2) & (x < 0.8)).all(dim=1)
    assert torch.equal(y[interior], y_nc[interior]) and not torch.equal(y, y_nc)
    # errors: unsupported level type, wrong block_ks dtype, CPU tensors
    with pytest.raises(RuntimeError):
        bad = _lotd.LoDMeta(3, [4, 8], [2, 2], ["Dense", "CPfast"], None, False)
        _lotd.lod_fwd((bad, fmeta), x, torch.zeros(bad.n_params * fmeta.n_trees, device=dev), bi, None, None, None, False)
    with pytest.raises(RuntimeError):
        f2 = _forest_meta(_occ_grid.ForestMeta, inp["forest"], dev); f2.block_ks = f2.block_ks.int()
        _lotd.lod_fwd((meta, f2), x, params, bi, None, None, None, False)
    with pytest.raises(RuntimeError):
        _lotd.lod_fwd((meta, fmeta), x.cpu(), params.cpu(), bi.cpu(), None, None, None, False)
    # empty batch
    y0, _ = _lotd.lod_fwd((meta, fmeta), x[:0], params, bi[:0], None, None, None, False)
    assert y0.shape == (0, meta.n_encoded_dims)


@pytest.mark.gpu
def test_forest_lotd_through_autograd_wrappers(dev):
    """metas=(LoDMeta, ForestMeta) through LoTDFunction (mirror of the reference wrapper, lotd.py:48-119), as the reference's
    forest unit test drives it (lotd/tests/math_test_forest.py:126-139)."""
    from nr3d_lib_b200.bindings import _lotd, _occ_grid
    from nr3d_lib_b200.lotd import LoTDFunction
    from oracle import lotd_oracle as O
    cfg = FOREST_LOTD_CONFIGS["mixed"]
    meta = _lotd.LoDMeta(*meta_args(cfg))
    inp = forest_lotd_inputs(cfg, meta.n_params, N=512, seed=13)
    metas = (meta, _forest_meta(_occ_grid.ForestMeta, inp["forest"], dev))
    x = inp["x"].to(dev).requires_grad_(True)
    params = inp["params"].to(dev).requires_grad_(True)
    y = LoTDFunction.apply(metas, x, params, inp["batch_inds"].to(dev), None, None, 1.0, None)
    (y * inp["dL_dy"].to(dev)).sum().backward()
    om = O.OracleMeta(*meta_args(cfg))
    gx, gp = O.bwd(om, inp["dL_dy"], inp["x"], inp["params"], batch_inds=inp["batch_inds"], forest=_oracle_forest(inp["forest"]))
    assert rel_err(x.grad.cpu(), gx) < 1e-5 and rel_err(params.grad.cpu(), gp) < 2e-5


@pytest.mark.gpu
def test_forest_lotd_batch_data_size_and_offsets(dev):
    """Block selection by `batch_data_size` (points grouped per block) and block tables addressed through `batch_offsets`
    (lotd_forest.h:214-224): a permuted parameter storage must give the same features / gradients at the permuted places."""
    from nr3d_lib_b200.bindings import _lotd, _occ_grid
    from oracle import lotd_oracle as O
    cfg = FOREST_LOTD_CONFIGS["hash_f4"]
    meta = _lotd.LoDMeta(*meta_args(cfg))
    om = O.OracleMeta(*meta_args(cfg))
    inp = forest_lotd_inputs(cfg, meta.n_params, N=5 * 300, seed=21)
    B = inp["forest"]["block_ks"].shape[0]
    assert B == 5
    fmeta = _forest_meta(_occ_grid.ForestMeta, inp["forest"], dev)
    F = _oracle_forest(inp["forest"])
    x, params, gy = inp["x"].to(dev), inp["params"].to(dev), inp["dL_dy"].to(dev)
    # (1) batch_data_size: point i lies in block i // 300
    y, dydx = _lotd.lod_fwd((meta, fmeta), x, params, None, None, 300, None, True)
    _, gp = _lotd.lod_bwd((meta, fmeta), gy, x, params, dydx, None, None, 300, None, False, True)
    y_o = O.encode(om, inp["x"], inp["params"], batch_data_size=300, forest=F)
    _, gp_o = O.bwd(om, inp["dL_dy"], inp["x"], inp["params"], batch_data_size=300, forest=F)
    assert rel_err(y.cpu(), y_o) < 1e-5 and rel_err(gp.cpu(), gp_o) < 2e-5
    # (2) batch_offsets: block b's table stored at slot perm[b]
    perm = torch.tensor([3, 0, 4, 1, 2])
    offs = (perm * meta.n_params).to(dev)
    p_perm = torch.empty_like(params)
    for b in range(B):
        p_perm[int(perm[b]) * meta.n_params:(int(perm[b]) + 1) * meta.n_params] = params[b * meta.n_params:(b + 1) * meta.n_params]
    bi = inp["batch_inds"].to(dev)
    y1, _ = _lotd.lod_fwd((meta, fmeta), x, params, bi, None, None, None, False)
    y2, _ = _lotd.lod_fwd((meta, fmeta), x, p_perm, bi, offs, None, None, False)
    assert rel_err(y2.cpu(), y1.cpu()) < 1e-6
    _, g1 = _lotd.lod_bwd((meta, fmeta), gy, x, params, None, bi, None, None, None, False, True)
    _, g2 = _lotd.lod_bwd((meta, fmeta), gy, x, p_perm, None, bi, offs, None, None, False, True)
    for b in range(B):
        a = g1[b * meta.n_params:(b + 1) * meta.n_params]
        c = g2[int(perm[b]) * meta.n_params:(int(perm[b]) + 1) * meta.n_params]
        assert rel_err(c.cpu(), a.cpu()) < 2e-5
    with pytest.raises(RuntimeError):      # batch_offsets must have one entry per block
        _lotd.lod_fwd((meta, fmeta), x, p_perm, bi, offs[:3], None, None, False)
