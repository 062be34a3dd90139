"""Full-size checks of BASELINE.json's configs[2] (march + composite, 1024^2 rays) and configs[3] (batched mixed-type LoTD with second
order, 2 Mi points) -- the sizes the CPU oracles cannot process in seconds.  Rays / packs / points are independent of each other, so a
random SUBSET of the full-size result must equal what the oracle computes for exactly those rays / points; global outputs are tied
down by size-independent identities (offsets = exclusive scan of counts, every sample in an occupied voxel, Euler's identity for the
multilinear level types)."""
import numpy as np
import pytest
import torch

from tests.util import C4_ARGS, c4_inputs, elementwise_excess, load_ref, rel_err, run_ref_worker

pytestmark = pytest.mark.gpu


def _rays(n, seed):
    rs = np.random.RandomState(seed)
    o = rs.randn(n, 3)
    o = (4.0 * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    tgt = (rs.rand(n, 3) - 0.5).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    with np.errstate(divide="ignore"):
        t1, t2 = (-1.0 - o) / d, (1.0 - o) / d
    near = np.maximum(np.minimum(t1, t2).max(1), 0.0).astype(np.float32)
    far = np.maximum(t1, t2).min(1).astype(np.float32)
    far[far <= near] = near[far <= near]
    return o, d, near, far


def test_config2_march_and_composite_at_1024x1024_rays(dev):
    from nr3d_lib_b200.bindings import _occ_grid, _pack_ops
    from oracle import march_oracle as MO, pack_oracle as PO
    R = 1024 * 1024
    o, d, near, far = _rays(R, 7)
    rs = np.random.RandomState(8)
    grid = rs.rand(128, 128, 128) > 0.5
    roi = np.array([-1, -1, -1, 1, 1, 1], dtype=np.float32)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi, t0, t1, ridx, gidx = _occ_grid.ray_marching(t(o), t(d), t(near), t(far), t(roi), t(grid), _occ_grid.ContractionType.AABB, 0.01, 1e10, 0.0, 512, True)
    S = t0.shape[0]
    assert S > 100_000_000 and int(pi[:, 1].max()) <= 512
    # global identities
    cnt = pi[:, 1].long()
    assert torch.equal(pi[:, 0].long(), torch.cumsum(cnt, 0) - cnt) and int(cnt.sum()) == S
    assert bool(t(grid).reshape(-1)[gidx.long()].all())                                   # every sample sits in an occupied voxel
    assert bool((t1 > t0).all()) and bool((ridx[1:] >= ridx[:-1]).all())
    assert torch.equal(torch.repeat_interleave(torch.arange(R, device=dev, dtype=torch.int32), cnt), ridx)
    # a random subset of rays against the C oracle, bit for bit
    sub = np.sort(rs.choice(R, 3000, replace=False))
    m = MO.ray_marching(o[sub], d[sub], near[sub], far[sub], roi, grid, 0, 0.01, 1e10, 0.0, 512)
    pi_c = pi.cpu().numpy()
    assert np.array_equal(pi_c[sub, 1], m["packed_info"][:, 1])
    sel = np.concatenate([np.arange(b, b + n) for b, n in pi_c[sub]])
    assert np.array_equal(t0.cpu().numpy()[sel, 0], m["t_starts"]) and np.array_equal(t1.cpu().numpy()[sel, 0], m["t_ends"])
    assert np.array_equal(gidx.cpu().numpy()[sel], m["gidx"])
    # composite on all 1.2e8 samples; the subset's packs against the sequential numpy oracle, bit for bit
    g = torch.Generator(device=dev).manual_seed(3)
    alphas = (torch.rand(S, device=dev, generator=g) ** 2 * 0.2).contiguous()
    hit = cnt > 0
    pack_infos = torch.stack([pi[:, 0].long(), cnt], 1)[hit].contiguous()
    w = _pack_ops.packed_alpha_to_vw_forward(alphas, pack_infos, 1e-4, 0.0, False)[0]
    acc = _pack_ops.packed_sum(w, pack_infos)
    assert float(acc.max()) <= 1.0 + 1e-5 and float(w.min()) >= 0.0                       # opacities are probabilities
    sub_packs = pi_c[sub][pi_c[sub, 1] > 0].astype(np.int64)
    a_sub = alphas.cpu().numpy()
    w_sub = w.cpu().numpy()
    local = np.stack([np.cumsum(sub_packs[:, 1]) - sub_packs[:, 1], sub_packs[:, 1]], 1)
    a_cat = np.concatenate([a_sub[b:b + n] for b, n in sub_packs])
    w_ref, _, _ = PO.alpha_to_vw_forward(a_cat, local, 1e-4, 0.0)
    assert np.array_equal(np.concatenate([w_sub[b:b + n] for b, n in sub_packs]), w_ref)


def test_config3_batched_mixed_second_order_at_2Mi_points(dev):
    from nr3d_lib_b200.bindings import _lotd
    from oracle import lotd_oracle as O
    args = C4_ARGS
    meta, om = _lotd.LoDMeta(*args), O.OracleMeta(*args)
    N, B = 2 * 1024 * 1024, 8
    inp = c4_inputs(N, meta.n_params, seed=11, B=B)
    x, bi, params, gy, ddx = inp["x"], inp["batch_inds"], inp["params"], inp["dL_dy"], inp["dL_ddLdx"]
    g = torch.Generator().manual_seed(12)
    xd, bd, pd, gd, dd = x.to(dev), bi.to(dev), params.to(dev), gy.to(dev), ddx.to(dev)
    kw = dict(batch_inds=bd, batch_offsets=None, batch_data_size=None, max_level=None)
    y, dydx = _lotd.lod_fwd(meta, xd, pd, need_input_grad=True, **kw)
    dL_dx, dL_dp = _lotd.lod_bwd(meta, gd, xd, pd, dydx, need_input_grad=True, need_param_grad=True, **kw)
    a, b, c = _lotd.lod_bwd_bwd_input(meta, dd, gd, xd, pd, dydx, need_dLdinput_ddLdoutput=True, need_dLdinput_dparams=True,
                                      need_dLdinput_dinput=True, **kw)
    # ---- table against table: the reference's own CUDA kernels on the same 2 Mi points (first order: stock build, live) ----
    ref = load_ref("_lotd")
    if ref is not None:
        m_r = ref.LoDMeta(*args)
        y_r, _ = ref.lod_fwd(m_r, xd, pd, need_input_grad=False, **kw)
        assert rel_err(y, y_r) < 1e-5
        _, gp_r = ref.lod_bwd(m_r, gd, xd, pd, None, need_input_grad=False, need_param_grad=True, **kw)
        # sum of |terms| per entry (all weights and factor products are taken positive): what fp32 summation order may change
        _, mag = _lotd.lod_bwd(meta, gd.abs(), xd, pd.abs(), None, need_input_grad=False, need_param_grad=True, **kw)
        bad, worst = elementwise_excess(dL_dp.cpu().numpy(), gp_r.cpu().numpy(), mag.cpu().numpy(), rel=1e-5, c_eps=16.0)
        assert bad == 0, f"first-order dL/dparam vs the reference build, element-wise: {bad} entries out of tolerance (worst x{worst:.2f})"
        del y_r, gp_r, mag
    # ---- second order against the -G build of the reference's generic kernels (subprocess; per-level max-norm 5e-5) ----
    chk = run_ref_worker([dict(name="c4", dtype="f32", N=N, seed=11, keys=["dL_dparam2", "dL_ddLdy", "dy_dx", "dL_dx"])], variant="G", timeout=3000)
    if chk is not None:
        chk = chk[0]
        off = list(meta.level_offsets)
        g2 = b.view(B, -1).cpu()
        r2 = torch.from_numpy(chk["dL_dparam2"]).view(B, -1)
        for lvl in range(meta.n_levels):
            e = rel_err(g2[:, off[lvl]:off[lvl + 1]], r2[:, off[lvl]:off[lvl + 1]])
            # line / plane tables of 96 - 1536 entries per scene receive 2 Mi x 24 fp32 terms in arbitrary order in BOTH builds
            assert e < 5e-5, ("dL_dparam2 level", lvl, e)
        assert rel_err(a.cpu(), chk["dL_ddLdy"]) < 1e-5 and rel_err(dL_dx.cpu(), chk["dL_dx"]) < 1e-5
        assert rel_err(dydx.view(N, -1, 3).cpu(), chk["dy_dx"]) < 1e-5
        del chk, g2, r2
    # per-point outputs of a random subset against the float64 oracle
    sub = torch.sort(torch.randperm(N, generator=g)[:4000]).values
    okw = dict(batch_inds=bi[sub])
    y_o, dydx_o = O.fwd_dydx(om, x[sub], params, **okw)
    gx_o, _ = O.bwd(om, gy[sub], x[sub], params, **okw)
    a_o, _, c_o = O.bwd_bwd_input(om, ddx[sub], gy[sub], x[sub], params, **okw)
    sd = sub.to(dev)
    assert rel_err(y[sd].cpu(), y_o) < 1e-5 and rel_err(dydx.view(N, -1, 3)[sd].cpu(), dydx_o) < 1e-5
    assert rel_err(dL_dx[sd].cpu(), gx_o) < 1e-5 and rel_err(a[sd].cpu(), a_o) < 1e-5 and rel_err(c[sd].cpu(), c_o) < 5e-5
    assert float(y[bd < 0].abs().max()) == 0.0 and float(dL_dx[bd < 0].abs().max()) == 0.0
    # global gradients through Euler's identity: every level type here is homogeneous in its own table -- degree 1 (Dense), 2 (VM:
    # plane x line), 3 (CP: three lines) -- so <params_l, dL/dparams_l> = degree_l * <y_l, dL_dy_l> per level, summed over scenes
    off, deg = list(meta.level_offsets), {"Dense": 1.0, "VM": 2.0, "CP": 3.0}
    pb, gb, g2b = pd.view(B, -1).double(), dL_dp.view(B, -1).double(), b.view(B, -1).double()
    f0 = 0
    for lvl, (tp, nf) in enumerate(zip(args[3], args[2])):
        lhs = float((pb[:, off[lvl]:off[lvl + 1]] * gb[:, off[lvl]:off[lvl + 1]]).sum())
        rhs = deg[tp] * float((y[:, f0:f0 + nf].double() * gd[:, f0:f0 + nf].double()).sum())
        assert abs(lhs - rhs) < 2e-4 * max(1.0, abs(rhs)), (lvl, tp, lhs, rhs)
        # second order: <params_l, d(dL/dx . v)/dparams_l> = degree_l * <dL/dx_l, v> with dL/dx_l = sum_j dL_dy_j dy_j/dx restricted to the level
        dldx_l = (dydx.view(N, -1, 3)[:, f0:f0 + nf].double() * gd[:, f0:f0 + nf].double().unsqueeze(-1)).sum(1)
        lhs2 = float((pb[:, off[lvl]:off[lvl + 1]] * g2b[:, off[lvl]:off[lvl + 1]]).sum())
        rhs2 = deg[tp] * float((dldx_l * dd.double()).sum())
        assert abs(lhs2 - rhs2) < 5e-4 * max(1.0, abs(rhs2)), (lvl, tp, lhs2, rhs2)
        f0 += nf
