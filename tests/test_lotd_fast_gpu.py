"""GPU parity tests of the cell-sorted fast path (lotd_fast.cu, LoDMeta.c_sort_points) against the generic kernels,
the reference CUDA build and the float64 oracle -- at small sizes and at BASELINE.json's full size (4 Mi points, 16-level NGP)."""
import numpy as np
import pytest
import torch

from tests.util import LOTD_CONFIGS, elementwise_excess, load_ref, lotd_inputs, meta_args, rel_err

pytestmark = pytest.mark.gpu


def _ngp16():
    res = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
    return (3, res, [2] * 16, ["Dense" if r ** 3 <= 2 ** 19 else "Hash" for r in res], 2 ** 19, False)


@pytest.mark.parametrize("name", ["ngp8", "ngp_smooth", "batched_hash"])
@pytest.mark.parametrize("N", [1, 33, 5000])
def test_sorted_path_matches_generic_and_oracle(name, N, dev):
    from nr3d_lib_b200.bindings import _lotd
    from oracle import lotd_oracle as O
    cfg = dict(LOTD_CONFIGS[name], B=1)
    meta_s, meta_g = _lotd.LoDMeta(*meta_args(cfg)), _lotd.LoDMeta(*meta_args(cfg))
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    inp = lotd_inputs(cfg, meta_g.n_params, N=N, seed=N)
    x, p, gy = inp["x"].to(dev), inp["params"].to(dev), inp["dL_dy"].to(dev)
    _lotd.clear_sort_cache()
    for max_level in (None, 1):
        y_s, none = _lotd.lod_fwd(meta_s, x, p, max_level=max_level, need_input_grad=False)
        y_g, _ = _lotd.lod_fwd(meta_g, x, p, max_level=max_level, need_input_grad=False)
        assert none is None and y_s.is_contiguous() and y_s.shape == y_g.shape
        assert rel_err(y_s.cpu(), y_g.cpu()) < 1e-6
        _, g_s = _lotd.lod_bwd(meta_s, gy, x, p, None, max_level=max_level, need_input_grad=False, need_param_grad=True)
        _, g_g = _lotd.lod_bwd(meta_g, gy, x, p, None, max_level=max_level, need_input_grad=False, need_param_grad=True)
        assert rel_err(g_s.cpu(), g_g.cpu()) < 2e-5
        # strided dL_dy (feature-major storage) goes through the same kernel
        gy_t = gy.t().contiguous().t()
        _, g_t = _lotd.lod_bwd(meta_s, gy_t, x, p, None, max_level=max_level, need_input_grad=False, need_param_grad=True)
        assert rel_err(g_t.cpu(), g_g.cpu()) < 2e-5
    om = O.OracleMeta(*meta_args(cfg))
    assert rel_err(y_s.cpu(), O.encode(om, inp["x"], inp["params"], max_level=1)) < 1e-5
    # the sorted records must follow in-place edits of x -- with and without a version-counter bump
    x.mul_(0.5)
    y2, _ = _lotd.lod_fwd(meta_s, x, p, need_input_grad=False)
    assert rel_err(y2.cpu(), O.encode(om, x.cpu(), inp["params"])) < 1e-5
    x.data.mul_(0.9)                      # no _version bump (stale-sort hazard of a host-side cache)
    x.data.add_(0.01)
    y2b, _ = _lotd.lod_fwd(meta_s, x, p, need_input_grad=False)
    assert rel_err(y2b.cpu(), O.encode(om, x.cpu(), inp["params"])) < 1e-5
    _, g2b = _lotd.lod_bwd(meta_s, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    _, g2g = _lotd.lod_bwd(meta_g, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    assert rel_err(g2b.cpu(), g2g.cpu()) < 2e-5
    x.data.mul_(1.0 / 0.9)                # edit between forward and backward of the "same" tensor: the backward must see the new points
    _, g2c = _lotd.lod_bwd(meta_s, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    _, g2d = _lotd.lod_bwd(meta_g, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    assert rel_err(g2c.cpu(), g2d.cpu()) < 2e-5
    y2, _ = _lotd.lod_fwd(meta_s, x, p, need_input_grad=False)
    # dy_dx requests stay on the fast path: same shapes as the reference returns, contiguous, values equal to the generic kernels
    y3, dy = _lotd.lod_fwd(meta_s, x, p, need_input_grad=True)
    y3g, dyg = _lotd.lod_fwd(meta_g, x, p, need_input_grad=True)
    assert dy.shape == dyg.shape and dy.is_contiguous() and y3.is_contiguous() and torch.equal(y3, y2)
    assert rel_err(dy.cpu(), dyg.cpu()) < 1e-5 and rel_err(y3.cpu(), y3g.cpu()) < 1e-6


@pytest.mark.parametrize("name", ["ngp8", "ngp_smooth"])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.float16])
def test_sorted_path_nablas_and_second_order(name, pdtype, dev):
    """NeuS-style call pattern on the fast path: forward with dy/dx, dL/dx, and the second-order backward (dL_ddLdy, dparam, dx)
    against the generic kernels and the float64 oracle."""
    from nr3d_lib_b200.bindings import _lotd
    from oracle import lotd_oracle as O
    cfg = LOTD_CONFIGS[name]
    meta_s, meta_g = _lotd.LoDMeta(*meta_args(cfg)), _lotd.LoDMeta(*meta_args(cfg))
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    N = 6000
    inp = lotd_inputs(cfg, meta_g.n_params, N=N, seed=5)
    x, p, gy, ddx = inp["x"].to(dev), inp["params"].to(dev).to(pdtype), inp["dL_dy"].to(dev).to(pdtype), inp["dL_ddLdx"].to(dev)
    half = pdtype == torch.float16
    tol, tol_at = (2e-3, 3e-2) if half else (1e-5, 2e-5)
    out = {}
    for tag, meta in (("s", meta_s), ("g", meta_g)):
        _lotd.clear_sort_cache()
        y, dydx = _lotd.lod_fwd(meta, x, p, need_input_grad=True)
        dL_dx, dL_dp = _lotd.lod_bwd(meta, gy, x, p, dydx, need_input_grad=True, need_param_grad=True)
        a, b, c = _lotd.lod_bwd_bwd_input(meta, ddx, gy, x, p, dydx, need_dLdinput_ddLdoutput=True, need_dLdinput_dparams=True,
                                          need_dLdinput_dinput=True)
        out[tag] = dict(y=y, dy_dx=dydx.reshape(N, -1, 3), dL_dx=dL_dx, dL_dparam=dL_dp, dL_ddLdy=a, dL_dparam2=b, dL_dx2=c)
    for k in out["s"]:
        t = tol_at if k in ("dL_dparam", "dL_dparam2", "dL_dx2") else tol
        assert rel_err(out["s"][k].float().cpu(), out["g"][k].float().cpu()) < t, k
    om = O.OracleMeta(*meta_args(cfg))
    pp, gg = (inp["params"].half().float(), inp["dL_dy"].half().float()) if half else (inp["params"], inp["dL_dy"])
    y_o, dydx_o = O.fwd_dydx(om, inp["x"], pp)
    a_o, b_o, c_o = O.bwd_bwd_input(om, inp["dL_ddLdx"], gg, inp["x"], pp)
    assert rel_err(out["s"]["y"].float().cpu(), y_o) < tol and rel_err(out["s"]["dy_dx"].cpu(), dydx_o) < tol
    assert rel_err(out["s"]["dL_dparam2"].float().cpu(), b_o) < tol_at and rel_err(out["s"]["dL_ddLdy"].float().cpu(), a_o) < tol
    # max_level and an empty batch
    y1, d1 = _lotd.lod_fwd(meta_s, x, p, max_level=1, need_input_grad=True)
    y1g, d1g = _lotd.lod_fwd(meta_g, x, p, max_level=1, need_input_grad=True)
    assert rel_err(d1.reshape(N, -1, 3).cpu(), d1g.reshape(N, -1, 3).cpu()) < tol and rel_err(y1.float().cpu(), y1g.float().cpu()) < tol


def test_sorted_path_clustered_points(dev):
    """All points inside one coarse cell / on one line: maximal run merging and atomic contention."""
    from nr3d_lib_b200.bindings import _lotd
    cfg = LOTD_CONFIGS["ngp8"]
    meta_s, meta_g = _lotd.LoDMeta(*meta_args(cfg)), _lotd.LoDMeta(*meta_args(cfg))
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    rs = np.random.RandomState(0)
    N = 20000
    pts = np.concatenate([0.3 + 0.001 * rs.rand(N // 2, 3), np.stack([rs.rand(N // 2), np.full(N // 2, 0.7), np.full(N // 2, 0.2)], 1)]).astype(np.float32)
    x = torch.from_numpy(pts).to(dev)
    p = torch.randn(meta_g.n_params, device=dev) * 0.1
    gy = torch.randn(N, meta_g.n_encoded_dims, device=dev)
    y_s, _ = _lotd.lod_fwd(meta_s, x, p, need_input_grad=False)
    y_g, _ = _lotd.lod_fwd(meta_g, x, p, need_input_grad=False)
    assert rel_err(y_s.cpu(), y_g.cpu()) < 1e-6
    _, g_s = _lotd.lod_bwd(meta_s, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    _, g_g = _lotd.lod_bwd(meta_g, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    assert rel_err(g_s.cpu().double(), g_g.cpu().double()) < 1e-4   # thousands of fp32 terms per slot, different orders


def test_full_size_headline_config(dev):
    """BASELINE.json configs[1] at full size: 16-level NGP LoTD, 4 Mi uniform points (size-independent properties +
    sampled comparison with the oracle and, when built, the reference CUDA kernels)."""
    from nr3d_lib_b200.bindings import _lotd
    from oracle import lotd_oracle as O
    args = _ngp16()
    meta_s, meta_g = _lotd.LoDMeta(*args), _lotd.LoDMeta(*args)
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    N = 4 * 1024 * 1024
    g = torch.Generator(device=dev).manual_seed(42)
    x = torch.rand(N, 3, device=dev, generator=g).clamp(1e-6, 1 - 1e-6)
    p = (torch.rand(meta_g.n_params, device=dev, generator=g) * 2 - 1) * 1e-4
    gy = torch.randn(N, 32, device=dev, generator=g) * 1e-4
    y_s, _ = _lotd.lod_fwd(meta_s, x, p, need_input_grad=False)
    y_g, _ = _lotd.lod_fwd(meta_g, x, p, need_input_grad=False)
    assert rel_err(y_s, y_g) < 1e-6
    _, g_s = _lotd.lod_bwd(meta_s, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    _, g_g = _lotd.lod_bwd(meta_g, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    assert rel_err(g_s.double(), g_g.double()) < 1e-4
    # linearity of the backward in dL_dy and adjointness <y, gy> == <params, dL/dparams> (y is linear in the parameters)
    lhs = (y_g.double() * gy.double()).sum().item()
    rhs = (p.double() * g_s.double()).sum().item()
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), 1e-30), (lhs, rhs)
    # sampled oracle check (float64) of forward and of the gradient restricted to the sampled points' contribution
    idx = torch.randperm(N, device=dev)[:4096]
    om = O.OracleMeta(*args)
    y_o = O.encode(om, x[idx].cpu(), p.cpu())
    assert rel_err(y_s[idx].cpu(), y_o) < 1e-5
    # ---- the whole gradient table, element-wise, with the float64 oracle as arbiter (north_star: 1e-5) ----
    #   |ours - f64| <= 2 |ref - f64| + 1e-5 |f64| + 8 eps32 sum|terms|
    g64, mag = O.bwd_param_hash_f64(om, gy.cpu(), x.cpu())
    ref = load_ref("_lotd")
    g_r = None
    if ref is not None:
        m_r = ref.LoDMeta(*args)
        y_r, _ = ref.lod_fwd(m_r, x, p, need_input_grad=False)
        assert rel_err(y_s, y_r) < 1e-5
        _, g_r = ref.lod_bwd(m_r, gy, x, p, None, need_input_grad=False, need_param_grad=True)
        g_r = g_r.cpu().numpy()
    for tag, tab in (("fast path", g_s), ("generic kernels", g_g)):
        bad, worst = elementwise_excess(tab.cpu().numpy(), g64, mag, ref=g_r)
        assert bad == 0, f"{tag}: {bad} of {g64.size} gradient entries out of the element-wise tolerance (worst x{worst:.2f})"
    assert rel_err(g_s.cpu().double(), torch.from_numpy(g64)) < 1e-5          # and the max-norm figure of north_star


@pytest.mark.parametrize("name,N", [("ngp8", 37), ("ngp8", 6000), ("ngp_smooth", 3000)])
def test_sorted_path_fp16_params(name, N, dev):
    """fp16 tables on the fast path: y in half (every term rounded to half like the reference, linear_interpolate.cuh:118),
    dL/dparam scattered with packed-half reductions.  Checked against the generic kernels, the float64 oracle and, when
    built, the reference CUDA extension -- at fp16 tolerances (the reference's own fp16 atomics are order dependent)."""
    from nr3d_lib_b200.bindings import _lotd
    from oracle import lotd_oracle as O
    cfg = dict(LOTD_CONFIGS[name], B=1)
    meta_s, meta_g = _lotd.LoDMeta(*meta_args(cfg)), _lotd.LoDMeta(*meta_args(cfg))
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    inp = lotd_inputs(cfg, meta_g.n_params, N=N, seed=11)
    x, p, gy = inp["x"].to(dev), inp["params"].to(dev).half(), inp["dL_dy"].to(dev).half()
    _lotd.clear_sort_cache()
    y_s, _ = _lotd.lod_fwd(meta_s, x, p, need_input_grad=False)
    y_g, _ = _lotd.lod_fwd(meta_g, x, p, need_input_grad=False)
    assert y_s.dtype == torch.float16 and y_s.is_contiguous()
    om = O.OracleMeta(*meta_args(cfg))
    y_o = O.encode(om, inp["x"], p.float().cpu())
    assert rel_err(y_s.float().cpu(), y_o) < 4e-3 and rel_err(y_g.float().cpu(), y_o) < 4e-3     # 8 half roundings per feature
    _, g_s = _lotd.lod_bwd(meta_s, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    _, g_g = _lotd.lod_bwd(meta_g, gy, x, p, None, need_input_grad=False, need_param_grad=True)
    assert g_s.dtype == torch.float16
    g_o = O.bwd(om, gy.float().cpu(), inp["x"], p.float().cpu())[1]
    assert rel_err(g_s.float().cpu(), g_o) < 2e-2 and rel_err(g_g.float().cpu(), g_o) < 3e-2
    ref = load_ref("_lotd")
    if ref is not None:
        rmeta = ref.LoDMeta(*meta_args(cfg))
        y_r, _ = ref.lod_fwd(rmeta, x, p, None, None, None, None, False)
        assert rel_err(y_s.float().cpu(), y_r.float().cpu()) < 4e-3


WIDE_CONFIGS = {
    # F_pl = 4 (one level is 8 wide: two pseudo levels on one table) and F_pl = 8, Dense + Hash, non power-of-two hash size included
    "hash_f4": LOTD_CONFIGS["hash_f4"],
    "hash_f8": dict(D=3, res=[9, 24, 50], feats=[8, 8, 8], types=["Dense", "Hash", "Hash"], T=1000, smooth=False, B=1),
    "hash_f4_smooth": dict(D=3, res=[[8, 10, 12], 31], feats=[4, 4], types=["Dense", "Hash"], T=2 ** 10, smooth=True, B=1),
}


@pytest.mark.parametrize("name", list(WIDE_CONFIGS))
@pytest.mark.parametrize("pdtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("N", [7, 4000])
def test_sorted_path_wide_pseudo_levels(name, pdtype, N, dev):
    """F_pl = 4 and 8 on the fast path (the reference's hash-only kernels serve F in {2, 4, 8}, lotd_hash_only.h:35-55): forward, dy/dx,
    first- and second-order scatter against the generic kernels and the float64 oracle."""
    from nr3d_lib_b200.bindings import _lotd
    from oracle import lotd_oracle as O
    cfg = WIDE_CONFIGS[name]
    meta_s, meta_g = _lotd.LoDMeta(*meta_args(cfg)), _lotd.LoDMeta(*meta_args(cfg))
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    assert meta_s.c_hash_only and meta_s.n_feat_per_pseudo_lvl in (4, 8)
    inp = lotd_inputs(cfg, meta_g.n_params, N=N, seed=N + 1)
    half = pdtype == torch.float16
    tol, tol_at = (4e-3, 3e-2) if half else (1e-5, 2e-5)
    x, p, gy, ddx = inp["x"].to(dev), inp["params"].to(dev).to(pdtype), inp["dL_dy"].to(dev).to(pdtype), inp["dL_ddLdx"].to(dev)
    out = {}
    for tag, meta in (("s", meta_s), ("g", meta_g)):
        y0, _ = _lotd.lod_fwd(meta, x, p, need_input_grad=False)
        y, dydx = _lotd.lod_fwd(meta, x, p, need_input_grad=True)
        assert torch.equal(y, y0)
        dL_dx, dL_dp = _lotd.lod_bwd(meta, gy, x, p, dydx, need_input_grad=True, need_param_grad=True)
        a, b, c = _lotd.lod_bwd_bwd_input(meta, ddx, gy, x, p, dydx, need_dLdinput_ddLdoutput=True, need_dLdinput_dparams=True, need_dLdinput_dinput=True)
        y1, _ = _lotd.lod_fwd(meta, x, p, max_level=0, need_input_grad=False)
        out[tag] = dict(y=y, dy_dx=dydx.reshape(N, -1, 3), dL_dx=dL_dx, dL_dparam=dL_dp, dL_ddLdy=a, dL_dparam2=b, dL_dx2=c, y_max0=y1)
    assert out["s"]["y"].is_contiguous()
    for k in out["s"]:
        t = tol_at if k in ("dL_dparam", "dL_dparam2", "dL_dx2") else tol
        assert rel_err(out["s"][k].float().cpu(), out["g"][k].float().cpu()) < t, k
    om = O.OracleMeta(*meta_args(cfg))
    pp, gg = (inp["params"].half().float(), inp["dL_dy"].half().float()) if half else (inp["params"], inp["dL_dy"])
    y_o, dydx_o = O.fwd_dydx(om, inp["x"], pp)
    _, gp_o = O.bwd(om, gg, inp["x"], pp)
    assert rel_err(out["s"]["y"].float().cpu(), y_o) < tol and rel_err(out["s"]["dy_dx"].cpu(), dydx_o) < tol
    assert rel_err(out["s"]["dL_dparam"].float().cpu(), gp_o) < tol_at


@pytest.mark.parametrize("mode", ["inds", "size"])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.float16])
def test_sorted_path_batched_scenes(mode, pdtype, dev):
    """Several scenes on the fast path (LoTDBatched, lotd_batched.py:118-156): scene index in the sort key, per-scene tables, points with
    batch_inds < 0 skipped (zero features, no gradient) -- against the generic kernels, the oracle and, when built, the reference build."""
    from nr3d_lib_b200.bindings import _lotd
    from oracle import lotd_oracle as O
    cfg = LOTD_CONFIGS["batched_hash"]
    B = cfg["B"]
    meta_s, meta_g = _lotd.LoDMeta(*meta_args(cfg)), _lotd.LoDMeta(*meta_args(cfg))
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    N = 6000
    inp = lotd_inputs(cfg, meta_g.n_params, N=N, seed=23, batch_mode=mode)
    half = pdtype == torch.float16
    tol, tol_at = (4e-3, 3e-2) if half else (1e-5, 2e-5)
    x, p, gy, ddx = inp["x"].to(dev), inp["params"].to(dev).to(pdtype), inp["dL_dy"].to(dev).to(pdtype), inp["dL_ddLdx"].to(dev)
    bi = None if inp["batch_inds"] is None else inp["batch_inds"].to(dev)
    kw = dict(batch_inds=bi, batch_data_size=inp["batch_data_size"] or None)
    out = {}
    for tag, meta in (("s", meta_s), ("g", meta_g)):
        y, dydx = _lotd.lod_fwd(meta, x, p, need_input_grad=True, **kw)
        dL_dx, dL_dp = _lotd.lod_bwd(meta, gy, x, p, dydx, need_input_grad=True, need_param_grad=True, **kw)
        a, b, c = _lotd.lod_bwd_bwd_input(meta, ddx, gy, x, p, dydx, need_dLdinput_ddLdoutput=True, need_dLdinput_dparams=True, need_dLdinput_dinput=True, **kw)
        out[tag] = dict(y=y, dy_dx=dydx.reshape(N, -1, 3), dL_dx=dL_dx, dL_dparam=dL_dp, dL_ddLdy=a, dL_dparam2=b, dL_dx2=c)
    for k in out["s"]:
        t = tol_at if k in ("dL_dparam", "dL_dparam2", "dL_dx2") else tol
        assert rel_err(out["s"][k].float().cpu(), out["g"][k].float().cpu()) < t, k
    if bi is not None:
        skipped = (bi < 0)
        assert skipped.any() and out["s"]["y"][skipped].abs().max() == 0 and out["s"]["dy_dx"][skipped].abs().max() == 0
    om = O.OracleMeta(*meta_args(cfg))
    pp, gg = (inp["params"].half().float(), inp["dL_dy"].half().float()) if half else (inp["params"], inp["dL_dy"])
    okw = dict(batch_inds=inp["batch_inds"], batch_data_size=inp["batch_data_size"])
    y_o = O.encode(om, inp["x"], pp, **okw)
    _, gp_o = O.bwd(om, gg, inp["x"], pp, **okw)
    assert rel_err(out["s"]["y"].float().cpu(), y_o) < tol and rel_err(out["s"]["dL_dparam"].float().cpu(), gp_o) < tol_at
    ref = load_ref("_lotd")
    if ref is not None and not half:
        m_r = ref.LoDMeta(*meta_args(cfg))
        y_r, _ = ref.lod_fwd(m_r, x, p, need_input_grad=False, **kw)
        _, g_r = ref.lod_bwd(m_r, gy, x, p, None, need_input_grad=False, need_param_grad=True, **kw)
        assert rel_err(out["s"]["y"].cpu(), y_r.cpu()) < tol and rel_err(out["s"]["dL_dparam"].cpu(), g_r.cpu()) < tol_at


def test_sorted_path_streams_and_sizes(dev):
    """The sorted-record buffers are per (device, stream): interleaved calls with different point sets, sizes and streams never see each
    other's records."""
    from nr3d_lib_b200.bindings import _lotd
    cfg = LOTD_CONFIGS["ngp8"]
    meta_s, meta_g = _lotd.LoDMeta(*meta_args(cfg)), _lotd.LoDMeta(*meta_args(cfg))
    meta_s.c_sort_points, meta_g.c_sort_points = True, False
    p = torch.randn(meta_g.n_params, device=dev) * 0.1
    xa, xb, xc = torch.rand(3000, 3, device=dev), torch.rand(3000, 3, device=dev), torch.rand(777, 3, device=dev)
    torch.cuda.synchronize()
    side = torch.cuda.Stream(dev)
    want = {k: _lotd.lod_fwd(meta_g, v, p, need_input_grad=False)[0] for k, v in (("a", xa), ("b", xb), ("c", xc))}
    torch.cuda.synchronize()
    got = {}
    got["a"] = _lotd.lod_fwd(meta_s, xa, p, need_input_grad=False)[0]
    with torch.cuda.stream(side):
        got["b"] = _lotd.lod_fwd(meta_s, xb, p, need_input_grad=False)[0]
    got["c"] = _lotd.lod_fwd(meta_s, xc, p, need_input_grad=False)[0]
    got["a2"] = _lotd.lod_fwd(meta_s, xa, p, need_input_grad=False)[0]
    with torch.cuda.stream(side):
        got["b2"] = _lotd.lod_fwd(meta_s, xb, p, need_input_grad=False)[0]     # same points, same stream: verified hit
    torch.cuda.synchronize()
    for k, ref_k in (("a", "a"), ("b", "b"), ("c", "c"), ("a2", "a"), ("b2", "b")):
        assert rel_err(got[k].cpu(), want[ref_k].cpu()) < 1e-6, k


def test_two_level_sort_matches_one_level(dev):
    """The two-level sort (coarse partition with CTA-local staging + fine sort, csrc/lotd_sort.cu; default from 12 Mi points) forced at small
    sizes: the records are a permutation of the points, non-decreasing in the x-fastest bin key, the fingerprint path reuses them, and the
    encoder gives the same values as on the one-level records."""
    from nr3d_lib_b200 import _lib
    from nr3d_lib_b200.bindings import _lotd
    lib = _lib.get_lib()
    cfg = LOTD_CONFIGS["ngp8"]
    meta = _lotd.LoDMeta(*meta_args(cfg))
    g = torch.Generator().manual_seed(21)
    by_index = lambda r: r[r[:, 3].contiguous().view(torch.int32).argsort()]

    def bin_keys(r, res):
        b = (r[:, :3] * float(res)).clamp_min(0).floor().clamp_max(res - 1).long()
        return (b[:, 2] * res + b[:, 1]) * res + b[:, 0]

    try:
        for N, res in ((5000, 32), (300000, 64), (2500000, 128)):
            # ray-like input: short runs of neighbouring points, like the samples of the marcher
            base = torch.rand((N + 15) // 16, 1, 3, generator=g)
            x = (base + torch.arange(16).view(1, 16, 1) * 0.004 * torch.randn((N + 15) // 16, 1, 3, generator=g)).reshape(-1, 3)[:N]
            x = x.clamp(1e-6, 1 - 1e-6).contiguous().to(dev)
            x0 = x.clone()
            _lib.check(lib.nr3d_lotd_sort_set_two_level_min(1000))
            _lotd.clear_sort_cache()
            xs2, _ = _lotd._sorted_points(x, expect_new=True)
            k2 = bin_keys(xs2, res)
            assert torch.equal(by_index(xs2)[:, :3], x), N
            assert bool((k2[1:] >= k2[:-1]).all()), N
            ptr = xs2.data_ptr()
            xs2b, _ = _lotd._sorted_points(x)                        # fingerprint check: same records, nothing re-sorted
            assert xs2b.data_ptr() == ptr and torch.equal(bin_keys(xs2b, res), k2)
            p = (torch.rand(meta.n_params, generator=g) - 0.5).to(dev)
            y2, _ = _lotd.lod_fwd(meta, x, p, need_input_grad=False)
            _, g2 = _lotd.lod_bwd(meta, y2, x, p, None, need_input_grad=False, need_param_grad=True)
            x.mul_(0.999)                                            # new points through the verify path of the two-level sort
            _, g2n = _lotd.lod_bwd(meta, y2, x, p, None, need_input_grad=False, need_param_grad=True)
            _lib.check(lib.nr3d_lotd_sort_set_two_level_min(0))
            _lotd.clear_sort_cache()
            xs1, _ = _lotd._sorted_points(x, expect_new=True)
            k1 = bin_keys(xs1, res)
            assert bool((k1[1:] >= k1[:-1]).all())
            _, g1n = _lotd.lod_bwd(meta, y2, x, p, None, need_input_grad=False, need_param_grad=True)
            assert rel_err(g2n.cpu(), g1n.cpu()) < 2e-5
            y1, _ = _lotd.lod_fwd(meta, x0, p, need_input_grad=False)
            assert rel_err(y2.cpu(), y1.cpu()) < 1e-5
            _, g1 = _lotd.lod_bwd(meta, y2, x0, p, None, need_input_grad=False, need_param_grad=True)
            assert rel_err(g2.cpu(), g1.cpu()) < 2e-5
    finally:
        _lib.check(lib.nr3d_lotd_sort_set_two_level_min(0))
        _lotd.clear_sort_cache()
