"""GPU parity tests for pack_ops and the occupancy-grid marcher (through the C-ABI bindings).

Checkers: golden vectors from the reference CUDA build, the reference build live, and the CPU oracles
(oracle/pack_oracle.py, oracle/march_oracle.c).  Integer outputs (pack offsets, counts, indices, selectors) must
be bit-exact; float outputs within 1e-5 relative (t_starts / t_ends of the marcher: bit-exact).
"""
import numpy as np
import pytest
import torch

from tests.util import golden, load_ref, march_inputs, pack_inputs, rel_err

pytestmark = pytest.mark.gpu


def _pk():
    from nr3d_lib_b200.bindings import _pack_ops
    return _pack_ops


def _og():
    from nr3d_lib_b200.bindings import _occ_grid
    return _occ_grid


def _close(a, b, tol=1e-5, what=""):
    a, b = torch.as_tensor(a), torch.as_tensor(b)
    assert tuple(a.shape) == tuple(b.shape), (what, a.shape, b.shape)
    if a.dtype in (torch.bool, torch.uint8, torch.int32, torch.int64) or b.dtype in (torch.bool, torch.int32, torch.int64):
        assert torch.equal(a.cpu().long(), b.cpu().long()), f"{what}: integer outputs differ"
    else:
        e = rel_err(a.cpu(), b.cpu())
        assert e <= tol, f"{what}: rel err {e} > {tol}"


def _pack_all(be, d, dev, skip_exclusive_cumprod=False):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi = t(d["pack_infos"])
    f1, fC, pr, ap, apC = t(d["feats1"]), t(d["featsC"]), t(d["prod1"]), t(d["other1"]), t(d["otherC"])
    out = {"sum1": be.packed_sum(f1, pi), "sumC": be.packed_sum(fC, pi)}
    for ex in (0, 1):
        for rv in (0, 1):
            out[f"cumsum1_e{ex}r{rv}"] = be.packed_cumsum(f1, pi, bool(ex), bool(rv))
            out[f"cumsumC_e{ex}r{rv}"] = be.packed_cumsum(fC, pi, bool(ex), bool(rv))
            if not (ex and skip_exclusive_cumprod):
                out[f"cumprod1_e{ex}r{rv}"] = be.packed_cumprod(pr, pi, bool(ex), bool(rv))
    out["diff1"], out["diffC"] = be.packed_diff(f1, pi, None, None), be.packed_diff(fC, pi, None, None)
    out["diff1_append"], out["diffC_fill"] = be.packed_diff(f1, pi, ap, None), be.packed_diff(fC, pi, None, apC)
    out["bdiff1"], out["bdiffC_prepend"] = be.packed_backward_diff(f1, pi, None, None), be.packed_backward_diff(fC, pi, apC, None)
    out["bdiff1_fill"] = be.packed_backward_diff(f1, pi, None, ap)
    for nm in ("add", "sub", "mul", "div", "gt", "geq", "lt", "leq", "eq", "neq"):
        out[f"{nm}1"] = getattr(be, "packed_" + nm)(f1, ap, pi)
        out[f"{nm}C"] = getattr(be, "packed_" + nm)(fC, apC, pi)
    al, gw = t(d["alphas"]), t(d["grad_w"])
    for tag, eps, thre in (("a", 1e-4, 0.0), ("b", 0.3, 0.05)):
        w, _, _ = be.packed_alpha_to_vw_forward(al, pi, eps, thre, False)
        _, cpi, sel = be.packed_alpha_to_vw_forward(al, pi, eps, thre, True)
        out[f"vw_w_{tag}"], out[f"vw_cpi_{tag}"], out[f"vw_sel_{tag}"] = w, cpi, sel
        out[f"vw_ga_{tag}"] = be.packed_alpha_to_vw_backward(w, gw, al, pi, eps, thre)
    out["arange"], out["arange_idx"] = be.interleave_arange(t(d["n"]), True)
    out["linstep"], out["linstep_idx"] = be.interleave_linstep(t(d["near"]), t(d["n"]), t(d["other1"] * 0.01), True)
    out["ss_t"], out["ss_d"], out["ss_idx"], out["ss_pi"] = be.interleave_sample_step_wrt_depth_clamped(t(d["near"]), t(d["far"]), 64, 0.02, 0.01, 0.2)
    out["boundaries"] = be.mark_pack_boundaries_cuda(t(d["ids"]))
    return out


def test_pack_vs_golden(dev):
    g = golden("pack_ops")
    if g is None:
        pytest.skip("golden fixture not generated yet")
    mine = _pk()
    mine.EXCLUSIVE_CUMPROD_BUG_COMPAT = True   # the fixture records the reference CUDA output (all zeros, quirk Q2)
    try:
        got = _pack_all(mine, g | {"n": g["pack_infos"][:, 1].copy()}, dev)
    finally:
        mine.EXCLUSIVE_CUMPROD_BUG_COMPAT = False
    for k, v in got.items():
        _close(v, g[k], what="golden:" + k)


@pytest.mark.parametrize("seed,P,max_len,C", [(1, 300, 200, 4), (2, 4096, 96, 1), (3, 50, 3000, 2)])
def test_pack_vs_reference_build(seed, P, max_len, C, dev):
    ref = load_ref("_pack_ops")
    if ref is None:
        pytest.skip("oracle/_ref/_pack_ops.so not built")
    d = pack_inputs(P=P, max_len=max_len, C=C, seed=seed)
    want = _pack_all(ref, d, dev, skip_exclusive_cumprod=True)
    got = _pack_all(_pk(), d, dev, skip_exclusive_cumprod=True)
    for k, v in want.items():
        # long packs: fp32 scan order differs from the sequential loop
        _close(got[k], v, tol=3e-5 if ("cum" in k or "sum" in k) else 1e-5, what=f"ref:{k}")


@pytest.mark.parametrize("P,max_len,min_len", [(20000, 230, 1), (33, 5000, 2000), (4097, 40, 0)])
def test_staged_composite_bit_exact_vs_reference_build(P, max_len, min_len, dev):
    """The staged thread-per-pack kernels (csrc/pack_staged.cu) walk every pack in the reference thread's order: fp32 weights, selectors,
    counts, dL/dalpha and the one-channel pack sum are BIT-identical to the reference's CUDA build (and to the sequential numpy oracle)."""
    from oracle import pack_oracle as PO
    d = pack_inputs(P=P, max_len=max_len, C=1, seed=P, min_len=max(min_len, 1))
    if min_len == 0:      # zero-length packs in between (undefined in the reference: oracle only)
        n = d["n"].copy(); n[::7] = 0
        d["pack_infos"] = np.stack([np.cumsum(n) - n, n], 1)
        d["S"] = int(n.sum())
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi, al, gw = t(d["pack_infos"]), t(d["alphas"][: max(d["S"], 1)]), t(d["grad_w"][: max(d["S"], 1)])
    mine = _pk()
    ref = load_ref("_pack_ops") if min_len else None
    for eps, thre in ((1e-4, 0.0), (0.3, 0.01), (0.0, 0.0)):
        w, _, _ = mine.packed_alpha_to_vw_forward(al, pi, eps, thre, False)
        _, cpi, sel = mine.packed_alpha_to_vw_forward(al, pi, eps, thre, True)
        ga = mine.packed_alpha_to_vw_backward(w, gw, al, pi, eps, thre)
        sm = mine.packed_sum(w, pi)
        w_o, cnt_o, sel_o = PO.alpha_to_vw_forward(al.cpu().numpy(), d["pack_infos"], eps, thre)
        assert np.array_equal(w.cpu().numpy(), w_o) and np.array_equal(sel.cpu().numpy().astype(bool), sel_o.astype(bool))
        assert np.array_equal(cpi.cpu().numpy()[:, 1], cnt_o)
        if ref is not None:
            w_r, _, _ = ref.packed_alpha_to_vw_forward(al, pi, eps, thre, False)
            _, cpi_r, sel_r = ref.packed_alpha_to_vw_forward(al, pi, eps, thre, True)
            ga_r = ref.packed_alpha_to_vw_backward(w_r, gw, al, pi, eps, thre)
            assert torch.equal(w, w_r) and torch.equal(sel, sel_r) and torch.equal(cpi, cpi_r)
            assert torch.equal(ga, ga_r), f"dL/dalpha not bit-exact: max diff {(ga - ga_r).abs().max().item()}"
            assert torch.equal(sm, ref.packed_sum(w_r, pi)), "pack sum not bit-exact"
    if ref is not None:     # one-channel scans walk the pack like the reference thread: bit-identical (the exclusive product differs by design, Q2)
        pr = t((1.0 + 0.2 * np.random.RandomState(P).randn(max(d["S"], 1))).astype(np.float32))
        for ex in (False, True):
            for rv in (False, True):
                assert torch.equal(mine.packed_cumsum(gw, pi, ex, rv), ref.packed_cumsum(gw, pi, ex, rv)), ("cumsum", ex, rv)
                if not ex:
                    assert torch.equal(mine.packed_cumprod(pr, pi, ex, rv), ref.packed_cumprod(pr, pi, ex, rv)), ("cumprod", ex, rv)


def test_staged_composite_untiled_packs(dev):
    """pack_infos that do not tile one span (gaps, reversed order, overlapping packs): the staged kernels fall back to one pack per span and
    still reproduce the sequential oracle bit for bit; entries outside every pack keep the caller's zeros."""
    from oracle import pack_oracle as PO
    rs = np.random.RandomState(5)
    S = 9000
    al = np.clip(rs.rand(S) ** 2 * 0.6, 0, 0.999).astype(np.float32)
    begins = np.sort(rs.choice(S - 200, size=70, replace=False)).astype(np.int64)
    lens = np.minimum(rs.randint(0, 120, size=70), np.diff(np.append(begins, S))).astype(np.int64)     # disjoint, with gaps
    order = rs.permutation(70)
    pinf = np.stack([begins, lens], 1)[order]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    mine = _pk()
    w, _, _ = mine.packed_alpha_to_vw_forward(t(al), t(pinf), 1e-3, 0.0, False)
    w_o, _, _ = PO.alpha_to_vw_forward(al, pinf, 1e-3, 0.0)
    assert np.array_equal(w.cpu().numpy(), w_o)
    gw = rs.randn(S).astype(np.float32)
    ga = mine.packed_alpha_to_vw_backward(w, t(gw), t(al), t(pinf), 1e-3, 0.0)
    ga_o = PO.alpha_to_vw_backward(w_o, gw, al, pinf, 1e-3, 0.0)
    _close(ga, ga_o, 1e-6, "untiled alpha bwd")
    assert np.array_equal(mine.packed_sum(t(al), t(pinf)).cpu().numpy(), np.array([al[b:b + n].astype(np.float32).cumsum(dtype=np.float32)[-1] if n else 0.0
                                                                                     for b, n in pinf], dtype=np.float32))


def test_pack_vs_oracle_with_empty_packs(dev):
    """oracle comparison incl. zero-length packs (undefined behaviour in the reference, defined as no-ops here)."""
    from oracle import pack_oracle as PO
    mine = _pk()
    d = pack_inputs(P=64, max_len=90, C=3, seed=4, min_len=0)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi = t(d["pack_infos"])
    _close(mine.packed_sum(t(d["featsC"]), pi), PO.packed_sum(d["featsC"], d["pack_infos"]), what="sum")
    for ex in (False, True):
        for rv in (False, True):
            _close(mine.packed_cumsum(t(d["feats1"]), pi, ex, rv), PO.packed_cumsum(d["feats1"], d["pack_infos"], ex, rv), 3e-5, "cumsum")
            _close(mine.packed_cumprod(t(d["prod1"]), pi, ex, rv), PO.packed_cumprod(d["prod1"], d["pack_infos"], ex, rv), 3e-5, "cumprod")
    _close(mine.packed_diff(t(d["featsC"]), pi, None, t(d["otherC"])), PO.packed_diff(d["featsC"], d["pack_infos"], None, d["otherC"]), what="diff")
    _close(mine.packed_backward_diff(t(d["feats1"]), pi, t(d["other1"]), None), PO.packed_backward_diff(d["feats1"], d["pack_infos"], d["other1"], None), what="bdiff")
    for eps, thre in ((1e-4, 0.0), (0.2, 0.1)):
        w, cnt, sel = PO.alpha_to_vw_forward(d["alphas"], d["pack_infos"], eps, thre)
        gw_, _, _ = mine.packed_alpha_to_vw_forward(t(d["alphas"]), pi, eps, thre, False)
        assert torch.equal(gw_.cpu(), torch.from_numpy(w)), "alpha weights must be bit-exact (sequential product order)"
        _, cpi, gsel = mine.packed_alpha_to_vw_forward(t(d["alphas"]), pi, eps, thre, True)
        assert torch.equal(cpi[:, 1].cpu().long(), torch.from_numpy(cnt)) and torch.equal(gsel.cpu(), torch.from_numpy(sel))
        ga = PO.alpha_to_vw_backward(w, d["grad_w"], d["alphas"], d["pack_infos"], eps, thre)
        _close(mine.packed_alpha_to_vw_backward(gw_, t(d["grad_w"]), t(d["alphas"]), pi, eps, thre), ga, 2e-5, "alpha bwd")
    # documented exclusive cumprod == kaolin's known-answer semantics (leading 1)
    e = mine.packed_cumprod(t(d["prod1"]), pi, True, False).cpu().numpy()
    firsts = d["pack_infos"][d["pack_infos"][:, 1] > 0, 0]
    assert np.all(e[firsts] == 1.0)
    # float64 / int64 dtypes
    _close(mine.packed_cumsum(t(d["feats1"].astype(np.float64)), pi, False, False), PO.packed_cumsum(d["feats1"].astype(np.float64), d["pack_infos"]), 1e-12, "f64")
    ints = (d["feats1"] * 10).astype(np.int64)
    _close(mine.packed_sum(t(ints), pi), PO.packed_sum(ints, d["pack_infos"]), what="i64 sum")


def test_pack_autograd_wrappers(dev):
    """Gradients of the host-side mirror vs torch reference formulas (the reference's own check style, unit_test.py:188-241)."""
    from nr3d_lib_b200 import pack_ops as P
    d = pack_inputs(P=40, max_len=50, C=2, seed=6)
    pi = torch.from_numpy(d["pack_infos"]).to(dev)
    n = pi[:, 1]
    f = torch.from_numpy(d["featsC"]).to(dev).double().requires_grad_(True)
    o = torch.from_numpy(d["otherC"]).to(dev).double().requires_grad_(True)
    for op, tf in ((P.packed_add, lambda a, b: a + b), (P.packed_sub, lambda a, b: a - b), (P.packed_mul, lambda a, b: a * b),
                   (P.packed_div, lambda a, b: a / b)):
        w = torch.randn_like(f)
        g1 = torch.autograd.grad((op(f, o, pi) * w).sum(), [f, o])
        g2 = torch.autograd.grad((tf(f, o.repeat_interleave(n, 0)) * w).sum(), [f, o])
        assert rel_err(g1[0], g2[0]) < 1e-10 and rel_err(g1[1], g2[1]) < 1e-10
    a = torch.from_numpy(d["alphas"]).to(dev).double().clamp(1e-3, 0.9).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: P.packed_alpha_to_vw(t, pi, 1e-9, 0.0), (a,), eps=1e-6, atol=1e-6, rtol=1e-4)
    s = torch.from_numpy(d["feats1"]).to(dev).double().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: P.packed_cumsum(t, pi, False, False), (s,), eps=1e-6, atol=1e-7)
    assert torch.autograd.gradcheck(lambda t: P.packed_sum(t, pi), (s,), eps=1e-6, atol=1e-7)
    assert torch.autograd.gradcheck(lambda t: P.packed_diff(t, pi), (s,), eps=1e-6, atol=1e-7)
    assert torch.autograd.gradcheck(lambda t: P.packed_backward_diff(t, pi), (s,), eps=1e-6, atol=1e-7)


# ------------------------------------------------------------------------------------------------------------------
# marcher
# ------------------------------------------------------------------------------------------------------------------
def _march(be, d, dev, ct, step, mx, gamma, ms, bds=None):
    t = lambda a: None if a is None else torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    if d["grid"].ndim == 4:
        r = be.batched_ray_marching(t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), t(d.get("batch_inds")) if not bds else None,
                                    bds, t(d["roi"]), t(d["grid"]), be.ContractionType(ct), step, mx, gamma, ms, True)
        return dict(packed_info=r[0], t_starts=r[1], t_ends=r[2], ridx=r[3], bidx=r[4], gidx=r[5])
    r = be.ray_marching(t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), t(d["roi"]), t(d["grid"]), be.ContractionType(ct),
                        step, mx, gamma, ms, True)
    return dict(packed_info=r[0], t_starts=r[1], t_ends=r[2], ridx=r[3], gidx=r[4])


def _march_equal(got, want, what, exact_float=True):
    for k in ("packed_info", "ridx", "bidx", "gidx"):
        if k in want and want[k] is not None and k in got:
            w = torch.as_tensor(want[k]).cpu().long().flatten()
            g = got[k].cpu().long().flatten()
            assert g.shape == w.shape and torch.equal(g, w), f"{what}: {k} differs"
    for k in ("t_starts", "t_ends"):
        w, g = torch.as_tensor(want[k]).cpu().flatten(), got[k].cpu().flatten()
        if exact_float:
            assert torch.equal(g, w), f"{what}: {k} not bit-exact (max diff {(g - w).abs().max().item() if g.numel() else 0})"
        else:
            assert rel_err(g, w) < 1e-6


MARCH_GOLDEN = ["aabb", "aabb_gamma", "aabb_shell", "aabb_maxsteps", "sphere", "tanh", "batched_inds", "batched_size"]


@pytest.fixture(params=["single_pass", "two_pass"])
def march_mode(request, monkeypatch):
    """Both marchers: record + compact (every ray marched once; default when the scratch fits) and count + fill (the reference's two passes)."""
    if request.param == "two_pass":
        monkeypatch.setattr(_og(), "MARCH_SCRATCH_BYTES", 0)
    return request.param


@pytest.mark.parametrize("name", MARCH_GOLDEN)
def test_march_vs_golden(name, dev, march_mode):
    g = golden("march_" + name)
    if g is None:
        pytest.skip("golden fixture not generated yet")
    cfg = g["cfg"]
    bds = int(cfg[5]) if len(cfg) > 5 and cfg[5] else None
    got = _march(_og(), g, dev, int(cfg[0]), float(cfg[1]), float(cfg[2]), float(cfg[3]), int(cfg[4]), bds)
    assert got["t_starts"].shape[1:] == (1,) and got["packed_info"].dtype == torch.int32
    _march_equal(got, g, "golden:" + name)


@pytest.mark.parametrize("R,res,B", [(20000, 64, 1), (8192, 32, 4), (33, 32, 1)])
def test_march_vs_reference_build_and_oracle(R, res, B, dev, march_mode):
    from oracle import march_oracle as MO
    d = march_inputs(R=R, res=res, seed=21, B=B)
    got = _march(_og(), d, dev, 0, 0.005, 1e10, 0.0, 512)
    o = MO.ray_marching(d["rays_o"], d["rays_d"], d["near"], d["far"], d["roi"], d["grid"], 0, 0.005, 1e10, 0.0, 512, batch_inds=d["batch_inds"])
    _march_equal(got, o, "oracle")
    ref = load_ref("_occ_grid")
    if ref is not None:
        want = _march(ref, d, dev, 0, 0.005, 1e10, 0.0, 512)
        _march_equal(got, want, "ref")


def test_march_wrapper_and_edges(dev):
    from nr3d_lib_b200.occgrid_raymarch import occgrid_raymarch, occgrid_raymarch_batched
    d = march_inputs(R=1000, res=32, seed=8)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ret = occgrid_raymarch(t(d["grid"]), t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), step_size=0.01)
    assert ret.num_hit_rays > 0 and ret.pack_infos.dtype == torch.int64 and ret.samples.shape == (ret.ridx.numel(), 3)
    assert int(ret.pack_infos[:, 1].sum()) == ret.ridx.numel() and (ret.deltas > 0).all()
    # sortedness / pack structure: ridx is non-decreasing and constant within each pack
    assert (ret.ridx[1:] >= ret.ridx[:-1]).all()
    # empty grid -> no hits
    ret0 = occgrid_raymarch(torch.zeros(16, 16, 16, dtype=torch.bool, device=dev), t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), step_size=0.01)
    assert ret0.num_hit_rays == 0 and ret0.samples is None
    # zero rays
    z = torch.zeros(0, 3, device=dev)
    r = _og().ray_marching(z, z, torch.zeros(0, device=dev), torch.zeros(0, device=dev), torch.tensor([-1., -1, -1, 1, 1, 1], device=dev),
                           t(d["grid"]), _og().ContractionType.AABB, 0.01, 1e10, 0.0, 64, True)
    assert r[0].shape == (0, 2) and r[1].shape == (0, 1)
    # batched with negative batch indices: skipped rays get zero samples (defined behaviour, unlike the reference)
    db = march_inputs(R=600, res=16, seed=9, B=2)
    bi = db["batch_inds"].copy()
    bi[::3] = -1
    rb = occgrid_raymarch_batched(t(db["grid"]), t(db["rays_o"]), t(db["rays_d"]), t(bi), t(db["near"]), t(db["far"]), step_size=0.02)
    assert rb.num_hit_rays > 0 and not np.isin(rb.ridx_hit.cpu().numpy(), np.arange(0, 600, 3)).any()
