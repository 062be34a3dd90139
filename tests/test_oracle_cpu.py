"""CPU tests (no GPU): pin the oracles against reference-owned data.

* golden vectors recorded from the reference's own CUDA build on a B200 (tests/golden/*.npz, make_golden.py);
* the reference's compiled LoDMeta (oracle/_ref/_lotd.so is host code for this class) when present;
* the reference's pure-PyTorch Dense path (F.grid_sample, lotd_helpers.py:327-346);
* the literal known-answer vectors of the reference's pack_ops unit test (unit_test.py:547-563);
* finite-difference identities (the reference's own LoTD test strategy, lotd/tests/math_test.py:99-171).
"""
import numpy as np
import pytest
import torch

from oracle import lotd_oracle as O
from oracle import pack_oracle as PO
from tests.util import LOTD_CONFIGS, golden, load_ref, lotd_inputs, march_inputs, meta_args, rel_err

META_ATTRS = ("level_res level_res_multidim level_n_params level_n_feats level_types level_sizes level_offsets map_levels map_cnt "
              "n_levels n_pseudo_levels n_feat_per_pseudo_lvl n_dims_to_encode n_encoded_dims n_params").split()


def _ngp16():
    res = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
    return (3, res, [2] * 16, ["Dense" if r ** 3 <= 2 ** 19 else "Hash" for r in res], 2 ** 19, False)


def test_meta_headline_config():
    """C2 of BASELINE.json: gen_ngp_cfg defaults (lotd_cfg.py:48-57) -> 12,131,648 params, 32 features."""
    m = O.OracleMeta(*_ngp16())
    assert m.level_res[:6] == [16, 22, 30, 42, 58, 80] and m.level_res[-1] == 2049
    assert m.level_types == [0] * 6 + [7] * 10
    assert m.n_params == 12131648 and m.n_encoded_dims == 32 and m.n_pseudo_levels == 16 and m.c_hash_only


@pytest.mark.parametrize("name", list(LOTD_CONFIGS) + ["ngp16"])
def test_meta_matches_shim_and_reference(name):
    args = _ngp16() if name == "ngp16" else meta_args(LOTD_CONFIGS[name])
    from nr3d_lib_b200.bindings import _lotd as shim   # host-only call into libnr3d_b200.so (no GPU needed)
    o, s = O.OracleMeta(*args), shim.LoDMeta(*args)
    ref = load_ref("_lotd")
    r = ref.LoDMeta(*args) if ref is not None else None
    for a in META_ATTRS:
        assert getattr(o, a) == getattr(s, a), a
        if r is not None:
            assert list(getattr(r, a)) == getattr(o, a) if isinstance(getattr(o, a), list) else getattr(r, a) == getattr(o, a), a
    assert o.c_hash_only == s.c_hash_only and int(s.interpolation_type) == o.interpolation_type


def test_meta_errors():
    from nr3d_lib_b200.bindings import _lotd as shim
    for cls in (O.OracleMeta, shim.LoDMeta):
        with pytest.raises(RuntimeError):
            cls(3, [2, 8], [2, 2], ["Dense", "Dense"])               # res must be >= 3
        with pytest.raises(RuntimeError):
            cls(3, [8], [3], ["Dense"])                              # odd feature width
        with pytest.raises(RuntimeError):
            cls(3, [8], [2], ["Hash"])                               # hash without hashmap_size
        with pytest.raises(RuntimeError):
            cls(2, [8], [2], ["VM"])                                 # VM is 3-D only
        with pytest.raises(RuntimeError):
            cls(5, [8], [2], ["Dense"])
        with pytest.raises(RuntimeError):
            cls(3, [8], [2], ["Bogus"])
    assert shim.LoDMeta(3, [8], [2], ["vm"]).level_types == [1] and shim.LoDMeta(3, [8], [2], ["NPlane"]).level_types == [6]


def test_dense_level_matches_reference_grid_sample():
    """Dense coordinate convention (quirk Q1) against the reference's own F.grid_sample path (config C1 of BASELINE.json)."""
    torch.manual_seed(0)
    R, F = 32, 4
    m = O.OracleMeta(3, [R], [F], ["Dense"])
    p = torch.randn(R ** 3 * F) * 1e-2
    x = torch.rand(4096, 3).clamp(1e-6, 1 - 1e-6)
    y = O.encode(m, x, p)
    y_ref = O.grid_sample_dense_reference(p.view(R, R, R, F).double(), x.double(), R)
    assert rel_err(y, y_ref) < 1e-5
    # param cycle equivalence (lotd_helpers.py:451-484): sampling at the vertices returns the parameters
    idx = torch.stack(torch.meshgrid(*[torch.arange(R)] * 3, indexing="ij"), -1).reshape(-1, 3)
    xv = ((idx.double() - 0.5) / (R - 2)).clamp(1e-6, 1 - 1e-6)
    inner = ((idx > 0) & (idx < R - 1)).all(-1)
    yv = O.encode(m, xv.float(), p)
    assert rel_err(yv[inner], p.view(-1, F).double()[inner]) < 1e-4


def _golden_inputs(g):
    return dict(x=torch.from_numpy(g["x"]), params=torch.from_numpy(g["params"]).float(), dL_dy=torch.from_numpy(g["dL_dy"]).float(),
                dL_ddLdx=torch.from_numpy(g["dL_ddLdx"]), batch_inds=torch.from_numpy(g["batch_inds"]) if "batch_inds" in g else None)


@pytest.mark.parametrize("name", list(LOTD_CONFIGS))
def test_lotd_oracle_vs_golden(name):
    """The float64 oracle reproduces what the reference's CUDA build produced on a B200 (fp32 params: 1e-5)."""
    from tests.test_lotd_gpu import _compare
    g = golden(f"lotd_{name}_f32")
    if g is None:
        pytest.skip("golden fixture missing")
    cfg = LOTD_CONFIGS[name]
    om = O.OracleMeta(*meta_args(cfg))
    inp = _golden_inputs(g)
    kw = dict(batch_inds=inp["batch_inds"])
    y, dydx = O.fwd_dydx(om, inp["x"], inp["params"], **kw)
    gx, gp = O.bwd(om, inp["dL_dy"], inp["x"], inp["params"], **kw)
    g_gy, g_p2, g_x2 = O.bwd_bwd_input(om, inp["dL_ddLdx"], inp["dL_dy"], inp["x"], inp["params"], **kw)
    got = dict(y=y, dy_dx=dydx, dL_dx=gx, dL_dparam=gp, dL_ddLdy=g_gy, dL_dparam2=g_p2, dL_dx2=g_x2,
               y_maxlevel1=O.encode(om, inp["x"], inp["params"], max_level=1, **kw))
    if om.c_hash_only:
        got["grid_index"] = O.grid_index(om, inp["x"], **kw)
    want = {k: g[k] for k in got if k in g}

    if not om.c_hash_only and cfg["D"] >= 3:
        assert "checker_build" in g, "fixture predates the -G checker pass (tests/golden/make_golden.py:make_lotd_checker)"
    _compare(got, want, torch.float32, f"oracle-vs-golden:{name}")      # every output, nothing masked


@pytest.mark.parametrize("name", ["mixed", "mixed_smooth", "cuboid_vm", "d4", "batched"])
def test_lotd_oracle_gradients_vs_finite_differences(name):
    """Analytic (autograd) derivatives of the oracle vs central differences in float64 -- pins dy/dx and the second-order
    products independently of any build of the reference (its -O3 build of the generic kernels is miscompiled there, tests/test_lotd_gpu.py)."""
    cfg = LOTD_CONFIGS[name]
    om = O.OracleMeta(*meta_args(cfg))
    inp = lotd_inputs(cfg, om.n_params, N=64, seed=2)
    x = inp["x"].clamp(0.02, 0.98)
    keep = torch.ones(x.shape[0], dtype=torch.bool)
    h = 1e-5
    for R in om.level_res_multidim:
        s = torch.tensor([r - 2 for r in R], dtype=torch.float64)
        keep &= (torch.floor((x.double() + 4 * h) * s + 0.5) == torch.floor((x.double() - 4 * h) * s + 0.5)).all(-1)
    x = x[keep]
    bi = None if inp["batch_inds"] is None else inp["batch_inds"][keep]
    p = inp["params"].double()
    kw = dict(batch_inds=bi)

    def f64_encode(xq, pq):
        # evaluate the same piecewise polynomial at float64 inputs (cells fixed by the unperturbed fp32 x)
        return O.encode(om, xq, pq, **kw)

    _, dydx = O.fwd_dydx(om, x, p, **kw)
    if not cfg["smooth"]:
        # dy/dx vs central differences of the oracle's own forward (exact inside a cell for multilinear interpolation)
        hx = 1e-3
        ok = torch.ones(x.shape[0], dtype=torch.bool)
        for R in om.level_res_multidim:
            s_ = torch.tensor([r - 2 for r in R], dtype=torch.float64)
            ok &= (torch.floor((x.double() + 2 * hx) * s_ + 0.5) == torch.floor((x.double() - 2 * hx) * s_ + 0.5)).all(-1)
        for d in range(x.shape[1]):
            e = torch.zeros_like(x)
            e[:, d] = hx
            xp, xm = (x + e), (x - e)
            fd = (O.encode(om, xp, p, **kw) - O.encode(om, xm, p, **kw)) / (xp[:, d:d + 1].double() - xm[:, d:d + 1].double())
            # the value path rounds x*scale+0.5 to float32 (as the kernels do): ~1e-6 noise in the fraction / 2e-3 step
            assert rel_err(dydx[ok][:, :, d], fd[ok]) < 2e-4, (name, d)
    # d/dparams by finite differences (exact for multilinear-in-params types up to products)
    w = inp["dL_dy"][keep].double()
    _, gp = O.bwd(om, w, x, p, **kw)
    rs = np.random.RandomState(0)
    idxs = rs.choice(np.nonzero(gp.abs().numpy() > 0)[0], size=12, replace=False)
    for i in idxs:
        dp = torch.zeros_like(p)
        dp[i] = 1e-4
        fd = ((f64_encode(x, p + dp) - f64_encode(x, p - dp)) * w).sum() / 2e-4
        assert abs(fd.item() - gp[i].item()) <= 1e-6 * max(1.0, abs(gp[i].item())), (i, fd.item(), gp[i].item())
    # second order: d/dparams and d/d(dL_dy) of <dL_dx, u> by finite differences of the first-order oracle
    u = inp["dL_ddLdx"][keep].double()
    g_gy, g_p2, _ = O.bwd_bwd_input(om, u, w, x, p, **kw)
    for i in idxs[:6]:
        dp = torch.zeros_like(p)
        dp[i] = 1e-4
        fd = ((O.bwd(om, w, x, p + dp, **kw)[0] - O.bwd(om, w, x, p - dp, **kw)[0]) * u).sum() / 2e-4
        assert abs(fd.item() - g_p2[i].item()) <= 1e-6 * max(1.0, abs(g_p2[i].item())), (i, fd.item(), g_p2[i].item())
    dw = torch.zeros_like(w)
    dw[0, 0] = 1.0
    fd = ((O.bwd(om, w + dw, x, p, **kw)[0] - O.bwd(om, w - dw, x, p, **kw)[0]) * u).sum() / 2.0
    assert abs(fd.item() - g_gy[0, 0].item()) <= 1e-8 * max(1.0, abs(fd.item()))
    # dy/dx consistency: dL_dx = sum_j dL_dy * dy_dx
    gx, _ = O.bwd(om, w, x, p, **kw)
    assert rel_err((dydx * w.unsqueeze(-1)).sum(1), gx) < 1e-12


# ------------------------------------------------------------------------------------------------------------------
# pack ops
# ------------------------------------------------------------------------------------------------------------------
def test_pack_known_answers_from_reference_unit_test():
    """Literal vectors printed in nr3d_lib/graphics/pack_ops/unit_test.py:547-563 (kaolin semantics of `exclusive`)."""
    feat = np.array([0.8750, 0.0581, 0.9378, 0.9638, 0.9859, 0.4652, 0.9105, 0.5071, 0.0173, 0.6071, 0.7123, 0.7371, 0.8094], dtype=np.float32)
    incl = np.array([0.8750, 0.9331, 1.8709, 2.8347, 0.9859, 1.4512, 2.3617, 2.8688, 2.8860, 3.4931, 4.2054, 0.7371, 1.5465])
    excl = np.array([0.0000, 0.8750, 0.9331, 1.8709, 0.0000, 0.9859, 1.4512, 2.3617, 2.8688, 2.8860, 3.4931, 0.0000, 0.7371])
    pi = PO.pack_infos_from_counts([4, 7, 2])
    assert np.allclose(PO.packed_cumsum(feat, pi), incl, atol=2e-4) and np.allclose(PO.packed_cumsum(feat, pi, exclusive=True), excl, atol=2e-4)
    feat = np.array([-0.8033, 1.2413, 0.1971, 1.2183, 1.3434, 1.7485, 0.0624, -0.3419, -0.1997, -1.4790, 1.1720, 0.1686, -0.0704], dtype=np.float32)
    incl = np.array([-0.8033, -0.9972, -0.1965, -0.2394, 1.3434, 2.3489, 0.1465, -0.0501, 0.0100, -0.0148, -0.0173, 0.1686, -0.0119])
    excl = np.array([1.0000, -0.8033, -0.9972, -0.1965, 1.0000, 1.3434, 2.3489, 0.1465, -0.0501, 0.0100, -0.0148, 1.0000, 0.1686])
    assert np.allclose(PO.packed_cumprod(feat, pi), incl, atol=2e-4) and np.allclose(PO.packed_cumprod(feat, pi, exclusive=True), excl, atol=2e-4)
    assert np.all(PO.packed_cumprod(feat, pi, exclusive=True, bug_compat=True) == 0)  # what the reference CUDA kernel returns (Q2)


def test_pack_oracle_vs_golden():
    g = golden("pack_ops")
    if g is None:
        pytest.skip("golden fixture missing")
    pi = g["pack_infos"]
    close = lambda a, b, tol=3e-5: rel_err(a.astype(np.float64), b.astype(np.float64)) <= tol
    assert close(PO.packed_sum(g["feats1"], pi), g["sum1"]) and close(PO.packed_sum(g["featsC"], pi), g["sumC"])
    for ex in (0, 1):
        for rv in (0, 1):
            assert np.array_equal(PO.packed_cumsum(g["feats1"], pi, bool(ex), bool(rv)), g[f"cumsum1_e{ex}r{rv}"])   # same sequential order
            assert np.array_equal(PO.packed_cumsum(g["featsC"], pi, bool(ex), bool(rv)), g[f"cumsumC_e{ex}r{rv}"])
            assert np.array_equal(PO.packed_cumprod(g["prod1"], pi, bool(ex), bool(rv), bug_compat=True), g[f"cumprod1_e{ex}r{rv}"])
    assert np.array_equal(PO.packed_diff(g["feats1"], pi), g["diff1"]) and np.array_equal(PO.packed_diff(g["featsC"], pi), g["diffC"])
    assert np.array_equal(PO.packed_diff(g["feats1"], pi, appends=g["other1"]), g["diff1_append"])
    assert np.array_equal(PO.packed_diff(g["featsC"], pi, last_fill=g["otherC"]), g["diffC_fill"])
    assert np.array_equal(PO.packed_backward_diff(g["feats1"], pi), g["bdiff1"])
    assert np.array_equal(PO.packed_backward_diff(g["featsC"], pi, prepends=g["otherC"]), g["bdiffC_prepend"])
    assert np.array_equal(PO.packed_backward_diff(g["feats1"], pi, first_fill=g["other1"]), g["bdiff1_fill"])
    for op, nm in ((0, "add"), (1, "sub"), (2, "mul"), (3, "div"), (5, "gt"), (6, "geq"), (7, "lt"), (8, "leq"), (9, "eq"), (10, "neq")):
        assert np.array_equal(PO.packed_binary(op, g["feats1"], g["other1"], pi), g[nm + "1"]), nm
        assert np.array_equal(PO.packed_binary(op, g["featsC"], g["otherC"], pi), g[nm + "C"]), nm
    for tag, eps, thre in (("a", 1e-4, 0.0), ("b", 0.3, 0.05)):
        w, cnt, sel = PO.alpha_to_vw_forward(g["alphas"], pi, eps, thre)
        assert np.array_equal(w, g[f"vw_w_{tag}"]) and np.array_equal(sel, g[f"vw_sel_{tag}"])
        assert np.array_equal(PO.pack_infos_from_counts(cnt), g[f"vw_cpi_{tag}"].astype(np.int64))
        ga = PO.alpha_to_vw_backward(w, g["grad_w"], g["alphas"], pi, eps, thre)
        assert close(ga, g[f"vw_ga_{tag}"], 1e-6)
    n = pi[:, 1]
    ar, ar_idx = PO.interleave_linstep(np.zeros(len(n), dtype=np.int64), n, 1)
    assert np.array_equal(ar, g["arange"]) and np.array_equal(ar_idx, g["arange_idx"])
    ls, ls_idx = PO.interleave_linstep(g["near"], n, g["other1"] * np.float32(0.01))
    assert np.array_equal(ls, g["linstep"]) and np.array_equal(ls_idx, g["linstep_idx"])
    ts, ds, ni, spi = PO.sample_step_wrt_depth_clamped(g["near"], g["far"], 64, 0.02, 0.01, 0.2)
    assert np.array_equal(spi, g["ss_pi"]) and np.array_equal(ts, g["ss_t"]) and np.array_equal(ds, g["ss_d"]) and np.array_equal(ni, g["ss_idx"])
    assert np.array_equal(PO.mark_pack_boundaries(g["ids"]), g["boundaries"])


def test_pack_torch_equivalences():
    """pure-torch equivalents used by the reference's own unit test (unit_test.py:188-241) and nerf_utils.py:98-110."""
    from tests.util import pack_inputs
    d = pack_inputs(P=20, max_len=30, C=2, seed=3)
    pi, n = d["pack_infos"], d["pack_infos"][:, 1]
    rep = torch.from_numpy(d["otherC"]).repeat_interleave(torch.from_numpy(n), 0).numpy()
    assert np.array_equal(PO.packed_binary(0, d["featsC"], d["otherC"], pi), d["featsC"] + rep)
    assert np.array_equal(PO.packed_binary(3, d["featsC"], d["otherC"], pi), d["featsC"] / rep)
    # alpha -> weights against the dense shifted-cumprod formulation of ray_alpha_to_vw
    w, _, _ = PO.alpha_to_vw_forward(d["alphas"], pi, 0.0, -1.0)
    for (b, k) in pi:
        a = torch.from_numpy(d["alphas"][b:b + k]).double()
        ref = a * torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1.0 - a[:-1]]), 0)
        assert rel_err(w[b:b + k], ref) < 1e-6


# ------------------------------------------------------------------------------------------------------------------
# marcher
# ------------------------------------------------------------------------------------------------------------------
MARCH_GOLDEN = ["aabb", "aabb_gamma", "aabb_shell", "aabb_maxsteps", "batched_inds", "batched_size"]


@pytest.mark.parametrize("name", MARCH_GOLDEN + ["sphere", "tanh"])
def test_march_oracle_vs_golden(name):
    """The C oracle reproduces the reference CUDA build bit for bit: pack offsets, counts, voxel indices, t_starts / t_ends."""
    from oracle import march_oracle as MO
    g = golden("march_" + name)
    if g is None:
        pytest.skip("golden fixture missing")
    cfg = g["cfg"]
    bds = int(cfg[5]) if len(cfg) > 5 else 0
    o = MO.ray_marching(g["rays_o"], g["rays_d"], g["near"], g["far"], g["roi"], g["grid"], int(cfg[0]), float(cfg[1]), float(cfg[2]),
                        float(cfg[3]), int(cfg[4]), batch_inds=g.get("batch_inds"), batch_data_size=bds)
    if name in ("sphere", "tanh"):
        # libm tanhf/sqrtf vs CUDA's: sample counts may differ by a few at voxel borders; require >= 99.5 % identical rays
        same = (o["packed_info"][:, 1] == g["packed_info"][:, 1]).mean()
        assert same > 0.995, same
        return
    assert np.array_equal(o["packed_info"], g["packed_info"])
    assert np.array_equal(o["t_starts"], g["t_starts"][:, 0]) and np.array_equal(o["t_ends"], g["t_ends"][:, 0])
    assert np.array_equal(o["ridx"], g["ridx"]) and np.array_equal(o["gidx"], g["gidx"])
    if "bidx" in g:
        assert np.array_equal(o["bidx"], g["bidx"])


def test_march_oracle_properties():
    from oracle import march_oracle as MO
    d = march_inputs(R=300, res=16, seed=5)
    o = MO.ray_marching(d["rays_o"], d["rays_d"], d["near"], d["far"], d["roi"], d["grid"], 0, 0.02, 1e10, 0.0, 64)
    pi = o["packed_info"]
    assert pi[0, 0] == 0 and np.array_equal(pi[1:, 0], np.cumsum(pi[:-1, 1])) and pi[:, 1].max() <= 64
    assert np.all(o["t_ends"] > o["t_starts"]) and np.all(np.diff(o["ridx"]) >= 0)
    # every sample's midpoint lies in an occupied voxel
    mid = 0.5 * (o["t_starts"] + o["t_ends"])
    p = d["rays_o"][o["ridx"]] + mid[:, None] * d["rays_d"][o["ridx"]]
    ijk = np.clip(((p + 1) / 2 * 16).astype(int), 0, 15)
    assert d["grid"][ijk[:, 0], ijk[:, 1], ijk[:, 2]].mean() > 0.999
    empty = MO.ray_marching(d["rays_o"], d["rays_d"], d["near"], d["far"], d["roi"], np.zeros_like(d["grid"]), 0, 0.02, 1e10, 0.0, 64)
    assert empty["packed_info"][:, 1].sum() == 0


# ------------------------------------------------------------------------------------------------------------------
# plain-C fp32 port of the Dense/Hash kernels (oracle/lotd_port.c): the CPU baseline of bench.py and a second, op-order
# faithful oracle.  Pinned by the reference build's golden vectors and by the float64 oracle.
# ------------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["ngp8", "ngp_smooth", "hash_f4"])
def test_lotd_port_vs_golden_and_oracle(name):
    from oracle import lotd_oracle as O, lotd_port as P
    cfg = LOTD_CONFIGS[name]
    om = O.OracleMeta(*meta_args(cfg))
    g = golden(f"lotd_{name}_f32")
    assert g is not None
    y = P.fwd(om, g["x"], g["params"])
    gp = P.bwd_param(om, g["dL_dy"], g["x"])
    assert rel_err(y, g["y"]) < 1e-6 and rel_err(gp, g["dL_dparam"]) < 1e-6           # the reference's own CUDA outputs
    inp = lotd_inputs(cfg, om.n_params, N=3000, seed=17)
    y2 = P.fwd(om, inp["x"].numpy(), inp["params"].numpy(), n_threads=3)
    assert rel_err(y2, O.encode(om, inp["x"], inp["params"])) < 1e-6
    gp2 = P.bwd_param(om, inp["dL_dy"].numpy(), inp["x"].numpy(), n_threads=3)
    assert rel_err(gp2, O.bwd(om, inp["dL_dy"], inp["x"], inp["params"])[1]) < 1e-5
    assert abs(float((y2.astype(np.float64) * inp["dL_dy"].numpy()).sum() - (inp["params"].numpy().astype(np.float64) * gp2).sum())) < 1e-3   # adjoint
    with pytest.raises(ValueError):
        P.fwd(O.OracleMeta(*meta_args(LOTD_CONFIGS["mixed"])), g["x"], g["params"])


def test_fast_path_hash_arithmetic_matches_oracle():
    """The fast kernels (csrc/lotd_pair.cuh:pair_geo) derive the four Hash corners of a lane from TWO products and two additions -- the product of the
    +1 corner is the product of the cell plus the prime, in uint32 arithmetic -- and reduce with a mask for power-of-two tables (host-side
    `fast_levels`: hmask = size - 1 for size > 1, else modulo).  Restated here in numpy uint32 and held against the oracle's index function
    (lotd_cuda.h hash: x * 1 ^ y * 2654435761 ^ z * 805459861, mod size) for every corner, power-of-two and odd table sizes, cells up to 2^11."""
    rs = np.random.RandomState(0)
    c = rs.randint(0, 2049, size=(20000, 3)).astype(np.uint32)
    P1, P2 = np.uint32(2654435761), np.uint32(805459861)
    with np.errstate(over="ignore"):
        y0 = c[:, 1] * P1
        y1 = y0 + P1
        z0 = c[:, 2] * P2
        z1 = z0 + P2
    for size in (1, 2, 4096, 2 ** 19, 2 ** 19 - 1, 300007, 2 ** 24):
        hmask = np.uint32(size - 1) if (size > 1 and size & (size - 1) == 0) else np.uint32(0)
        for side in (0, 1):
            hx = c[:, 0] + np.uint32(side)
            for q in range(4):
                dy, dz = q & 1, q >> 1
                h = hx ^ (y1 if dy else y0) ^ (z1 if dz else z0)
                mine = (h & hmask) if hmask else (h % np.uint32(size))
                pos = torch.from_numpy(np.stack([c[:, 0] + side, c[:, 1] + dy, c[:, 2] + dz], -1).astype(np.int64))
                want = O._idx_hash(pos, size).numpy().astype(np.uint32)
                assert np.array_equal(mine, want), (size, side, q)
