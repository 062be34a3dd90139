"""Drop-in proof on the GPU: the reference's UNMODIFIED python wrappers -- nr3d_lib/models/grid_encodings/lotd/lotd.py:48-458
(LoTDFunction / LoTDFunctionFwdDydx / LoTDFunctionBwdDydx / LoTD), nr3d_lib/graphics/pack_ops/pack_ops.py:97-392 and
nr3d_lib/graphics/raymarch/occgrid_raymarch.py:25-221 -- run forward / backward / double backward on top of nr3d_lib_b200's shims
(`install()` registers them as nr3d_lib.bindings._lotd / _pack_ops / _occ_grid) and are checked against the CPU oracles.

The wrapper files are loaded from /root/reference when present, else from the verbatim staging copy oracle/_ref/pyref/ (tests/util.py:
reference_wrappers).  Every LoTD case runs with the cell-sorted fast path on (the default a drop-in user gets: row-major y flowing through
code written for the reference's transposed views) and off (reference strides, generic kernels).
"""
import numpy as np
import pytest
import torch

from tests.util import LOTD_CONFIGS, lotd_inputs, march_inputs, meta_args, pack_inputs, reference_wrappers, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    mods = reference_wrappers()
    if mods is None:
        pytest.skip("reference python wrappers not available (neither /root/reference nor oracle/_ref/pyref)")
    return mods


def _launched(names_before):
    from nr3d_lib_b200 import _lib
    return _lib.launch_count() - names_before


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("name", ["ngp8", "hash_f4", "ngp_smooth", "mixed"])
def test_reference_lotd_forward_backward(name, pdtype, fast, dev, ref):
    """LoTD.forward -> .backward() through the reference's LoTDFunction (lotd.py:48-119): clamp, flatten, loss scale 128 for half tables."""
    from nr3d_lib_b200 import _lib
    from oracle import lotd_oracle as O
    L = ref["lotd"]
    cfg = LOTD_CONFIGS[name]
    enc = L.LoTD(cfg["D"], cfg["res"], cfg["feats"], cfg["types"], hashmap_size=cfg["T"], use_smooth_step=cfg["smooth"], dtype=pdtype)
    enc.meta.c_sort_points = fast
    assert enc.loss_scale == (128.0 if pdtype == torch.float16 else 1.0)
    om = O.OracleMeta(*meta_args(cfg))
    N = 3000
    inp = lotd_inputs(cfg, enc.n_params, N=N, seed=21)
    half = pdtype == torch.float16
    # a [6, 500, 3] batch of points: the wrapper flattens and restores the prefix
    x = inp["x"].to(dev).view(6, 500, 3).requires_grad_(True)
    p = inp["params"].to(dev).to(pdtype).requires_grad_(True)
    w = inp["dL_dy"].to(dev).view(6, 500, -1)
    n0 = _lib.launch_count()
    y = enc(x, p)
    assert y.shape == (6, 500, enc.out_features) and y.dtype == pdtype
    (y.float() * w).sum().backward()
    assert _lib.launch_count() > n0          # our kernels ran (no silent fallback)
    pp = inp["params"].half().float() if half else inp["params"]
    gy = inp["dL_dy"].half().float() if half else inp["dL_dy"]
    y_o = O.encode(om, inp["x"], pp)
    gx_o, gp_o = O.bwd(om, gy, inp["x"], pp)
    tol, tol_at = (4e-3, 3e-2) if half else (1e-5, 2e-5)
    assert rel_err(y.detach().float().cpu().view(N, -1), y_o) < tol
    assert rel_err(p.grad.float().cpu(), gp_o) < tol_at
    assert rel_err(x.grad.cpu().view(N, 3), gx_o) < (2e-2 if half else 1e-5)


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("pdtype", [torch.float32, torch.float16])
@pytest.mark.parametrize("name", ["ngp8", "mixed"])
def test_reference_lotd_nablas_double_backward(name, pdtype, fast, dev, ref):
    """forward_dydx -> backward_dydx (nablas) -> backward of an eikonal-style loss onto dL_dy and the parameters
    (LoTDFunctionFwdDydx / LoTDFunctionBwdDydx, lotd.py:121-268)."""
    from oracle import lotd_oracle as O
    L = ref["lotd"]
    cfg = LOTD_CONFIGS[name]
    enc = L.LoTD(cfg["D"], cfg["res"], cfg["feats"], cfg["types"], hashmap_size=cfg["T"], use_smooth_step=cfg["smooth"], dtype=pdtype)
    enc.meta.c_sort_points = fast
    om = O.OracleMeta(*meta_args(cfg))
    N = 2000
    inp = lotd_inputs(cfg, enc.n_params, N=N, seed=22)
    half = pdtype == torch.float16
    x = inp["x"].to(dev).requires_grad_(True)
    p = inp["params"].to(dev).to(pdtype).requires_grad_(True)
    w = inp["dL_dy"].to(dev)
    h, dy_dx = enc.forward_dydx(x, p)
    sdf = (h.float() * w).sum(-1)
    dL_dh = torch.autograd.grad(sdf.sum(), h, create_graph=True)[0]
    nablas = enc.backward_dydx(dL_dh, dy_dx, x, p)
    # half tables: the wrapper multiplies dL_dy by the loss scale 128 and the second-order table accumulates res * 128 * dL_dy * v in
    # half -- realistic eikonal weights are small, N(0,1) weights overflow half in either build
    vs = 1.0e-3 if half else 1.0
    v = inp["dL_ddLdx"].to(dev) * vs
    (nablas * v).sum().backward()
    pp = inp["params"].half().float() if half else inp["params"]
    gy = inp["dL_dy"].half().float() if half else inp["dL_dy"]
    y_o, dydx_o = O.fwd_dydx(om, inp["x"], pp)
    gx_o, _ = O.bwd(om, gy, inp["x"], pp)
    _, gp2_o, _ = O.bwd_bwd_input(om, inp["dL_ddLdx"] * vs, gy, inp["x"], pp)
    tol, tol_at = (4e-3, 3e-2) if half else (1e-5, 2e-5)
    assert rel_err(h.detach().float().cpu(), y_o) < tol
    assert rel_err(dy_dx.detach().reshape(N, -1, 3).cpu(), dydx_o) < tol
    assert rel_err(nablas.detach().cpu(), gx_o) < (2e-2 if half else 1e-5)
    assert rel_err(p.grad.float().cpu(), gp2_o) < tol_at


@pytest.mark.parametrize("fast", [True, False])
@pytest.mark.parametrize("mode", ["bidx", "input_batched"])
def test_reference_lotd_batched(mode, fast, dev, ref):
    """Several scenes through the reference's wrapper: per-point `bidx` (incl. -1) and `input_batched=True` ([B, n, 3] points)."""
    from oracle import lotd_oracle as O
    L = ref["lotd"]
    cfg = LOTD_CONFIGS["batched_hash"]
    B = cfg["B"]
    enc = L.LoTD(cfg["D"], cfg["res"], cfg["feats"], cfg["types"], hashmap_size=cfg["T"], dtype=torch.float)
    enc.meta.c_sort_points = fast
    om = O.OracleMeta(*meta_args(cfg))
    N = 3000
    inp = lotd_inputs(cfg, enc.n_params, N=N, seed=24, batch_mode="inds" if mode == "bidx" else "size")
    p = inp["params"].to(dev).requires_grad_(True)
    w = inp["dL_dy"].to(dev)
    if mode == "bidx":
        x = inp["x"].to(dev)
        y = enc(x, p, bidx=inp["batch_inds"].to(dev))
        okw = dict(batch_inds=inp["batch_inds"])
    else:
        x = inp["x"].to(dev).view(B, N // B, 3)
        y = enc(x, p, input_batched=True)
        w = w.view(B, N // B, -1)
        okw = dict(batch_data_size=N // B)
    (y * w).sum().backward()
    y_o = O.encode(om, inp["x"], inp["params"], **okw)
    _, gp_o = O.bwd(om, inp["dL_dy"], inp["x"], inp["params"], **okw)
    assert rel_err(y.detach().cpu().reshape(N, -1), y_o) < 1e-5
    assert rel_err(p.grad.cpu(), gp_o) < 2e-5


def test_reference_pack_ops(dev, ref):
    """packed_alpha_to_vw / packed_sum / packed_cumprod / packed_cumsum / packed_diff and the broadcast arithmetic through the reference's
    own autograd Functions (pack_ops.py:97-392): values against the sequential numpy oracle, gradients against torch formulas."""
    from oracle import pack_oracle as PO
    P = ref["pack_ops"]
    d = pack_inputs(P=300, max_len=120, C=3, seed=8)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi = t(d["pack_infos"])
    n = pi[:, 1]
    # values
    a = t(d["alphas"])
    w = P.packed_alpha_to_vw(a, pi, 1e-4, 0.0)
    w_o, _, _ = PO.alpha_to_vw_forward(d["alphas"], d["pack_infos"], 1e-4, 0.0)
    assert np.array_equal(w.cpu().numpy(), w_o)
    assert rel_err(P.packed_sum(t(d["feats1"]), pi).cpu(), PO.packed_sum(d["feats1"], d["pack_infos"])) < 3e-5
    assert rel_err(P.packed_cumsum(t(d["feats1"]), pi).cpu(), PO.packed_cumsum(d["feats1"], d["pack_infos"], False, False)) < 3e-5
    assert rel_err(P.packed_cumprod(t(d["prod1"]), pi).cpu(), PO.packed_cumprod(d["prod1"], d["pack_infos"], False, False)) < 3e-5
    assert np.array_equal(P.packed_diff(t(d["feats1"]), pi).cpu().numpy(), PO.packed_diff(d["feats1"], d["pack_infos"]))
    # gradients (float64 inputs go through the same kernels)
    f = t(d["featsC"]).double().requires_grad_(True)
    o = t(d["otherC"]).double().requires_grad_(True)
    for op, tf in ((P.packed_add, lambda x, y: x + y), (P.packed_sub, lambda x, y: x - y), (P.packed_mul, lambda x, y: x * y),
                   (P.packed_div, lambda x, y: x / y)):
        ww = torch.randn_like(f)
        g1 = torch.autograd.grad((op(f, o, pi) * ww).sum(), [f, o])
        g2 = torch.autograd.grad((tf(f, o.repeat_interleave(n, 0)) * ww).sum(), [f, o])
        assert rel_err(g1[0], g2[0]) < 1e-10 and rel_err(g1[1], g2[1]) < 1e-10
    a64 = t(d["alphas"][:2000]).double().clamp(1e-3, 0.9).requires_grad_(True)
    pi_s = pi[: int((pi[:, 0] + pi[:, 1] <= 2000).sum())]
    a64s = a64[: int(pi_s[-1, 0] + pi_s[-1, 1])].detach().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda u: P.packed_alpha_to_vw(u, pi_s, 1e-9, 0.0), (a64s,), eps=1e-6, atol=1e-6, rtol=1e-4)
    s = t(d["feats1"])[: a64s.shape[0]].double().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda u: P.packed_sum(u, pi_s), (s,), eps=1e-6, atol=1e-7)
    assert torch.autograd.gradcheck(lambda u: P.packed_cumsum(u, pi_s), (s,), eps=1e-6, atol=1e-7)
    pr = t(d["prod1"])[: a64s.shape[0]].double().requires_grad_(True)
    assert torch.autograd.gradcheck(lambda u: P.packed_cumprod(u, pi_s), (pr,), eps=1e-6, atol=1e-6, rtol=1e-4)


@pytest.mark.parametrize("fast", [True, False])
def test_reference_march_encode_composite(fast, dev, ref):
    """The reference's occgrid_raymarch (occgrid_raymarch.py:25-110) -> its LoTD -> its packed_alpha_to_vw / packed_sum, end to end with a
    backward pass to the tables: marcher output bit-exact against the C oracle, features / weights / gradients against the oracles."""
    from oracle import lotd_oracle as O, march_oracle as MO, pack_oracle as PO
    L, P, M = ref["lotd"], ref["pack_ops"], ref["occgrid_raymarch"]
    d = march_inputs(R=4096, res=32, seed=12)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ret = M.occgrid_raymarch(t(d["grid"]), t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), step_size=0.02, max_steps=256)
    o = MO.ray_marching(d["rays_o"], d["rays_d"], d["near"], d["far"], d["roi"], d["grid"], 0, 0.02, 1e10, 0.0, 256)
    hit = np.nonzero(o["packed_info"][:, 1])[0]
    assert ret.num_hit_rays == len(hit) and np.array_equal(ret.ridx_hit.cpu().numpy(), hit)
    assert np.array_equal(ret.pack_infos.cpu().numpy(), o["packed_info"][hit].astype(np.int64))
    assert np.array_equal(ret.depth_samples.cpu().numpy(), o["t_starts"]) and np.array_equal(ret.ridx.cpu().numpy(), o["ridx"].astype(np.int64))
    assert np.array_equal(ret.deltas.cpu().numpy(), (o["t_ends"] - o["t_starts"]))
    cfg = LOTD_CONFIGS["ngp8"]
    enc = L.LoTD(cfg["D"], cfg["res"], cfg["feats"], cfg["types"], hashmap_size=cfg["T"], dtype=torch.float)
    enc.meta.c_sort_points = fast
    rs = np.random.RandomState(5)
    p = torch.from_numpy((rs.randn(enc.n_params) * 0.1).astype(np.float32)).to(dev).requires_grad_(True)
    x01 = ret.samples * 0.5 + 0.5
    h = enc(x01, p)
    sigma = torch.nn.functional.softplus(h.sum(-1) * 4.0)
    alpha = 1.0 - torch.exp(-sigma * ret.deltas)
    w = P.packed_alpha_to_vw(alpha, ret.pack_infos, 1e-4, 0.0)
    depth = P.packed_sum(w * ret.depth_samples, ret.pack_infos)
    depth.sum().backward()
    om = O.OracleMeta(*meta_args(cfg))
    xc = x01.detach().cpu().clamp(1e-6, 1 - 1e-6)
    h_o = O.encode(om, xc, p.detach().cpu())
    assert rel_err(h.detach().cpu(), h_o) < 1e-5
    w_o, _, _ = PO.alpha_to_vw_forward(alpha.detach().cpu().numpy(), ret.pack_infos.cpu().numpy(), 1e-4, 0.0)
    assert np.array_equal(w.detach().cpu().numpy(), w_o)
    # gradient of the same composition in float64 torch on the CPU (dense loop over packs), through the oracle's encode
    p64 = p.detach().cpu().double().requires_grad_(True)
    h64 = O.encode(om, xc, p64)
    s64 = torch.nn.functional.softplus(h64.sum(-1) * 4.0)
    a64 = 1.0 - torch.exp(-s64 * ret.deltas.cpu().double())
    tot = 0.0
    pin = ret.pack_infos.cpu().numpy()
    ts = ret.depth_samples.cpu().double()
    for b, k in pin:
        aa = a64[b:b + k]
        T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1.0 - aa[:-1]]), 0)
        keep = (T >= 1e-4).double()          # early stop of the compositing kernel (transmittance below eps contributes nothing)
        tot = tot + (aa * T * keep * ts[b:b + k]).sum()
    tot.backward()
    assert rel_err(p.grad.cpu(), p64.grad) < 1e-4
