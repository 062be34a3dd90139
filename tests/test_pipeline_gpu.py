"""End-to-end parity of march -> encode -> composite (BASELINE.json configs[2] shape, reduced size): the B200 path
against a pipeline assembled from the three CPU oracles, forward values and dL/dparams."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests.util import march_inputs, rel_err

pytestmark = pytest.mark.gpu


def _ngp(levels=10, T=2 ** 14):
    res = (16 * 1.382 ** np.arange(levels)).astype(int).tolist()
    return dict(in_features=3, lod_res=res, lod_n_feats=[2] * levels, lod_types=["Dense" if r ** 3 <= T else "Hash" for r in res], hashmap_size=T)


@pytest.mark.parametrize("sort_points,fuse_head", [(False, False), (True, False), (True, True)])
def test_march_encode_composite_matches_oracles(sort_points, fuse_head, dev):
    from nr3d_lib_b200.lotd import LoTD
    from nr3d_lib_b200.pipeline import march_encode_composite
    from oracle import lotd_oracle as O, march_oracle as MO, pack_oracle as PO
    cfg = _ngp()
    enc = LoTD(dtype=torch.float, **cfg)
    enc.meta.c_sort_points = sort_points
    d = march_inputs(R=3000, res=32, seed=12, occupancy=0.25)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    rs = np.random.RandomState(1)
    p_host = torch.from_numpy((rs.randn(enc.n_params) * 0.05).astype(np.float32))
    params = p_host.to(dev).requires_grad_(True)
    out = march_encode_composite(enc, params, t(d["grid"]), t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), step_size=0.02, max_steps=256,
                                 fuse_head=fuse_head)
    assert out.march.num_hit_rays > 100
    loss = (out.depth ** 2).sum() + out.acc.sum()
    loss.backward()

    # ---- oracle pipeline (float64 everywhere after the bit-exact march)
    m = MO.ray_marching(d["rays_o"], d["rays_d"], d["near"], d["far"], d["roi"], d["grid"], 0, 0.02, 1e10, 0.0, 256)
    assert np.array_equal(out.march.pack_infos.cpu().numpy()[:, 1], m["packed_info"][m["packed_info"][:, 1] > 0, 1])
    t0, t1, ridx = torch.from_numpy(m["t_starts"]), torch.from_numpy(m["t_ends"]), torch.from_numpy(m["ridx"]).long()
    samples = torch.addcmul(torch.from_numpy(d["rays_o"])[ridx], torch.from_numpy(d["rays_d"])[ridx], t0.unsqueeze(-1))   # fp32 like the wrapper
    assert torch.equal(samples, out.march.samples.cpu())
    x01 = (samples * 0.5 + 0.5).clamp(1e-6, 1 - 1e-6)
    om = O.OracleMeta(3, cfg["lod_res"], cfg["lod_n_feats"], cfg["lod_types"], cfg["hashmap_size"])
    pd = p_host.double().requires_grad_(True)
    h = O.encode(om, x01, pd)
    sigma = F.softplus(h.sum(-1) * 20.0)
    alpha = 1.0 - torch.exp(-sigma * (t1 - t0).double())
    pi = m["packed_info"][m["packed_info"][:, 1] > 0].astype(np.int64)
    w = torch.zeros_like(alpha)
    for b, n in pi:   # differentiable front-to-back compositing (early stop never triggers at eps=1e-4 for this scene's opacities)
        a = alpha[b:b + n]
        T = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.float64), 1.0 - a[:-1]]), 0)
        w[b:b + n] = torch.where(T >= 1e-4, a * T, torch.zeros_like(a))
    w_np, _, _ = PO.alpha_to_vw_forward(alpha.detach().float().numpy(), pi, 1e-4, 0.0)
    assert rel_err(out.weights.detach().cpu(), w_np) < 1e-5 and rel_err(w.detach(), w_np) < 1e-5
    depth = torch.stack([(w[b:b + n] * t0[b:b + n].double()).sum() for b, n in pi])
    acc = torch.stack([w[b:b + n].sum() for b, n in pi])
    assert rel_err(out.depth.detach().cpu(), depth.detach()) < 1e-5 and rel_err(out.acc.detach().cpu(), acc.detach()) < 1e-5
    ((depth ** 2).sum() + acc.sum()).backward()
    assert rel_err(params.grad.cpu(), pd.grad) < 5e-5


@pytest.mark.gpu
def test_host_fed_steps_match_direct_calls(dev):
    """The overlapped host-fed driver (3 streams, double buffers) returns, for every step, exactly what the direct calls give."""
    from nr3d_lib_b200.bindings import _lotd
    from nr3d_lib_b200.pipeline import HostFedLoTDStep
    meta = _lotd.LoDMeta(3, [16, 22, 30, 42, 58, 80, 111, 154], [2] * 8, ["Dense"] * 3 + ["Hash"] * 5, 2 ** 14)
    meta.c_sort_points = True
    N, steps = 20000, 5
    g = torch.Generator().manual_seed(3)
    xs = [torch.rand(N, 3, generator=g).clamp(1e-6, 1 - 1e-6).pin_memory() for _ in range(steps)]
    params = (torch.rand(meta.n_params, generator=g) * 0.2 - 0.1).to(dev)
    outs = [torch.empty(meta.n_params).pin_memory() for _ in range(steps)]
    pipe = HostFedLoTDStep(meta, params, N, dev, grad_of_y=lambda y: y * 0.5)
    pipe.prefetch(xs[0])
    for k in range(steps):
        pipe.step(xs[k + 1] if k + 1 < steps else None, outs[k])
    pipe.drain()
    torch.cuda.synchronize(dev)
    for k in range(steps):
        xd = xs[k].to(dev)
        y, _ = _lotd.lod_fwd(meta, xd, params, need_input_grad=False)
        _, want = _lotd.lod_bwd(meta, y * 0.5, xd, params, None, need_input_grad=False, need_param_grad=True)
        assert rel_err(outs[k], want.cpu()) < 1e-5, k         # atomics: summation order differs between runs
        assert outs[k].abs().max() > 0


@pytest.mark.gpu
@pytest.mark.parametrize("C", [32, 6, 5])
def test_density_alpha_matches_torch_composition(C, dev):
    """One-pass density head + opacity (csrc/pipeline_ops.cu) == softplus(sum * gain) -> 1 - exp(-sigma * delta), values and gradient."""
    from nr3d_lib_b200.pipeline import density_alpha, density_proxy
    g = torch.Generator().manual_seed(5)
    S = 10007
    h = (torch.randn(S, C, generator=g) * 0.3).to(dev)
    h[:4] = 3.0                                               # gain * sum > 20: softplus' linear branch
    deltas = (torch.rand(S, generator=g) * 0.05).to(dev)
    w = torch.randn(S, generator=g).to(dev)
    for hh in (h, h.t().contiguous().t()):                    # row-major and feature-major (the generic kernels' y layout)
        a = hh.clone().requires_grad_(True)
        alpha, sigma = density_alpha(a, deltas, 20.0)
        (alpha * w).sum().backward()
        b = hh.double().clone().requires_grad_(True)
        sig_ref = density_proxy(b, 20.0)
        alpha_ref = 1.0 - torch.exp(-sig_ref * deltas.double())
        (alpha_ref * w.double()).sum().backward()
        assert rel_err(sigma, sig_ref.detach()) < 1e-5 and rel_err(alpha.detach(), alpha_ref.detach()) < 1e-5
        assert rel_err(a.grad, b.grad) < 1e-5
    with pytest.raises(RuntimeError):
        density_alpha(h.cpu(), deltas.cpu(), 20.0)


@pytest.mark.gpu
@pytest.mark.parametrize("pdtype", [torch.float32, torch.float16])
def test_fused_density_head_matches_composition(pdtype, dev):
    """encode + density head fused into the pair kernels (lotd_fast.cu, HEAD) == density_alpha(LoTD(x)): alpha, sigma and dL/dparams."""
    from nr3d_lib_b200 import _lib
    from nr3d_lib_b200.lotd import LoTD
    from nr3d_lib_b200.pipeline import density_alpha, encode_density_alpha
    from oracle import lotd_oracle as O
    cfg = _ngp()
    enc = LoTD(dtype=pdtype, **cfg)
    g = torch.Generator().manual_seed(8)
    S = 30011
    x = torch.rand(S, 3, generator=g)
    x[:3] = torch.tensor([[0.0, 1.0, 0.5], [1.0, 1.0, 1.0], [0.0, 0.0, 0.0]])          # box boundary: clamped like the unfused path
    p_host = (torch.randn(enc.n_params, generator=g) * 0.05).to(pdtype)
    deltas = (torch.rand(S, generator=g) * 0.05).to(dev)
    w = torch.randn(S, generator=g).to(dev)
    xd = x.to(dev)
    pa = p_host.to(dev).requires_grad_(True)
    n0 = _lib.launch_count()
    alpha, sigma = encode_density_alpha(enc, xd, pa, deltas, 2.0)
    (alpha * w).sum().backward()
    fused_launches = _lib.launch_count() - n0
    pb = p_host.to(dev).requires_grad_(True)
    alpha_u, sigma_u = density_alpha(enc(xd, pb).float(), deltas, 2.0)
    (alpha_u * w).sum().backward()
    tol = 1e-5 if pdtype == torch.float32 else 2e-3
    assert rel_err(sigma, sigma_u) < tol and rel_err(alpha.detach(), alpha_u.detach()) < tol
    assert rel_err(pa.grad.float(), pb.grad.float()) < (2e-5 if pdtype == torch.float32 else 3e-2)
    assert fused_launches <= 10, fused_launches         # sort (4) + head forward + fingerprint check (4, early exits) + head backward
    if pdtype == torch.float32:                          # and against the float64 oracle
        om = O.OracleMeta(3, cfg["lod_res"], cfg["lod_n_feats"], cfg["lod_types"], cfg["hashmap_size"])
        pd = p_host.double().requires_grad_(True)
        h = O.encode(om, x.clamp(1e-6, 1 - 1e-6), pd)
        sig = F.softplus(h.sum(-1) * 2.0)
        al = 1.0 - torch.exp(-sig * deltas.cpu().double())
        (al * w.cpu().double()).sum().backward()
        assert rel_err(sigma.cpu(), sig.detach()) < 1e-5 and rel_err(alpha.detach().cpu(), al.detach()) < 1e-5
        assert rel_err(pa.grad.cpu(), pd.grad) < 2e-5
    enc.meta.c_sort_points = False                       # not eligible: falls back to the two-kernel composition, same values
    pc = p_host.to(dev).requires_grad_(True)
    alpha_g, _ = encode_density_alpha(enc, xd, pc, deltas, 2.0)
    assert rel_err(alpha_g.detach(), alpha_u.detach()) < tol


@pytest.mark.gpu
def test_fused_density_head_records_across_interleaved_calls(dev):
    """The fused head keeps a token of the sorted records for its backward.  Two point sets of DIFFERENT sizes share the one record slot and the
    one sort workspace of the stream (re-armed, not re-allocated, when the point count changes): forward A, forward B, backward A (records
    overwritten -> re-sorted behind the fingerprint), backward B (re-sorted again), then A alone (token valid -> no fingerprint pass)."""
    from nr3d_lib_b200 import _lib
    from nr3d_lib_b200.bindings import _lotd
    from nr3d_lib_b200.lotd import LoTD
    from nr3d_lib_b200.pipeline import encode_density_alpha
    cfg = _ngp()
    enc = LoTD(dtype=torch.float32, **cfg)
    g = torch.Generator().manual_seed(5)
    p_host = torch.randn(enc.n_params, generator=g) * 0.05
    sets = []
    for S in (20011, 7001, 26003):       # smaller, then larger than the first: the workspace is reused, then grown
        sets.append((torch.rand(S, 3, generator=g).to(dev), (torch.rand(S, generator=g) * 0.05).to(dev), torch.randn(S, generator=g).to(dev)))

    def alone(k):
        _lotd.clear_sort_cache()
        p = p_host.to(dev).requires_grad_(True)
        x, d, w = sets[k]
        a, _ = encode_density_alpha(enc, x, p, d, 2.0)
        (a * w).sum().backward()
        return a.detach(), p.grad.clone()

    ref = [alone(k) for k in range(3)]
    _lotd.clear_sort_cache()
    p = p_host.to(dev).requires_grad_(True)
    outs = [encode_density_alpha(enc, x, p, d, 2.0)[0] for x, d, _ in sets]            # three forwards back to back: one slot, three sizes
    assert len(_lotd._sort_cache) == 1
    for k in (0, 2, 1):                                                                 # backwards in another order
        p.grad = None
        (outs[k] * sets[k][2]).sum().backward(retain_graph=True)
        assert torch.equal(outs[k].detach(), ref[k][0])
        assert rel_err(p.grad, ref[k][1]) < 2e-5, k
    # a forward directly followed by its backward: the token is still valid, the backward launches no sort kernels at all
    p.grad = None
    a, _ = encode_density_alpha(enc, sets[1][0], p, sets[1][1], 2.0)
    n0 = _lib.launch_count()
    (a * sets[1][2]).sum().backward()
    assert _lib.launch_count() - n0 == 1                                                # the head backward kernel
    assert rel_err(p.grad, ref[1][1]) < 2e-5
    _lotd.clear_sort_cache()


@pytest.mark.gpu
def test_packed_weighted_sums_bit_identical_to_composition(dev):
    """(acc, depth) in one pass each way == packed_sum(w), packed_sum(w * t) and their autograd, bit for bit."""
    from nr3d_lib_b200.pack_ops import packed_sum
    from nr3d_lib_b200.pipeline import packed_weighted_sums
    from tests.util import pack_inputs
    for P, max_len, min_len in ((5000, 300, 1), (700, 40, 0), (3, 9000, 5000)):
        d = pack_inputs(P=P, max_len=max_len, C=1, seed=P, min_len=max(min_len, 1))
        n = d["n"].copy()
        if min_len == 0:
            n[::5] = 0
        pi = torch.from_numpy(np.stack([np.cumsum(n) - n, n], 1)).to(dev)
        S = int(n.sum())
        g = torch.Generator().manual_seed(P)
        t = (torch.rand(S, generator=g) * 6).to(dev)
        w0 = torch.rand(S, generator=g).to(dev)
        ca, cd = torch.randn(P, generator=g).to(dev), torch.randn(P, generator=g).to(dev)
        wa = w0.clone().requires_grad_(True)
        acc, dep = packed_weighted_sums(wa, t, pi)
        ((acc * ca).sum() + (dep * cd).sum()).backward()
        wb = w0.clone().requires_grad_(True)
        acc_r, dep_r = packed_sum(wb, pi), packed_sum(wb * t, pi)
        ((acc_r * ca).sum() + (dep_r * cd).sum()).backward()
        assert torch.equal(acc, acc_r) and torch.equal(dep, dep_r)
        assert torch.equal(wa.grad, wb.grad)
    # the coordinate map of the point sort: sorting x in [-1, 1] with (0.5, 0.5, clamp) == sorting clamp(x * 0.5 + 0.5)
    from nr3d_lib_b200.bindings import _lotd
    x = (torch.rand(50000, 3, generator=g) * 2.2 - 1.1).to(dev)
    by_index = lambda r: r[r[:, 3].contiguous().view(torch.int32).argsort()]      # (the order inside a bin is not deterministic)
    xs_a, _ = _lotd._sorted_points(x, expect_new=True, coord_map=(0.5, 0.5, True))
    xs_a = by_index(xs_a.clone())
    xs_b, _ = _lotd._sorted_points((x * 0.5 + 0.5).clamp(1e-6, 1 - 1e-6), expect_new=True)
    assert torch.equal(xs_a, by_index(xs_b))
    assert torch.equal(xs_a[:, :3], (x * 0.5 + 0.5).clamp(1e-6, 1 - 1e-6))
    xs_c, _ = _lotd._sorted_points(x)                      # same tensor, other map: must NOT reuse the mapped records
    assert torch.equal(by_index(xs_c)[:, :3], x)


@pytest.mark.gpu
def test_march_samples_bit_identical_to_torch(dev):
    from nr3d_lib_b200.bindings import _occ_grid
    g = torch.Generator().manual_seed(6)
    R, S = 513, 40001
    o, d = torch.randn(R, 3, generator=g).to(dev), torch.randn(R, 3, generator=g).to(dev)
    ridx = torch.randint(0, R, (S,), generator=g).sort().values.to(dev)
    t0 = (torch.rand(S, generator=g) * 5).to(dev)
    t1 = t0 + 0.01
    want = torch.addcmul(o.index_select(0, ridx), d.index_select(0, ridx), t0.unsqueeze(-1))
    for r in (ridx, ridx.int()):
        s, dl = _occ_grid.march_samples(o, d, t0.unsqueeze(-1), t1.unsqueeze(-1), r)
        assert torch.equal(s, want) and torch.equal(dl, t1 - t0)
    s, dl = _occ_grid.march_samples(o, d, t0, None, ridx)
    assert dl is None and torch.equal(s, want)
    e, _ = _occ_grid.march_samples(o, d, t0[:0], None, ridx[:0])
    assert e.shape == (0, 3)
