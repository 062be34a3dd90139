"""Fused LoTD encoder + density decoder (csrc/lotd_fused.cu, tcgen05) against (a) the unfused path of this
package followed by a torch MLP on bf16-rounded operands (tight) and (b) the plain fp32 composition the reference runs
(lotd_nerf.py:169-178) at bf16 tolerance."""
import numpy as np
import pytest
import torch

from tests.util import rel_err

pytestmark = pytest.mark.gpu


def _ngp16(T=2 ** 14):
    res = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
    return (3, res, [2] * 16, ["Dense" if r ** 3 <= T else "Hash" for r in res], T, False)


def _mlp(h, w1, b1, w2, b2, act, bf16):
    r = (lambda t: t.to(torch.bfloat16).float()) if bf16 else (lambda t: t)
    hid = torch.relu(r(h) @ r(w1).t() + (0 if b1 is None else b1))
    out = r(hid) @ r(w2).t() + (0 if b2 is None else b2)
    s = out[:, 0]
    s = {"identity": s, "exp": s.exp(), "softplus": torch.nn.functional.softplus(s), "relu": s.relu()}[act]
    return s, out


@pytest.mark.parametrize("N,n_out,act,bias", [(1, 16, "identity", True), (127, 1, "exp", False), (5000, 16, "softplus", True), (70001, 7, "relu", True)])
def test_fused_density_matches_unfused(N, n_out, act, bias, dev):
    from nr3d_lib_b200.bindings import _lotd
    from nr3d_lib_b200.fused import FusedDensityDecoder
    meta = _lotd.LoDMeta(*_ngp16())
    g = torch.Generator(device=dev).manual_seed(N)
    x = torch.rand(N, 3, device=dev, generator=g).clamp(1e-6, 1 - 1e-6)
    params = (torch.rand(meta.n_params, device=dev, generator=g) - 0.5)
    w1 = torch.randn(64, 32, device=dev, generator=g) * 0.3
    w2 = torch.randn(n_out, 64, device=dev, generator=g) * 0.2
    b1 = torch.randn(64, device=dev, generator=g) * 0.1 if bias else None
    b2 = torch.randn(n_out, device=dev, generator=g) * 0.1 if bias else None
    dec = FusedDensityDecoder(meta, w1, b1, w2, b2, activation=act)
    sigma, out = dec.query_density(x, params, return_output=True)
    torch.cuda.synchronize(dev)
    h, _ = _lotd.lod_fwd(meta, x, params, need_input_grad=False)       # unfused features (parity-tested elsewhere)
    s_bf, o_bf = _mlp(h.float(), w1, b1, w2, b2, act, bf16=True)
    s_32, o_32 = _mlp(h.float(), w1, b1, w2, b2, act, bf16=False)
    assert out.shape == (N, n_out)
    assert rel_err(out.cpu(), o_bf.cpu()) < 2e-3, "vs MLP on bf16-rounded operands"
    assert rel_err(sigma.cpu(), s_bf.cpu()) < 4e-3
    assert rel_err(out.cpu(), o_32.cpu()) < 2e-2, "vs the fp32 composition (bf16 operand rounding)"
    # max_level and the sigma-only entry
    s2, none = dec.query_density(x, params, max_level=3)
    h3, _ = _lotd.lod_fwd(meta, x, params, max_level=3, need_input_grad=False)
    assert none is None and rel_err(s2.cpu(), _mlp(h3.float(), w1, b1, w2, b2, act, True)[0].cpu()) < 4e-3


def test_fused_density_errors(dev):
    from nr3d_lib_b200.bindings import _lotd
    from nr3d_lib_b200.fused import FusedDensityDecoder
    meta = _lotd.LoDMeta(*_ngp16())
    w1, w2 = torch.zeros(64, 32, device=dev), torch.zeros(4, 64, device=dev)
    with pytest.raises(RuntimeError):
        FusedDensityDecoder(meta, torch.zeros(32, 32, device=dev), None, w2, None)
    with pytest.raises(RuntimeError):
        FusedDensityDecoder(_lotd.LoDMeta(3, [8, 16], [2, 2], ["Dense", "VM"], None), w1, None, w2, None)
    dec = FusedDensityDecoder(meta, w1, None, w2, None)
    with pytest.raises(RuntimeError):
        dec.query_density(torch.zeros(4, 3), torch.zeros(meta.n_params))          # CPU tensors: no fallback
    s, _ = dec.query_density(torch.zeros(0, 3, device=dev), torch.zeros(meta.n_params, device=dev))
    assert s.shape == (0,)


def _bf(t):
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("N,n_out,act,bias,use_out", [(1, 1, "exp", True, False), (127, 16, "identity", True, True), (40000, 4, "softplus", True, True),
                                                         (70001, 1, "relu", False, False), (300001, 1, "exp", True, False)])
def test_fused_density_backward(N, n_out, act, bias, use_out, dev):
    """Training step through the fused encoder + decoder (csrc/lotd_fused_bwd.cu, tcgen05): dL/dparams, dW1, db1, dW2, db2 against
    (a) a step-by-step emulation with the kernel's bf16 operand roundings (tight) and (b) fp32 autograd through the unfused operators
    (lotd_nerf.py:136-178 composition) at bf16 tolerance."""
    from nr3d_lib_b200.bindings import _lotd
    from nr3d_lib_b200.fused import fused_density
    from nr3d_lib_b200.lotd import LoTDFunction
    meta = _lotd.LoDMeta(*_ngp16())
    g = torch.Generator(device=dev).manual_seed(N + 7)
    x = torch.rand(N, 3, device=dev, generator=g).clamp(1e-6, 1 - 1e-6)
    p0 = (torch.rand(meta.n_params, device=dev, generator=g) - 0.5)
    w1_0 = torch.randn(64, 32, device=dev, generator=g) * 0.3
    w2_0 = torch.randn(n_out, 64, device=dev, generator=g) * 0.2
    b1_0 = torch.randn(64, device=dev, generator=g) * 0.1 if bias else None
    b2_0 = torch.randn(n_out, device=dev, generator=g) * 0.1 if bias else None
    if act == "exp":
        w2_0 = w2_0 * 0.2            # keep exp() tame
    c_s = torch.randn(N, device=dev, generator=g)
    c_o = torch.randn(N, n_out, device=dev, generator=g)

    def leaves():
        mk = lambda t: None if t is None else t.clone().requires_grad_(True)
        return mk(p0), mk(w1_0), mk(b1_0), mk(w2_0), mk(b2_0)

    # ---- fused
    p, w1, b1, w2, b2 = leaves()
    sigma, out = fused_density(x, p, w1, b1, w2, b2, meta, activation=act, return_output=use_out)
    loss = (sigma * c_s).sum() + ((out * c_o).sum() if use_out else 0.0)
    loss.backward()
    torch.cuda.synchronize(dev)
    got = dict(p=p.grad, w1=w1.grad, w2=w2.grad, b1=None if b1 is None else b1.grad, b2=None if b2 is None else b2.grad)

    # ---- (b) fp32 autograd through the unfused operators
    pr, w1r, b1r, w2r, b2r = leaves()
    h = LoTDFunction.apply(meta, x, pr, None, None, 0, 1.0, None).float()
    s_r, o_r = _mlp(h, w1r, b1r, w2r, b2r, act, bf16=False)
    ((s_r * c_s).sum() + ((o_r * c_o).sum() if use_out else 0.0)).backward()
    ref = dict(p=pr.grad, w1=w1r.grad, w2=w2r.grad, b1=None if b1r is None else b1r.grad, b2=None if b2r is None else b2r.grad)

    # ---- (a) emulation with the kernel's roundings
    with torch.no_grad():
        hf, _ = _lotd.lod_fwd(meta, x, p0, need_input_grad=False)
        hb = _bf(hf.float())
        pre = hb @ _bf(w1_0).t() + (0 if b1_0 is None else b1_0)
        hid = _bf(pre.relu())
        o = hid @ _bf(w2_0).t() + (0 if b2_0 is None else b2_0)
        s = {"identity": o[:, 0], "exp": o[:, 0].exp(), "softplus": torch.nn.functional.softplus(o[:, 0]), "relu": o[:, 0].relu()}[act]
        dact = {"identity": torch.ones_like(s), "exp": s, "softplus": 1 - torch.exp(-s), "relu": (s > 0).float()}[act]
        G = torch.zeros(N, n_out, device=dev)
        if use_out:
            G += c_o
        G[:, 0] += c_s * dact
        Gb = _bf(G)
        dh = _bf((Gb @ _bf(w2_0)) * (pre > 0))
        dF = dh @ _bf(w1_0)
        _, gp = _lotd.lod_bwd(meta, dF.contiguous(), x, p0, None, need_input_grad=False, need_param_grad=True)
        emu = dict(p=gp, w1=dh.t() @ hb, w2=Gb.t() @ hid, b1=dh.sum(0) if bias else None, b2=G.sum(0) if bias else None)

    for k in ("p", "w1", "w2", "b1", "b2"):
        if got[k] is None:
            assert ref[k] is None
            continue
        assert got[k].shape == ref[k].shape, k
        e_emu, e_ref = rel_err(got[k].cpu(), emu[k].cpu()), rel_err(got[k].cpu(), ref[k].cpu())
        assert e_emu < 3e-3, (k, "vs bf16-operand emulation", e_emu)
        # bf16 operands against fp32 autograd (sanity check; the tight check is the emulation above): entries that few points contribute to
        # carry the full operand rounding, and ReLU / activation decisions taken on bf16-rounded pre-activations flip for values near zero
        # (seen: max-norm 1e-1, cosine 0.998 for db1 at N = 70001) -- the direction of the gradient is what training needs
        cos = torch.nn.functional.cosine_similarity(got[k].flatten().double(), ref[k].flatten().double(), dim=0)
        # (max-norm: a table entry of a fine level is fed by ONE sample; when that sample's output activation (relu) or a hidden unit flips
        # between the bf16 and the fp32 forward, the entry differs by its whole value -- seen 0.5 at N = 70001 -- so only the direction is asserted)
        assert cos > 0.99, (k, "vs fp32 autograd", e_ref, float(cos))
    # the module wrapper trains: one SGD step lowers a simple loss
    from nr3d_lib_b200.fused import FusedDensityMLP
    torch.manual_seed(0)
    mlp = FusedDensityMLP(meta, n_out=1, activation="softplus", device=dev)
    pt = p0.clone().requires_grad_(True)
    target = torch.rand(N, device=dev, generator=g)
    l0 = ((mlp(x, pt) - target) ** 2).mean()
    l0.backward()
    with torch.no_grad():
        for t in [pt] + list(mlp.parameters()):
            t -= t.grad * (1.0e-2 / (t.grad.abs().max() + 1e-20))     # small step along -grad: no entry moves by more than 1e-2
    l1 = ((mlp(x, pt) - target) ** 2).mean()
    assert float(l1.detach()) < float(l0.detach())
