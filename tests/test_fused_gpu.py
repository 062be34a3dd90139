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
