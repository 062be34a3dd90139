"""Hierarchical-sampling pack ops: searchsorted, invert_cdf, sorted merge, per-pack sort, matmul.
CPU part pins the oracle against the golden vectors of the reference CUDA build; GPU part checks the B200 kernels against
golden, the reference build live, and the oracle.  Index outputs are bit-exact."""
import numpy as np
import pytest
import torch

from oracle import pack_oracle as PO
from tests.util import golden, load_ref, pack_next_inputs, rel_err


def _oracle_all(d):
    pi, pib = d["pack_infos"], d["pack_infos_b"]
    out = {"ss": PO.packed_searchsorted(d["bins"], d["vals_q"], pi),
           "ss_packed": PO.packed_searchsorted(d["bins"], d["vals_b"], pi, pib)}
    out["icdf_samples"], out["icdf_idx"] = PO.packed_invert_cdf(d["bins"], d["cdfs"], d["u"], pi)
    out["merge_a"], out["merge_b"], out["merge_pi"] = PO.try_merge_two_packs_sorted_aligned(d["bins"], pi, d["vals_b"], pib)
    out["sorted"], out["sort_idx"] = PO.packed_sort(d["unsorted"], pi)
    out["matmul"] = PO.packed_matmul(d["feats"], d["mats"], pi)
    return out


def test_oracle_vs_golden():
    g = golden("pack_next")
    if g is None:
        pytest.skip("golden fixture missing")
    o = _oracle_all(g)
    for k in ("ss", "ss_packed", "icdf_idx", "merge_a", "merge_b", "merge_pi"):
        assert np.array_equal(o[k], g[k]), k
    assert np.array_equal(o["merge_a"], g["merge_a_unsorted_flag"]) and np.array_equal(o["merge_b"], g["merge_b_unsorted_flag"])
    assert np.array_equal(o["icdf_samples"], g["icdf_samples"])
    assert np.array_equal(o["sorted"], g["sorted"])
    assert np.array_equal(g["unsorted"][g["sort_idx"]], g["sorted"]) and np.array_equal(g["unsorted"][o["sort_idx"]], o["sorted"])
    assert rel_err(o["matmul"], g["matmul"]) < 1e-6


def test_oracle_merge_properties():
    d = pack_next_inputs(seed=3)
    pa, pb, pm = PO.try_merge_two_packs_sorted_aligned(d["bins"], d["pack_infos"], d["vals_b"], d["pack_infos_b"])
    merged = np.empty(len(pa) + len(pb), dtype=np.float32)
    merged[pa], merged[pb] = d["bins"], d["vals_b"]
    assert np.array_equal(np.sort(np.concatenate([pa, pb])), np.arange(len(merged)))   # a permutation
    for b, n in pm:
        assert np.all(np.diff(merged[b:b + n]) >= 0)                                    # every merged pack is sorted


# ------------------------------------------------------------------------------------------------------------------
def _gpu_all(be, d, dev):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi, pib = t(d["pack_infos"]), t(d["pack_infos_b"])
    out = {"ss": be.packed_searchsorted(t(d["bins"]), t(d["vals_q"]), pi),
           "ss_packed": be.packed_searchsorted_packed_vals(t(d["bins"]), pi, t(d["vals_b"]), pib)}
    out["icdf_samples"], out["icdf_idx"] = be.packed_invert_cdf(t(d["bins"]), t(d["cdfs"]), t(d["u"]), pi)
    out["merge_a"], out["merge_b"], out["merge_pi"] = be.try_merge_two_packs_sorted_aligned(t(d["bins"]), pi, t(d["vals_b"]), pib, True)
    v = t(d["unsorted"]).clone()
    out["sort_idx"] = be.packed_sort_qsort(v, pi, True)
    out["sorted"] = v
    out["matmul"] = be.packed_matmul(t(d["feats"]), t(d["mats"]), pi)
    return {k: v.cpu().numpy() for k, v in out.items()}


def _check(got, want, unsorted):
    for k in ("ss", "ss_packed", "icdf_idx", "merge_a", "merge_b", "merge_pi"):
        assert np.array_equal(got[k], want[k]), k
    assert rel_err(got["icdf_samples"], want["icdf_samples"]) < 1e-6
    assert np.array_equal(got["sorted"], want["sorted"])
    assert np.array_equal(unsorted[got["sort_idx"]], got["sorted"])
    assert rel_err(got["matmul"], want["matmul"]) < 1e-5


@pytest.mark.gpu
def test_gpu_vs_golden(dev):
    from nr3d_lib_b200.bindings import _pack_ops
    g = golden("pack_next")
    if g is None:
        pytest.skip("golden fixture missing")
    _check(_gpu_all(_pack_ops, g, dev), g, g["unsorted"])


@pytest.mark.gpu
@pytest.mark.parametrize("P,max_len,nq,seed", [(200, 100, 16, 1), (40, 5000, 64, 2), (3000, 12, 3, 3)])
def test_gpu_vs_oracle_and_reference(P, max_len, nq, seed, dev):
    from nr3d_lib_b200.bindings import _pack_ops
    d = pack_next_inputs(P=P, max_len=max_len, nq=nq, seed=seed)
    got = _gpu_all(_pack_ops, d, dev)
    if P * max_len <= 40000:
        _check(got, _oracle_all(d), d["unsorted"])
    ref = load_ref("_pack_ops")
    if ref is not None:
        _check(got, _gpu_all(ref, d, dev), d["unsorted"])
    # sortedness / permutation properties at any size
    for b, n in d["pack_infos"]:
        assert np.all(np.diff(got["sorted"][b:b + n]) >= 0)
    assert np.array_equal(np.sort(got["sort_idx"]), np.arange(len(d["unsorted"])))


@pytest.mark.gpu
def test_gpu_host_mirror_and_edges(dev):
    from nr3d_lib_b200 import pack_ops as P
    d = pack_next_inputs(P=10, max_len=30, nq=5, seed=5)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi, pib = t(d["pack_infos"]), t(d["pack_infos_b"])
    val, pm = P.merge_two_packs_sorted_aligned(t(d["bins"]), pi, t(d["vals_b"]), pib, return_val=True)
    for b, n in pm.cpu().numpy():
        assert torch.all(val[b:b + n][1:] >= val[b:b + n][:-1])
    sv, idx = P.packed_sort(t(d["unsorted"]), pi)
    assert torch.equal(sv, t(d["unsorted"])[idx])
    mm = P.packed_matmul(t(d["feats"]), t(d["mats"]), pi)
    assert rel_err(mm.cpu(), PO.packed_matmul(d["feats"], d["mats"], d["pack_infos"])) < 1e-5
    # int64 keys, float64 keys
    ints = (d["unsorted"] * 1000).astype(np.int64)
    vi = t(ints).clone()
    P.packed_sort_inplace(vi, pi, return_idx=False)
    assert np.array_equal(vi.cpu().numpy(), PO.packed_sort(ints, d["pack_infos"])[0])
    s64, i64 = P.packed_invert_cdf(t(d["bins"].astype(np.float64)), t(d["cdfs"].astype(np.float64)), t(d["u"].astype(np.float64)), pi)
    assert np.array_equal(i64.cpu().numpy(), PO.packed_invert_cdf(d["bins"].astype(np.float64), d["cdfs"].astype(np.float64), d["u"].astype(np.float64), d["pack_infos"])[1])


# ------------------------------------------------------------------------------------------------------------------
# segment-restricted depth sampler, deprecated depth sampler, octree consecutive-segment marker
# ------------------------------------------------------------------------------------------------------------------
from tests.util import seg_inputs  # noqa: E402

SEG_CFG = (48, 0.05, 0.02, 0.3)


def _seg_oracle(d, cfg=SEG_CFG):
    o = {}
    o["seg_t"], o["seg_d"], o["seg_sidx"], o["seg_nidx"], o["seg_pi"] = PO.sample_step_in_packed_segments(
        d["near"], d["far"], d["entry"], d["exit"], d["seg_pack_infos"], *cfg)
    o["dep_t"], o["dep_d"], o["dep_nidx"], o["dep_pi"] = PO.sample_step_wrt_depth_clamped(d["near"], d["far"], *cfg)
    o["mark_start"], o["mark_end"] = PO.octree_mark_consecutive_segments(d["pidx"], d["oct_pack_infos"], d["points"])
    return o


def _seg_gpu(be, d, dev, cfg=SEG_CFG):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    o = {}
    o["seg_t"], o["seg_d"], o["seg_sidx"], o["seg_nidx"], o["seg_pi"] = be.interleave_sample_step_wrt_depth_in_packed_segments(
        t(d["near"]), t(d["far"]), t(d["entry"]), t(d["exit"]), t(d["seg_pack_infos"]), *cfg)
    o["dep_t"], o["dep_d"], o["dep_nidx"], o["dep_pi"] = be.interleave_sample_step_wrt_depth_clamp_deprecated(t(d["near"]), t(d["far"]), *cfg)
    o["mark_start"], o["mark_end"] = be.octree_mark_consecutive_segments(t(d["pidx"]), t(d["oct_pack_infos"]), t(d["points"]))
    return {k: v.cpu().numpy() for k, v in o.items()}


def _seg_check(got, want):
    for k in ("seg_t", "seg_d", "seg_sidx", "seg_nidx", "seg_pi", "dep_t", "dep_d", "dep_nidx", "dep_pi", "mark_start", "mark_end"):
        assert np.array_equal(got[k], want[k]), k            # bit-exact: no fused operation on this path


def test_seg_oracle_vs_golden():
    g = golden("pack_seg")
    if g is None:
        pytest.skip("golden fixture missing")
    _seg_check(_seg_oracle(g), g)
    assert g["seg_t"].shape[0] > 100 and g["mark_end"].sum() > 41


def test_seg_oracle_properties():
    d = seg_inputs(seed=7)
    o = _seg_oracle(d)
    near, far = d["near"][o["seg_nidx"]], d["far"][o["seg_nidx"]]
    assert np.all(o["seg_t"] <= far) and np.all(o["seg_t"] > near)
    assert np.all(o["seg_t"] <= d["exit"][o["seg_sidx"]]) and np.all(o["seg_t"] >= d["entry"][o["seg_sidx"]])
    assert np.all(o["seg_pi"][:, 1] <= SEG_CFG[0])
    fixed = PO.octree_mark_consecutive_segments(d["pidx"], d["oct_pack_infos"], d["points"], offset_fix=True)
    assert fixed[0].sum() == fixed[1].sum() >= d["oct_pack_infos"].shape[0]


@pytest.mark.gpu
def test_seg_gpu_vs_golden(dev):
    from nr3d_lib_b200.bindings import _pack_ops
    g = golden("pack_seg")
    if g is None:
        pytest.skip("golden fixture missing")
    _seg_check(_seg_gpu(_pack_ops, g, dev), g)


@pytest.mark.gpu
@pytest.mark.parametrize("P,max_segs,seed", [(100, 4, 1), (5000, 9, 2)])
def test_seg_gpu_vs_oracle_and_reference(P, max_segs, seed, dev):
    from nr3d_lib_b200.bindings import _pack_ops
    d = seg_inputs(P=P, max_segs=max_segs, seed=seed, n_points=500)
    got = _seg_gpu(_pack_ops, d, dev)
    if P <= 1000:
        _seg_check(got, _seg_oracle(d))
    ref = load_ref("_pack_ops")
    if ref is not None:
        _seg_check(got, _seg_gpu(ref, d, dev))
    _pack_ops.OCTREE_SEGMENTS_OFFSET_FIX = True
    try:
        ms, me = _pack_ops.octree_mark_consecutive_segments(*(torch.from_numpy(d[k]).to(dev) for k in ("pidx", "oct_pack_infos", "points")))
    finally:
        _pack_ops.OCTREE_SEGMENTS_OFFSET_FIX = False
    want = PO.octree_mark_consecutive_segments(d["pidx"], d["oct_pack_infos"], d["points"], offset_fix=True)
    assert np.array_equal(ms.cpu().numpy(), want[0]) and np.array_equal(me.cpu().numpy(), want[1])


@pytest.mark.gpu
def test_seg_host_mirror(dev):
    from nr3d_lib_b200 import pack_ops as P
    d = seg_inputs(seed=9)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ts, ds, ridx, rpi, sidx, spi = P.interleave_sample_step_wrt_depth_in_packed_segments(
        t(d["near"]), t(d["far"]), t(d["entry"]), t(d["exit"]), t(d["seg_pack_infos"]), max_steps=48, dt_gamma=0.05, min_step_size=0.02, max_step_size=0.3)
    o = _seg_oracle(d)
    assert np.array_equal(ts.cpu().numpy(), o["seg_t"]) and np.array_equal(rpi.cpu().numpy(), o["seg_pi"])
    assert int(spi[:, 1].sum()) == ts.shape[0]
    ts2 = P.interleave_sample_step_wrt_depth_in_packed_segments(0.1, 5.0, t(d["entry"]), t(d["exit"]), t(d["seg_pack_infos"]), perturb=True)[0]
    assert torch.isfinite(ts2).all()
