"""Forest (multi-block) marcher -- SURVEY.md section 8f row n4, reference csrc/occ_grid/src/forest_marching.cu.

CPU: the C oracle against the golden vectors recorded from the reference's own CUDA build, plus structural properties.
GPU: the B200 kernel (through the C-ABI shim `nr3d_lib_b200.bindings._occ_grid.forest_ray_marching`) against the goldens, the
oracle on fresh inputs and -- when oracle/_ref is present -- the reference build live.  Everything bit-exact.
"""
import numpy as np
import pytest
import torch

from tests.util import FOREST_MARCH_CASES, forest_inputs, golden, load_ref


def _oracle(d, step, mx, gamma, ms):
    from oracle import march_oracle as MO
    return MO.forest_ray_marching(d["rays_o"], d["rays_d"], d["near"], d["far"], d["seg_block_inds"], d["seg_entries"], d["seg_exits"],
                                  d["seg_pack_infos"], d["block_ks"], d["world_origin"], d["world_block_size"], d["grid"], step, mx, gamma, ms)


def _equal(got, want, what):
    for k in ("packed_info", "ridx", "blidx", "gidx", "t_starts", "t_ends"):
        g, w = np.asarray(got[k]).reshape(-1), np.asarray(want[k]).reshape(-1)
        assert g.shape == w.shape and np.array_equal(g, w), f"{what}: {k} differs"


@pytest.mark.parametrize("name", list(FOREST_MARCH_CASES))
def test_forest_oracle_vs_golden(name):
    g = golden("forest_march_" + name)
    if g is None:
        pytest.skip("golden fixture not generated yet")
    cfg = g["cfg"]
    _equal(_oracle(g, float(cfg[0]), float(cfg[1]), float(cfg[2]), int(cfg[3])), g, "oracle vs golden:" + name)


def test_forest_inputs_are_what_the_fixture_stores():
    """The goldens embed their inputs; the generator must still produce the same ones (numpy RandomState is portable)."""
    for name, c in FOREST_MARCH_CASES.items():
        g = golden("forest_march_" + name)
        if g is None:
            pytest.skip("golden fixture not generated yet")
        d = forest_inputs(**c["inp"])
        for k in ("rays_o", "seg_entries", "seg_block_inds", "block_ks", "octree", "exsum"):
            assert np.array_equal(d[k], g[k]), (name, k)


def test_forest_oracle_properties():
    c = FOREST_MARCH_CASES["basic"]
    d = forest_inputs(**c["inp"])
    o = _oracle(d, c["step"], c["mx"], c["gamma"], c["ms"])
    pi = o["packed_info"]
    assert np.array_equal(pi[:, 0], np.cumsum(pi[:, 1]) - pi[:, 1]) and pi[:, 1].max() <= c["ms"]
    assert np.all(pi[d["seg_pack_infos"][:, 1] == 0, 1] == 0)                       # rays without segments get no samples
    assert np.all(o["t_ends"] > o["t_starts"]) and np.all(np.diff(o["ridx"]) >= 0)
    cells = d["grid"][0].size
    assert np.array_equal(o["gidx"] // cells, o["blidx"])                             # voxel index carries the block offset
    assert np.all(d["grid"].reshape(-1)[o["gidx"]])                                   # every sample sits in an occupied voxel
    # the block of a sample is one of its ray's segments, and samples of a ray visit blocks in segment order
    for r in np.nonzero(pi[:, 1])[0][:50]:
        b, n = d["seg_pack_infos"][r]
        segs = d["seg_block_inds"][b:b + n].tolist()
        seen = o["blidx"][pi[r, 0]:pi[r, 0] + pi[r, 1]].tolist()
        order = [segs.index(x) for x in seen]
        assert order == sorted(order)
    empty = _oracle(dict(d, grid=np.zeros_like(d["grid"])), c["step"], c["mx"], c["gamma"], c["ms"])
    assert empty["packed_info"][:, 1].sum() == 0


# ------------------------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------------------------
def _forest_meta(cls, d, t):
    m = cls()
    m.octree, m.exsum, m.block_ks = t(d["octree"]), t(d["exsum"]), t(d["block_ks"])
    m.world_origin, m.world_block_size = [float(v) for v in d["world_origin"]], [float(v) for v in d["world_block_size"]]
    m.n_trees, m.level, m.level_poffset = int(d["block_ks"].shape[0]), int(d["level"]), int(d["level_poffset"])
    m.resolution = [1 << int(d["level"])] * 3     # the reference's ForestMetaRef reads resolution[0..2] unconditionally (forest.h:77)
    return m


def _run(be, meta_cls, d, dev, step, mx, gamma, ms, gidx=True):
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    r = be.forest_ray_marching(_forest_meta(meta_cls, d, t), t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), t(d["seg_block_inds"]),
                               t(d["seg_entries"]), t(d["seg_exits"]), t(d["seg_pack_infos"]), t(d["grid"]), step, mx, gamma, ms, gidx)
    return r


def _as_dict(r):
    return {k: (None if v is None else v.cpu().numpy()) for k, v in zip(("packed_info", "t_starts", "t_ends", "ridx", "blidx", "gidx"), r)}


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(FOREST_MARCH_CASES))
def test_forest_march_vs_golden(name, dev):
    from nr3d_lib_b200.bindings import _occ_grid as og
    g = golden("forest_march_" + name)
    if g is None:
        pytest.skip("golden fixture not generated yet")
    cfg = g["cfg"]
    r = _run(og, og.ForestMeta, g, dev, float(cfg[0]), float(cfg[1]), float(cfg[2]), int(cfg[3]))
    assert r[0].dtype == torch.int32 and r[1].shape[1:] == (1,) and r[4].dtype == torch.int32
    _equal(_as_dict(r), g, "golden:" + name)


@pytest.mark.gpu
def test_forest_march_vs_oracle_and_reference_build(dev):
    from nr3d_lib_b200.bindings import _occ_grid as og
    d = forest_inputs(R=20000, res=32, seed=41, level=2, n_blocks=30, occupancy=0.3)
    r = _run(og, og.ForestMeta, d, dev, 0.004, 1e10, 0.0, 512)
    _equal(_as_dict(r), _oracle(d, 0.004, 1e10, 0.0, 512), "oracle")
    r_nog = _run(og, og.ForestMeta, d, dev, 0.004, 1e10, 0.0, 512, gidx=False)
    assert r_nog[5] is None and torch.equal(r_nog[4], r[4])
    ref, fm = load_ref("_occ_grid"), load_ref("_forest")
    if ref is not None and fm is not None:
        _equal(_as_dict(r), _as_dict(_run(ref, fm.ForestMeta, d, dev, 0.004, 1e10, 0.0, 512)), "reference build")


@pytest.mark.gpu
def test_forest_march_edges(dev):
    from nr3d_lib_b200.bindings import _occ_grid as og
    d = forest_inputs(R=64, res=8, seed=42, level=1, n_blocks=3)
    z = dict(d, rays_o=d["rays_o"][:0], rays_d=d["rays_d"][:0], near=d["near"][:0], far=d["far"][:0], seg_pack_infos=d["seg_pack_infos"][:0])
    r = _run(og, og.ForestMeta, z, dev, 0.01, 1e10, 0.0, 64)
    assert r[0].shape == (0, 2) and r[1].shape == (0, 1)
    e = _run(og, og.ForestMeta, dict(d, grid=np.zeros_like(d["grid"])), dev, 0.01, 1e10, 0.0, 64)
    assert int(e[0][:, 1].sum()) == 0 and e[3].numel() == 0
    with pytest.raises(RuntimeError):
        t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        og.forest_ray_marching(_forest_meta(og.ForestMeta, d, t), t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]),
                               t(d["seg_block_inds"]).long(), t(d["seg_entries"]), t(d["seg_exits"]), t(d["seg_pack_infos"]), t(d["grid"]),
                               0.01, 1e10, 0.0, 64, True)
    with pytest.raises(RuntimeError):
        c = lambda a: torch.from_numpy(np.ascontiguousarray(a))
        og.forest_ray_marching(_forest_meta(og.ForestMeta, d, c), c(d["rays_o"]), c(d["rays_d"]), c(d["near"]), c(d["far"]), c(d["seg_block_inds"]),
                               c(d["seg_entries"]), c(d["seg_exits"]), c(d["seg_pack_infos"]), c(d["grid"]), 0.01, 1e10, 0.0, 64, True)


@pytest.mark.gpu
def test_forest_wrapper(dev):
    """occgrid_raymarch_forest (mirror of the reference wrapper, occgrid_raymarch.py:223-272): packs per hit ray and per block run."""
    from nr3d_lib_b200.bindings import _occ_grid as og
    from nr3d_lib_b200.occgrid_raymarch import occgrid_raymarch_forest
    c = FOREST_MARCH_CASES["basic"]
    d = forest_inputs(**c["inp"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ret = occgrid_raymarch_forest(_forest_meta(og.ForestMeta, d, t), t(d["grid"]), t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]),
                                  t(d["seg_block_inds"]), t(d["seg_entries"]), t(d["seg_exits"]), t(d["seg_pack_infos"]), step_size=c["step"],
                                  max_steps=c["ms"])
    o = _oracle(d, c["step"], c["mx"], c["gamma"], c["ms"])
    hit = o["packed_info"][:, 1] > 0
    assert ret.num_hit_rays == int(hit.sum()) and np.array_equal(ret.ridx_hit.cpu().numpy(), np.nonzero(hit)[0])
    assert np.array_equal(ret.pack_infos.cpu().numpy(), o["packed_info"][hit].astype(np.int64))
    assert np.array_equal(ret.blidx.cpu().numpy(), o["blidx"].astype(np.int64))
    assert np.array_equal(ret.deltas.cpu().numpy(), o["t_ends"] - o["t_starts"])
    ridx = torch.from_numpy(o["ridx"].astype(np.int64)).to(dev)      # the reference composes the positions on the GPU (occgrid_raymarch.py:268)
    want = torch.addcmul(t(d["rays_o"]).index_select(0, ridx), t(d["rays_d"]).index_select(0, ridx), t(o["t_starts"]).unsqueeze(-1))
    assert torch.equal(ret.samples, want)
    bpi = ret.blidx_pack_infos.cpu().numpy()
    assert bpi[:, 1].sum() == o["blidx"].size and np.array_equal(bpi[:, 0], np.cumsum(bpi[:, 1]) - bpi[:, 1])
    for b, n in bpi[:200]:
        assert len(set(o["blidx"][b:b + n].tolist())) == 1
