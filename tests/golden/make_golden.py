#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/*.npz from the REFERENCE'S OWN CUDA BUILD (oracle/_ref/*.so, produced by
oracle/build_ref.py from the unmodified sources under /root/reference/csrc).  Must run on a GPU box:

    gpurun -- 'python tests/golden/make_golden.py --out gpurun_out/golden'     # then copy the .npz into tests/golden/

Inputs come from tests/util.py (numpy RandomState, portable); every fixture stores inputs AND outputs so that the CPU
tests can check the oracle against them without a GPU, and the GPU tests can check the B200 kernels against them.
Nothing from this repository's kernels or oracle participates in producing the numbers.
"""
import argparse
import zlib
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests.util import FOREST_LOTD_CONFIGS, FOREST_MARCH_CASES, LOTD_CONFIGS, forest_lotd_inputs, forest_inputs, lotd_inputs, load_ref, march_inputs, meta_args, pack_inputs, pack_next_inputs, seg_inputs  # noqa: E402

dev = torch.device("cuda:0")


def npy(t):
    return None if t is None else t.detach().cpu().numpy()


def save(out_dir, name, **arrays):
    arrays = {k: v for k, v in arrays.items() if v is not None}
    np.savez_compressed(os.path.join(out_dir, name + ".npz"), **arrays)
    print(f"[golden] {name}: " + ", ".join(f"{k}{tuple(v.shape)}" for k, v in arrays.items()))


def make_lotd(out_dir):
    ref = load_ref("_lotd")
    for name, cfg in LOTD_CONFIGS.items():
        meta = ref.LoDMeta(*meta_args(cfg))
        for pdtype, tag in ((torch.float32, "f32"), (torch.float16, "f16")):
            if tag == "f16" and name not in ("ngp8", "mixed", "batched"):
                continue
            inp = lotd_inputs(cfg, meta.n_params, N=192, seed=zlib.crc32(name.encode()) % 1000)
            x = inp["x"].to(dev)
            params = inp["params"].to(dev).to(pdtype)
            dL_dy = inp["dL_dy"].to(dev).to(pdtype)
            ddx = inp["dL_ddLdx"].to(dev)
            bi = None if inp["batch_inds"] is None else inp["batch_inds"].to(dev)
            kw = dict(batch_inds=bi, batch_offsets=None, batch_data_size=None, max_level=None)
            y, dy_dx = ref.lod_fwd(meta, x, params, need_input_grad=True, **kw)
            y0, _ = ref.lod_fwd(meta, x, params, need_input_grad=False, **kw)
            assert torch.equal(y, y0)
            dL_dx, dL_dparam = ref.lod_bwd(meta, dL_dy, x, params, dy_dx, need_input_grad=True, need_param_grad=True, **kw)
            g_dLdy, g_param2, g_x2 = ref.lod_bwd_bwd_input(meta, ddx, dL_dy, x, params, dy_dx, need_dLdinput_ddLdoutput=True,
                                                           need_dLdinput_dparams=True, need_dLdinput_dinput=True, **kw)
            ymax, _ = ref.lod_fwd(meta, x, params, batch_inds=bi, max_level=1, need_input_grad=False)
            gi = None
            if meta.c_hash_only:
                gi = ref.lod_get_grid_index(meta, x, **kw)
            dyv = dy_dx.reshape(192, meta.n_encoded_dims, meta.n_dims_to_encode) if dy_dx.dim() == 2 else dy_dx
            save(out_dir, f"lotd_{name}_{tag}", x=npy(x), params=npy(params), dL_dy=npy(dL_dy), dL_ddLdx=npy(ddx),
                 batch_inds=npy(bi), y=npy(y), dy_dx=npy(dyv), dL_dx=npy(dL_dx), dL_dparam=npy(dL_dparam), dL_ddLdy=npy(g_dLdy),
                 dL_dparam2=npy(g_param2), dL_dx2=npy(g_x2), y_maxlevel1=npy(ymax), grid_index=npy(gi))


# The reference's GENERIC kernels (compile_split_*.cu) are miscompiled by nvcc 12.9's optimiser for sm_100: dy/dx and everything derived
# from it is wrong for the n-linear level types at D >= 3 (scripts/ref_variant_check.py -> profiles/r2_ref_build_variants.txt; the same
# sources built with -G pass finite differences of their own forward, `-Xptxas -O0/-O1` do not).  The four affected outputs of the
# generic-path fixtures therefore come from the -G build (`python oracle/build_ref.py --variant G` -> oracle/_ref/_lotd__G.so), which a second
# pass merges into the fixtures written by the stock build:   python tests/golden/make_golden.py --only make_lotd_checker
CHECKER_KEYS = ("dy_dx", "dL_dx", "dL_ddLdy", "dL_dparam2")


def make_lotd_checker(out_dir):
    ref = load_ref("_lotd", variant="G")
    assert ref is not None, "oracle/_ref/_lotd__G.so missing: python oracle/build_ref.py --variant G"
    for name, cfg in LOTD_CONFIGS.items():
        meta = ref.LoDMeta(*meta_args(cfg))
        if meta.c_hash_only or cfg["D"] == 2:
            continue                                   # hash-only kernels / D = 2 are unaffected: the stock build's vectors stand
        for pdtype, tag in ((torch.float32, "f32"), (torch.float16, "f16")):
            path = os.path.join(out_dir, f"lotd_{name}_{tag}.npz")
            if not os.path.exists(path):
                continue
            g = dict(np.load(path, allow_pickle=False))
            x, params, dL_dy, ddx = (torch.from_numpy(g[k]).to(dev) for k in ("x", "params", "dL_dy", "dL_ddLdx"))
            bi = torch.from_numpy(g["batch_inds"]).to(dev) if "batch_inds" in g else None
            kw = dict(batch_inds=bi, batch_offsets=None, batch_data_size=None, max_level=None)
            y, dy_dx = ref.lod_fwd(meta, x, params, need_input_grad=True, **kw)
            dL_dx, dL_dparam = ref.lod_bwd(meta, dL_dy, x, params, dy_dx, need_input_grad=True, need_param_grad=True, **kw)
            g_dLdy, g_param2, g_x2 = ref.lod_bwd_bwd_input(meta, ddx, dL_dy, x, params, dy_dx, need_dLdinput_ddLdoutput=True,
                                                           need_dLdinput_dparams=True, need_dLdinput_dinput=True, **kw)
            N = x.shape[0]
            new = dict(dy_dx=npy(dy_dx.reshape(N, meta.n_encoded_dims, meta.n_dims_to_encode)), dL_dx=npy(dL_dx), dL_ddLdy=npy(g_dLdy), dL_dparam2=npy(g_param2))
            # the outputs the stock build gets right must not depend on the optimisation level beyond rounding
            for k, v in (("y", y), ("dL_dparam", dL_dparam), ("dL_dx2", g_x2)):
                err = np.abs(npy(v).astype(np.float64) - g[k].astype(np.float64)).max() / max(np.abs(g[k].astype(np.float64)).max(), 1e-30)
                assert err < (2e-2 if tag == "f16" else 1e-5), (name, tag, k, err)
            g.update(new)
            g["checker_build"] = np.asarray("nvcc -G for compile_split_*.cu (oracle/build_ref.py --variant G): " + ", ".join(CHECKER_KEYS))
            np.savez_compressed(path, **g)
            print(f"[golden] lotd_{name}_{tag}: {', '.join(CHECKER_KEYS)} replaced by the -G build's")


def make_pack(out_dir):
    ref = load_ref("_pack_ops")
    d = pack_inputs()
    t = lambda a: torch.from_numpy(a).to(dev)
    pi = t(d["pack_infos"])
    out = dict(pack_infos=d["pack_infos"], feats1=d["feats1"], featsC=d["featsC"], other1=d["other1"], otherC=d["otherC"],
               prod1=d["prod1"], alphas=d["alphas"], grad_w=d["grad_w"], near=d["near"], far=d["far"], ids=d["ids"])
    f1, fC, pr = t(d["feats1"]), t(d["featsC"]), t(d["prod1"])
    out["sum1"], out["sumC"] = npy(ref.packed_sum(f1, pi)), npy(ref.packed_sum(fC, pi))
    for ex in (0, 1):
        for rv in (0, 1):
            out[f"cumsum1_e{ex}r{rv}"] = npy(ref.packed_cumsum(f1, pi, bool(ex), bool(rv)))
            out[f"cumsumC_e{ex}r{rv}"] = npy(ref.packed_cumsum(fC, pi, bool(ex), bool(rv)))
            out[f"cumprod1_e{ex}r{rv}"] = npy(ref.packed_cumprod(pr, pi, bool(ex), bool(rv)))
    ap, apC = t(d["other1"]), t(d["otherC"])
    out["diff1"], out["diffC"] = npy(ref.packed_diff(f1, pi, None, None)), npy(ref.packed_diff(fC, pi, None, None))
    out["diff1_append"], out["diffC_fill"] = npy(ref.packed_diff(f1, pi, ap, None)), npy(ref.packed_diff(fC, pi, None, apC))
    out["bdiff1"], out["bdiffC_prepend"] = npy(ref.packed_backward_diff(f1, pi, None, None)), npy(ref.packed_backward_diff(fC, pi, apC, None))
    out["bdiff1_fill"] = npy(ref.packed_backward_diff(f1, pi, None, ap))
    for nm in ("add", "sub", "mul", "div", "gt", "geq", "lt", "leq", "eq", "neq"):
        out[f"{nm}1"] = npy(getattr(ref, "packed_" + nm)(f1, ap, pi))
        out[f"{nm}C"] = npy(getattr(ref, "packed_" + nm)(fC, apC, pi))
    al, gw = t(d["alphas"]), t(d["grad_w"])
    for tag, eps, thre in (("a", 1e-4, 0.0), ("b", 0.3, 0.05)):
        w, _, _ = ref.packed_alpha_to_vw_forward(al, pi, eps, thre, False)
        _, cpi, sel = ref.packed_alpha_to_vw_forward(al, pi, eps, thre, True)
        ga = ref.packed_alpha_to_vw_backward(w, gw, al, pi, eps, thre)
        out[f"vw_w_{tag}"], out[f"vw_cpi_{tag}"], out[f"vw_sel_{tag}"], out[f"vw_ga_{tag}"] = npy(w), npy(cpi), npy(sel), npy(ga)
    ar, ar_idx = ref.interleave_arange(t(d["n"]), True)
    out["arange"], out["arange_idx"] = npy(ar), npy(ar_idx)
    ls, ls_idx = ref.interleave_linstep(t(d["near"]), t(d["n"]), t(d["other1"] * 0.01), True)
    out["linstep"], out["linstep_idx"] = npy(ls), npy(ls_idx)
    ts, ds, ni, spi = ref.interleave_sample_step_wrt_depth_clamped(t(d["near"]), t(d["far"]), 64, 0.02, 0.01, 0.2)
    out["ss_t"], out["ss_d"], out["ss_idx"], out["ss_pi"] = npy(ts), npy(ds), npy(ni), npy(spi)
    out["boundaries"] = npy(ref.mark_pack_boundaries_cuda(t(d["ids"])))
    save(out_dir, "pack_ops", **out)


def make_pack_next(out_dir):
    ref = load_ref("_pack_ops")
    d = pack_next_inputs()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    pi, pib = t(d["pack_infos"]), t(d["pack_infos_b"])
    out = dict(d)
    out["ss"] = npy(ref.packed_searchsorted(t(d["bins"]), t(d["vals_q"]), pi))
    out["ss_packed"] = npy(ref.packed_searchsorted_packed_vals(t(d["bins"]), pi, t(d["vals_b"]), pib))
    smp, bidx = ref.packed_invert_cdf(t(d["bins"]), t(d["cdfs"]), t(d["u"]), pi)
    out["icdf_samples"], out["icdf_idx"] = npy(smp), npy(bidx)
    pa, pb, pm = ref.try_merge_two_packs_sorted_aligned(t(d["bins"]), pi, t(d["vals_b"]), pib, True)
    out["merge_a"], out["merge_b"], out["merge_pi"] = npy(pa), npy(pb), npy(pm)
    pa2, pb2, _ = ref.try_merge_two_packs_sorted_aligned(t(d["bins"]), pi, t(d["vals_b"]), pib, False)
    out["merge_a_unsorted_flag"], out["merge_b_unsorted_flag"] = npy(pa2), npy(pb2)
    v = t(d["unsorted"]).clone()
    idx = ref.packed_sort_qsort(v, pi, True)
    out["sorted"], out["sort_idx"] = npy(v), npy(idx)
    out["matmul"] = npy(ref.packed_matmul(t(d["feats"]), t(d["mats"]), pi))
    save(out_dir, "pack_next", **out)


def make_pack_seg(out_dir):
    ref = load_ref("_pack_ops")
    d = seg_inputs()
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    out = dict(d)
    cfg = (48, 0.05, 0.02, 0.3)
    ts, ds, sidx, nidx, pi = ref.interleave_sample_step_wrt_depth_in_packed_segments(t(d["near"]), t(d["far"]), t(d["entry"]), t(d["exit"]),
                                                                                      t(d["seg_pack_infos"]), *cfg)
    out.update(cfg=np.array(cfg, dtype=np.float64), seg_t=npy(ts), seg_d=npy(ds), seg_sidx=npy(sidx), seg_nidx=npy(nidx), seg_pi=npy(pi))
    ts, ds, nidx, pi = ref.interleave_sample_step_wrt_depth_clamp_deprecated(t(d["near"]), t(d["far"]), *cfg)
    out.update(dep_t=npy(ts), dep_d=npy(ds), dep_nidx=npy(nidx), dep_pi=npy(pi))
    ms, me = ref.octree_mark_consecutive_segments(t(d["pidx"]), t(d["oct_pack_infos"]), t(d["points"]))
    out.update(mark_start=npy(ms), mark_end=npy(me))
    save(out_dir, "pack_seg", **out)


def make_march(out_dir):
    ref = load_ref("_occ_grid")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    cases = {
        "aabb": dict(inp=dict(R=512, res=32, seed=1), ct=ref.ContractionType.AABB, step=0.01, mx=1e10, gamma=0.0, ms=512),
        "aabb_gamma": dict(inp=dict(R=384, res=24, seed=2, occupancy=0.3), ct=ref.ContractionType.AABB, step=0.005, mx=0.05, gamma=0.01, ms=256),
        "aabb_shell": dict(inp=dict(R=512, res=48, seed=3, shell=True), ct=ref.ContractionType.AABB, step=0.004, mx=1e10, gamma=0.0, ms=1024),
        "aabb_maxsteps": dict(inp=dict(R=256, res=16, seed=4, occupancy=0.9), ct=ref.ContractionType.AABB, step=0.01, mx=1e10, gamma=0.0, ms=17),
        "sphere": dict(inp=dict(R=256, res=32, seed=5), ct=ref.ContractionType.UN_BOUNDED_SPHERE, step=0.02, mx=1e10, gamma=0.0, ms=128),
        "tanh": dict(inp=dict(R=256, res=32, seed=6), ct=ref.ContractionType.UN_BOUNDED_TANH, step=0.02, mx=1e10, gamma=0.0, ms=128),
    }
    for name, c in cases.items():
        d = march_inputs(**c["inp"])
        res = ref.ray_marching(t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), t(d["roi"]), t(d["grid"]), c["ct"],
                               c["step"], c["mx"], c["gamma"], c["ms"], True)
        pi, t0, t1, ridx, gidx = res
        save(out_dir, "march_" + name, rays_o=d["rays_o"], rays_d=d["rays_d"], near=d["near"], far=d["far"], roi=d["roi"], grid=d["grid"],
             cfg=np.array([int(c["ct"]), c["step"], c["mx"], c["gamma"], c["ms"]], dtype=np.float64),
             packed_info=npy(pi), t_starts=npy(t0), t_ends=npy(t1), ridx=npy(ridx), gidx=npy(gidx))
    # batched (no negative batch indices: the reference leaves their counts uninitialised)
    for name, use_inds in (("batched_inds", True), ("batched_size", False)):
        d = march_inputs(R=384, res=24, seed=7, B=3)
        bi = t(d["batch_inds"]) if use_inds else None
        bds = None if use_inds else 128
        res = ref.batched_ray_marching(t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), bi, bds, t(d["roi"]), t(d["grid"]),
                                       ref.ContractionType.AABB, 0.01, 1e10, 0.0, 256, True)
        pi, t0, t1, ridx, bidx, gidx = res
        save(out_dir, "march_" + name, rays_o=d["rays_o"], rays_d=d["rays_d"], near=d["near"], far=d["far"], roi=d["roi"], grid=d["grid"],
             batch_inds=d["batch_inds"] if use_inds else None, cfg=np.array([0, 0.01, 1e10, 0.0, 256, bds or 0], dtype=np.float64),
             packed_info=npy(pi), t_starts=npy(t0), t_ends=npy(t1), ridx=npy(ridx), bidx=npy(bidx), gidx=npy(gidx))


def make_forest_march(out_dir):
    """forest_ray_marching of the reference build (oracle/_ref/_occ_grid.so); ForestMeta comes from oracle/_ref/_forest.so, our
    20-line pybind registration of the reference's struct (oracle/ref_forest_meta.cpp)."""
    fm = load_ref("_forest")
    ref = load_ref("_occ_grid")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for name, c in FOREST_MARCH_CASES.items():
        d = forest_inputs(**c["inp"])
        meta = fm.ForestMeta()
        meta.octree, meta.exsum, meta.block_ks = t(d["octree"]), t(d["exsum"]), t(d["block_ks"])
        meta.world_origin, meta.world_block_size = [float(v) for v in d["world_origin"]], [float(v) for v in d["world_block_size"]]
        meta.n_trees, meta.level, meta.level_poffset = int(d["block_ks"].shape[0]), int(d["level"]), int(d["level_poffset"])
        meta.resolution = [1 << int(d["level"])] * 3     # ForestMetaRef reads resolution[0..2] unconditionally (forest.h:77)
        res = ref.forest_ray_marching(meta, t(d["rays_o"]), t(d["rays_d"]), t(d["near"]), t(d["far"]), t(d["seg_block_inds"]), t(d["seg_entries"]),
                                      t(d["seg_exits"]), t(d["seg_pack_infos"]), t(d["grid"]), c["step"], c["mx"], c["gamma"], c["ms"], True)
        pi, t0, t1, ridx, blidx, gidx = res
        save(out_dir, "forest_march_" + name, cfg=np.array([c["step"], c["mx"], c["gamma"], c["ms"]], dtype=np.float64),
             packed_info=npy(pi), t_starts=npy(t0), t_ends=npy(t1), ridx=npy(ridx), blidx=npy(blidx), gidx=npy(gidx),
             **{k: np.asarray(v) for k, v in d.items()})


def _ref_forest_meta(fm, f, t):
    meta = fm.ForestMeta()
    meta.octree, meta.exsum, meta.block_ks = t(f["octree"]), t(f["exsum"]), t(f["block_ks"])
    meta.world_origin, meta.world_block_size = [float(v) for v in f["world_origin"]], [float(v) for v in f["world_block_size"]]
    meta.n_trees, meta.level, meta.level_poffset = int(f["block_ks"].shape[0]), int(f["level"]), int(f["level_poffset"])
    meta.resolution = [1 << int(f["level"])] * 3     # ForestMetaRef reads resolution[0..2] unconditionally (forest.h:77)
    return meta


def make_forest_lotd(out_dir):
    """lod_bwd / lod_bwd_bwd_input with metas=(LoDMeta, ForestMeta) of the reference build.  The reference's forest FORWARD cannot
    run at this commit (lod_fwd_common never allocates `output` / `dy_dx` on the forest branch, lotd_torch_api.cu:300-362), so the
    fixtures hold what does run: dL/dparam, d(dL/dx)/dparam, d(dL/dx)/dx.  No dy_dx is passed (it would have to come from our side)."""
    fm, ref = load_ref("_forest"), load_ref("_lotd")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    for name, cfg in FOREST_LOTD_CONFIGS.items():
        meta = ref.LoDMeta(*meta_args(cfg))
        for pdtype, tag in ((torch.float32, "f32"), (torch.float16, "f16")):
            if tag == "f16" and name != "mixed":
                continue
            inp = forest_lotd_inputs(cfg, meta.n_params, N=256, seed=zlib.crc32(name.encode()) % 1000)
            f = inp["forest"]
            metas = (meta, _ref_forest_meta(fm, f, t))
            x, params, dL_dy = inp["x"].to(dev), inp["params"].to(dev).to(pdtype), inp["dL_dy"].to(dev).to(pdtype)
            ddx, bi = inp["dL_ddLdx"].to(dev), inp["batch_inds"].to(dev)
            _, dL_dparam = ref.lod_bwd(metas, dL_dy, x, params, None, bi, None, None, None, False, True)
            _, g_param2, g_x2 = ref.lod_bwd_bwd_input(metas, ddx, dL_dy, x, params, None, bi, None, None, None, False, True, True)
            _, dL_dparam_ml = ref.lod_bwd(metas, dL_dy, x, params, None, bi, None, None, 1, False, True)
            save(out_dir, f"forest_lotd_{name}_{tag}", x=npy(x), params=npy(params), dL_dy=npy(dL_dy), dL_ddLdx=npy(ddx), batch_inds=npy(bi),
                 dL_dparam=npy(dL_dparam), dL_dparam2=npy(g_param2), dL_dx2=npy(g_x2), dL_dparam_maxlevel1=npy(dL_dparam_ml),
                 octree=f["octree"], exsum=f["exsum"], block_ks=f["block_ks"], level=np.asarray(f["level"]), level_poffset=np.asarray(f["level_poffset"]))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(os.path.abspath(__file__))))
    ap.add_argument("--only", default="", help="comma separated subset of make_lotd,make_pack,make_pack_next,make_march")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    assert torch.cuda.is_available(), "golden vectors are produced by the reference CUDA build: a GPU is required"
    only = set(args.only.split(',')) if args.only else None
    for fn in (make_lotd, make_lotd_checker, make_pack, make_pack_next, make_pack_seg, make_march, make_forest_march, make_forest_lotd):
        if only and fn.__name__ not in only:
            continue
        if fn is make_lotd_checker and not only:
            continue        # needs its own process: only one build of the reference's _lotd can be loaded at a time
        try:
            fn(args.out)
        except Exception as e:  # keep going so one failing family does not lose the others
            import traceback
            traceback.print_exc()
            print(f"[golden] {fn.__name__} FAILED: {e}")
