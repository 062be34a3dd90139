"""Shared test helpers: seeded input generators, config zoo, loaders for the oracle and the reference CUDA build."""
import importlib.util
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
REF_DIR = os.path.join(ROOT, "oracle", "_ref")

# --------------------------------------------------------------------------------------------------------------
# LoTD config zoo (name -> LoDMeta ctor args + batch info).  Small enough for committed golden fixtures.
# --------------------------------------------------------------------------------------------------------------
LOTD_CONFIGS = {
    # hash-only fast path: NGP-style Dense -> Hash ladder (gen_ngp_cfg geometry, lotd_cfg.py:48-57, reduced T)
    "ngp8": dict(D=3, res=[16, 22, 30, 42, 58, 80, 111, 154], feats=[2] * 8, types=["Dense"] * 3 + ["Hash"] * 5, T=2 ** 12, smooth=False, B=1),
    "hash_f4": dict(D=3, res=[12, 20, 33, 70], feats=[4, 4, 8, 4], types=["Dense", "Hash", "Hash", "Hash"], T=2 ** 10, smooth=False, B=1),
    "ngp_smooth": dict(D=3, res=[10, 17, 40], feats=[2, 2, 2], types=["Dense", "Hash", "Hash"], T=2 ** 9, smooth=True, B=1),
    # generic path: every level type, mixed widths (exercises pseudo levels)
    "mixed": dict(D=3, res=[8, 12, 16, 20, 24, 28], feats=[4, 4, 2, 2, 4, 2], types=["Dense", "VM", "CP", "CPfast", "NPlaneMul", "NPlaneSum"], T=None, smooth=False, B=1),
    "mixed_smooth": dict(D=3, res=[9, 11, 13, 10, 12], feats=[2, 2, 2, 2, 2], types=["VM", "CP", "CPfast", "NPlaneSum", "Dense"], T=None, smooth=True, B=1),
    "cuboid_vm": dict(D=3, res=[[8, 10, 12], [14, 9, 11]], feats=[2, 4], types=["VM", "Dense"], T=None, smooth=False, B=1),
    # batched scenes with per-point batch indices (-1 = skip)
    "batched": dict(D=3, res=[8, 16, 32], feats=[2, 2, 2], types=["Dense", "Hash", "VM"], T=2 ** 10, smooth=False, B=3),
    "batched_hash": dict(D=3, res=[8, 16, 32], feats=[2, 2, 2], types=["Dense", "Dense", "Hash"], T=2 ** 10, smooth=False, B=3),
    # other input dimensions
    "d2": dict(D=2, res=[10, 40, 90], feats=[2, 4, 2], types=["Dense", "Hash", "CP"], T=2 ** 9, smooth=False, B=1),
    "d2_hash": dict(D=2, res=[10, 40, 300], feats=[2, 2, 2], types=["Dense", "Dense", "Hash"], T=2 ** 9, smooth=False, B=1),
    "d4": dict(D=4, res=[5, 7, 9], feats=[2, 2, 2], types=["Dense", "Hash", "CP"], T=2 ** 10, smooth=False, B=1),
}


def meta_args(cfg):
    return (cfg["D"], cfg["res"], cfg["feats"], cfg["types"], cfg["T"], cfg["smooth"])


def lotd_inputs(cfg, n_params, N=192, seed=0, batch_mode="inds"):
    """Deterministic (numpy RandomState) inputs.  Returns dict of CPU torch tensors."""
    rs = np.random.RandomState(seed)
    D, B, E = cfg["D"], cfg["B"], sum(cfg["feats"])
    x = np.clip(rs.rand(N, D).astype(np.float32), 1e-6, 1 - 1e-6)
    # a few points close to cell borders / domain borders
    x[:4] = np.clip(np.round(x[:4] * 8) / 8 + 1e-4, 1e-6, 1 - 1e-6)
    params = (rs.randn(B * n_params) * 0.1).astype(np.float32)
    dL_dy = rs.randn(N, E).astype(np.float32)
    dL_ddLdx = rs.randn(N, D).astype(np.float32)
    out = dict(x=torch.from_numpy(x), params=torch.from_numpy(params), dL_dy=torch.from_numpy(dL_dy), dL_ddLdx=torch.from_numpy(dL_ddLdx),
               batch_inds=None, batch_data_size=0)
    if B > 1:
        if batch_mode == "inds":
            bi = rs.randint(0, B, size=N).astype(np.int64)
            bi[rs.rand(N) < 0.1] = -1
            out["batch_inds"] = torch.from_numpy(bi)
        else:
            out["batch_data_size"] = N // B
    return out


# --------------------------------------------------------------------------------------------------------------
# pack / march inputs
# --------------------------------------------------------------------------------------------------------------
def pack_inputs(P=37, max_len=70, C=3, seed=0, min_len=1):
    rs = np.random.RandomState(seed)
    n = rs.randint(min_len, max_len + 1, size=P).astype(np.int64)
    pack_infos = np.stack([np.cumsum(n) - n, n], 1)
    S = int(n.sum())
    d = dict(pack_infos=pack_infos, S=S, n=n,
             feats1=rs.randn(S).astype(np.float32), featsC=rs.randn(S, C).astype(np.float32),
             other1=(rs.randn(P) + 3.0).astype(np.float32), otherC=(rs.randn(P, C) + 3.0).astype(np.float32),
             prod1=(1.0 + 0.2 * rs.randn(S)).astype(np.float32),
             alphas=np.clip(rs.rand(S) ** 2 * 0.6, 0, 0.999).astype(np.float32), grad_w=rs.randn(S).astype(np.float32),
             near=(rs.rand(P) * 0.5 + 0.1).astype(np.float32), ids=np.repeat(rs.randint(0, 5, size=P), n).astype(np.int64))
    d["far"] = (d["near"] + rs.rand(P).astype(np.float32) * 2.0).astype(np.float32)
    d["alphas"][rs.rand(S) < 0.15] = 0.0   # exactly-zero alphas exercise the `<= thre` / `< thre` asymmetry
    return d


def march_inputs(R=512, res=32, seed=0, B=1, occupancy=0.5, shell=False):
    """Rays looking at the [-1,1]^3 box (the reference's smoke-test style, occgrid_raymarch.py:274-295, randomised)."""
    rs = np.random.RandomState(seed)
    if shell:
        g = np.stack(np.meshgrid(*[np.linspace(-1, 1, res)] * 3, indexing="ij"), -1)
        grid = (np.abs(np.linalg.norm(g, axis=-1) - 0.6) < 0.08)
        grid = np.broadcast_to(grid, (B, res, res, res)).copy()
    else:
        grid = rs.rand(B, res, res, res) > (1.0 - occupancy)
    o = rs.randn(R, 3)
    o = (4.0 * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    tgt = (rs.rand(R, 3) - 0.5).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    # ray / box [-1,1]^3 intersection (slab method) for near / far
    with np.errstate(divide="ignore"):
        t1 = (-1.0 - o) / d
        t2 = (1.0 - o) / d
    near = np.maximum(np.minimum(t1, t2).max(1), 0.0).astype(np.float32)
    far = np.maximum(t1, t2).min(1).astype(np.float32)
    miss = far <= near
    far[miss] = near[miss]
    roi = np.tile(np.array([-1, -1, -1, 1, 1, 1], dtype=np.float32), (B, 1))
    bi = rs.randint(0, B, size=R).astype(np.int32) if B > 1 else None
    return dict(rays_o=o, rays_d=d, near=near, far=far, roi=roi if B > 1 else roi[0], grid=grid if B > 1 else grid[0], batch_inds=bi)


# --------------------------------------------------------------------------------------------------------------
# loaders
# --------------------------------------------------------------------------------------------------------------
_ref_cache = {}


def load_ref(name, variant=""):
    """Load one of the reference's own CUDA extensions built by oracle/build_ref.py (None if absent).  `variant` selects a checker build
    of _lotd (oracle/build_ref.py --variant); pybind registers the reference's types per process, so only ONE _lotd build can be loaded
    into a process -- tests reach the other one through tests/ref_worker.py."""
    if name in _ref_cache:
        if _ref_cache[name] is not None and getattr(_ref_cache[name], "_nr3d_variant", "") != variant:
            raise RuntimeError(f"reference build {name} already loaded as variant {_ref_cache[name]._nr3d_variant!r}, cannot load {variant!r}")
        return _ref_cache[name]
    path = os.path.join(REF_DIR, name + (("__" + variant) if variant else "") + ".so")
    mod = None
    if os.path.exists(path):
        try:
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mod._nr3d_variant = variant
        except Exception as e:  # pragma: no cover
            print(f"[tests] could not load reference build {path}: {e}")
            mod = None
    _ref_cache[name] = mod
    return mod


def golden(name):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    if not os.path.exists(path):
        return None
    return dict(np.load(path, allow_pickle=False))


def rel_err(a, b):
    """max |a-b| / max(|b|_inf, tiny): error relative to the magnitude of the reference tensor."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    denom = max(b.abs().max().item(), 1e-30)
    return ((a - b).abs().max().item() / denom) if a.numel() else 0.0


def pack_next_inputs(P=29, max_len=60, nq=7, seed=0, min_len=2):
    """Sorted packs, CDFs and queries for the hierarchical-sampling ops (searchsorted / invert_cdf / merge / sort / matmul)."""
    rs = np.random.RandomState(seed)
    n = rs.randint(min_len, max_len + 1, size=P).astype(np.int64)
    nb = rs.randint(1, max_len // 2 + 1, size=P).astype(np.int64)
    pi = np.stack([np.cumsum(n) - n, n], 1)
    pib = np.stack([np.cumsum(nb) - nb, nb], 1)
    bins = np.concatenate([np.sort(rs.rand(k)).astype(np.float32) + i for i, k in enumerate(n)])          # strictly increasing per pack
    w = np.concatenate([rs.rand(k).astype(np.float32) + 0.01 for k in n])
    cdfs = np.concatenate([(np.cumsum(w[b:b + k]) / w[b:b + k].sum()).astype(np.float32) for b, k in pi])
    vals_q = (rs.rand(P, nq).astype(np.float32) * 1.2 - 0.1 + np.arange(P, dtype=np.float32)[:, None])   # some out of range on both sides
    u = rs.rand(P, nq).astype(np.float32)
    vals_b = np.concatenate([np.sort(rs.rand(k)).astype(np.float32) * 1.1 - 0.05 + i for i, k in enumerate(nb)])
    # duplicates: copy a few bins values into b so that ties (b == a[i]) are exercised
    for p in range(0, P, 3):
        vals_b[pib[p, 0]] = bins[pi[p, 0] + n[p] // 2]
        vals_b[pib[p, 0]:pib[p, 0] + nb[p]] = np.sort(vals_b[pib[p, 0]:pib[p, 0] + nb[p]])
    unsorted = rs.randn(int(n.sum())).astype(np.float32)
    feats = rs.randn(int(n.sum()), 3).astype(np.float32)
    mats = rs.randn(P, 4, 3).astype(np.float32)
    return dict(pack_infos=pi, pack_infos_b=pib, bins=bins, cdfs=cdfs, vals_q=vals_q, u=u, vals_b=vals_b, unsorted=unsorted, feats=feats, mats=mats)


def occ_inputs(res=(16, 12, 20), N=3000, B=3, seed=0):
    """Value grids + update samples for the occupancy-grid maintenance ops (single + batched, points and voxel indices)."""
    rs = np.random.RandomState(seed)
    r = np.array(res)
    grid = rs.rand(*res).astype(np.float32)
    bgrid = rs.rand(B, *res).astype(np.float32)
    pts = (rs.rand(N, 3) * 2.1 - 1.05).astype(np.float32)        # a few outside [-1,1]: exercises the clamp
    pts[:8] = np.array([-1.0, 1.0, 0.0])[rs.randint(0, 3, size=(8, 3))]   # exactly on cell / domain borders
    vals = (rs.rand(N) * 1.2 - 0.1).astype(np.float32)
    vals[rs.rand(N) < 0.2] *= -1.0                                 # negative values (raw sdf occupancy can be negative)
    gidx = np.stack([rs.randint(0, k, size=N) for k in res], 1).astype(np.int64)
    gidx[N // 2:] = gidx[: N - N // 2]                             # many duplicates -> the max matters
    bidx = rs.randint(0, B, size=N).astype(np.int64)
    n = N // B
    bpts = (rs.rand(B, n, 3) * 2 - 1).astype(np.float32)
    bgidx = np.stack([rs.randint(0, k, size=(B, n)) for k in res], -1).astype(np.int64)
    bvals = rs.rand(B, n).astype(np.float32)
    vox = np.stack(np.nonzero(rs.rand(*res) > 0.8), 1).astype(np.int64)
    return dict(grid=grid, bgrid=bgrid, pts=pts, vals=vals, gidx=gidx, bidx=bidx, bpts=bpts, bgidx=bgidx, bvals=bvals, vox=vox,
                ema=np.float32(0.9), thre=np.float32(0.55))


def seg_inputs(P=41, max_segs=6, seed=0, n_points=64):
    """Rays with sorted, disjoint [entry, exit] segments (some before near / after far, some rays without segments) for the
    in-packed-segments sampler, plus octree nuggets (point indices into an int16 xyz table) for the consecutive-segment marker."""
    rs = np.random.RandomState(seed)
    n = rs.randint(0, max_segs + 1, size=P).astype(np.int64)
    n[0] = max(n[0], 1)
    n[-1] = max(n[-1], 1)            # the reference reads seg_pack_infos[-1] to size-check `entry`
    spi = np.stack([np.cumsum(n) - n, n], 1)
    entry, exit_ = [], []
    for k in n:
        cuts = np.sort(rs.rand(2 * k).astype(np.float32) * 6.0 + 0.05)
        entry.append(cuts[0::2]); exit_.append(cuts[1::2])
    entry = np.concatenate(entry).astype(np.float32) if n.sum() else np.zeros(0, np.float32)
    exit_ = np.concatenate(exit_).astype(np.float32) if n.sum() else np.zeros(0, np.float32)
    near = (rs.rand(P) * 1.5).astype(np.float32)
    far = (near + rs.rand(P).astype(np.float32) * 5.0 + 0.2).astype(np.float32)
    # octree nuggets: walks on the integer lattice with occasional jumps
    points = rs.randint(0, 16, size=(n_points, 3)).astype(np.int16)
    for i in range(1, n_points):
        if rs.rand() < 0.7:
            step = np.zeros(3, np.int16); step[rs.randint(3)] = rs.choice([-1, 1])
            points[i] = points[i - 1] + step
    ln = rs.randint(1, 9, size=P).astype(np.int64)
    opi = np.stack([np.cumsum(ln) - ln, ln], 1)
    pidx = np.concatenate([np.sort(rs.randint(0, n_points, size=k)) for k in ln]).astype(np.int32)
    return dict(seg_pack_infos=spi, entry=entry, exit=exit_, near=near, far=far, points=points, oct_pack_infos=opi, pidx=pidx)


# --------------------------------------------------------------------------------------------------------------
# forest (multi-block) inputs: blocks of an octree level, SPC octree bytes, rays and their block segments
# --------------------------------------------------------------------------------------------------------------
def build_octree(block_ks, level):
    """kaolin-SPC style octree of the occupied integer blocks at `level` (what kaolin.ops.spc.unbatched_points_to_octree +
    Spc.exsum give the reference, lotd/tests/math_test_forest.py:36-57): BFS node bytes (bit `x<<2|y<<1|z` = child occupied),
    exsum = exclusive prefix sum of the child counts (int32 [n_nodes+1], exsum[0] = 0 ... the reference's `identify` walks
    `ord = exsum[ord] + inclusive_popc`, csrc/forest/forest.h:25-57), the level's blocks in Morton order and the number of
    nodes above `level` (ForestMeta.level_poffset)."""
    ks = sorted({tuple(int(v) for v in k) for k in block_ks})
    def morton(k, lv):
        m = 0
        for d in range(lv):
            m |= (((k[0] >> d) & 1) << 2 | ((k[1] >> d) & 1) << 1 | ((k[2] >> d) & 1)) << (3 * d)
        return m
    levels = [None] * (level + 1)
    levels[level] = sorted(ks, key=lambda k: morton(k, level))
    for lv in range(level - 1, -1, -1):
        parents = {(k[0] >> 1, k[1] >> 1, k[2] >> 1) for k in levels[lv + 1]}
        levels[lv] = sorted(parents, key=lambda k: morton(k, lv))
    octree = []
    for lv in range(level):
        children = set(levels[lv + 1])
        for k in levels[lv]:
            byte = 0
            for c in range(8):
                ck = (2 * k[0] + ((c >> 2) & 1), 2 * k[1] + ((c >> 1) & 1), 2 * k[2] + (c & 1))
                if ck in children:
                    byte |= 1 << c
            octree.append(byte)
    octree = np.array(octree, dtype=np.uint8)
    pop = np.array([bin(int(b)).count("1") for b in octree], dtype=np.int32)
    exsum = np.concatenate([[0], np.cumsum(pop)]).astype(np.int32)
    poffset = sum(len(levels[lv]) for lv in range(level))
    return octree, exsum, np.array(levels[level], dtype=np.int16), int(poffset)


def forest_inputs(R=400, res=16, seed=0, level=2, occupancy=0.5, n_blocks=6, block_size=0.5, origin=(-1.0, -1.0, -1.0)):
    """A few blocks of a 2^level grid, one occupancy grid per block, rays through the forest and -- per ray -- the blocks
    it crosses front to back with entry / exit depths (slab test per block; the reference gets these from kaolin's SPC ray trace)."""
    rs = np.random.RandomState(seed)
    n = 1 << level
    all_k = np.stack(np.meshgrid(*[np.arange(n)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pick = all_k[rs.permutation(len(all_k))[:n_blocks]]
    octree, exsum, block_ks, poffset = build_octree(pick, level)
    B = block_ks.shape[0]
    grid = rs.rand(B, res, res, res) > (1.0 - occupancy)
    origin = np.asarray(origin, dtype=np.float32)
    bs = np.full(3, block_size, dtype=np.float32)
    ext_lo, ext_hi = origin, origin + bs * n
    ctr, half = (ext_lo + ext_hi) / 2, (ext_hi - ext_lo) / 2
    o = rs.randn(R, 3)
    o = (ctr + 2.5 * half.max() * o / np.linalg.norm(o, axis=1, keepdims=True)).astype(np.float32)
    tgt = (ctr + (rs.rand(R, 3) - 0.5) * 2 * half * 0.9).astype(np.float32)
    d = tgt - o
    d = (d / np.linalg.norm(d, axis=1, keepdims=True)).astype(np.float32)
    lo = origin + block_ks.astype(np.float32) * bs          # [B,3]
    hi = lo + bs
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (lo[None] - o[:, None]) / d[:, None]            # [R,B,3]
        t2 = (hi[None] - o[:, None]) / d[:, None]
    ent = np.minimum(t1, t2).max(-1)
    ext = np.maximum(t1, t2).min(-1)
    hit = ext > np.maximum(ent, 0.0)
    seg_b, seg_in, seg_out, spi = [], [], [], []
    for r in range(R):
        idx = np.nonzero(hit[r])[0]
        idx = idx[np.argsort(ent[r, idx], kind="stable")]
        spi.append((len(seg_b), len(idx)))
        seg_b += idx.tolist()
        seg_in += np.maximum(ent[r, idx], 0.0).tolist()
        seg_out += ext[r, idx].tolist()
    spi = np.array(spi, dtype=np.int32)
    seg_in, seg_out = np.array(seg_in, dtype=np.float32), np.array(seg_out, dtype=np.float32)
    near = np.zeros(R, dtype=np.float32)
    far = np.full(R, 1e4, dtype=np.float32)
    has = spi[:, 1] > 0
    near[has] = seg_in[spi[has, 0]]
    far[has] = seg_out[spi[has, 0] + spi[has, 1] - 1]
    return dict(rays_o=o, rays_d=d, near=near, far=far, seg_block_inds=np.array(seg_b, dtype=np.int32), seg_entries=seg_in, seg_exits=seg_out,
                seg_pack_infos=spi, grid=grid, block_ks=block_ks, octree=octree, exsum=exsum, level=level, level_poffset=poffset,
                world_origin=origin, world_block_size=bs)


FOREST_MARCH_CASES = {
    "basic": dict(inp=dict(R=600, res=16, seed=31, level=2, n_blocks=24, occupancy=0.4), step=0.01, mx=1e10, gamma=0.0, ms=512),
    "gamma": dict(inp=dict(R=400, res=12, seed=32, level=2, n_blocks=40, occupancy=0.6), step=0.004, mx=0.05, gamma=0.01, ms=256),
    "maxsteps": dict(inp=dict(R=300, res=8, seed=33, level=1, n_blocks=7, occupancy=0.9, block_size=1.0), step=0.01, mx=1e10, gamma=0.0, ms=23),
    "sparse": dict(inp=dict(R=500, res=20, seed=34, level=3, n_blocks=60, occupancy=0.15, block_size=0.25), step=0.003, mx=1e10, gamma=0.0, ms=1024),
}


# forest LoTD: level zoo restricted to the types the reference's forest kernels implement (lotd_forest.h:263-311)
FOREST_LOTD_CONFIGS = {
    "mixed": dict(D=3, res=[4, 8, 6, 8, 12], feats=[2, 4, 2, 4, 2], types=["Dense", "VM", "NPlaneMul", "CP", "Hash"], T=2 ** 8, smooth=False,
                  forest=dict(seed=51, level=2, n_blocks=24)),
    "smooth": dict(D=3, res=[5, 16, 9], feats=[2, 2, 2], types=["Dense", "Hash", "VM"], T=2 ** 9, smooth=True, forest=dict(seed=52, level=3, n_blocks=70)),
    "hash_f4": dict(D=3, res=[6, 20], feats=[4, 8], types=["Dense", "Hash"], T=2 ** 9, smooth=False, forest=dict(seed=53, level=1, n_blocks=5)),
}


def forest_lotd_inputs(cfg, n_params, N=256, seed=0):
    """Block-local points for a forest LoTD: a good share hugs the block faces / edges / corners so that neighbour lookups
    (present and absent neighbours) are exercised; 10 % of the points carry block index -1 (skipped)."""
    rs = np.random.RandomState(seed)
    f = forest_inputs(R=4, res=2, **cfg["forest"])
    B, E = f["block_ks"].shape[0], sum(cfg["feats"])
    x = rs.rand(N, 3).astype(np.float32)
    edge = rs.rand(N, 3) < 0.3
    side = rs.rand(N, 3) < 0.5
    near = (rs.rand(N, 3) * 0.04).astype(np.float32)
    x = np.where(edge, np.where(side, near, 1.0 - near), x).astype(np.float32)
    x = np.clip(x, 1e-6, 1 - 1e-6)
    bi = rs.randint(0, B, size=N).astype(np.int64)
    bi[rs.rand(N) < 0.1] = -1
    return dict(x=torch.from_numpy(x), params=torch.from_numpy((rs.randn(B * n_params) * 0.1).astype(np.float32)),
                dL_dy=torch.from_numpy(rs.randn(N, E).astype(np.float32)), dL_ddLdx=torch.from_numpy(rs.randn(N, 3).astype(np.float32)),
                batch_inds=torch.from_numpy(bi), forest=f)


# --------------------------------------------------------------------------------------------------------------
# the reference's own python wrappers, UNMODIFIED, on top of this package's shims
# --------------------------------------------------------------------------------------------------------------
_refpy_cache = {}


def reference_wrappers():
    """Import the reference's unmodified `lotd.py`, `pack_ops.py` and `occgrid_raymarch.py` with `nr3d_lib.bindings.*` provided by
    nr3d_lib_b200 (install()).  The files come from /root/reference when it exists (this container) and otherwise from the verbatim staging
    copy `oracle/_ref/pyref/` that oracle/build_ref.py makes (git-ignored, shipped to the GPU box like the compiled reference .so files).
    Parent packages are namespace shells: their own __init__ files pull in the whole model zoo and third-party deps that are out of scope.
    Returns a dict(lotd=..., pack_ops=..., occgrid_raymarch=..., root=...) or None when neither location is available."""
    if "mods" in _refpy_cache:
        return _refpy_cache["mods"]
    import types
    root = None
    for cand in ("/root/reference/nr3d_lib", os.path.join(REF_DIR, "pyref", "nr3d_lib")):
        if os.path.isfile(os.path.join(cand, "models", "grid_encodings", "lotd", "lotd.py")):
            root = cand
            break
    if root is None:
        _refpy_cache["mods"] = None
        return None
    from nr3d_lib_b200.install import install

    def shell(name, path):
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
        return m
    shell("nr3d_lib", root)
    shell("nr3d_lib.models", root + "/models")
    shell("nr3d_lib.models.grid_encodings", root + "/models/grid_encodings")
    shell("nr3d_lib.models.grid_encodings.lotd", root + "/models/grid_encodings/lotd")
    shell("nr3d_lib.graphics", root + "/graphics")
    install()
    import importlib
    mods = dict(root=root,
                lotd=importlib.import_module("nr3d_lib.models.grid_encodings.lotd.lotd"),
                pack_ops=importlib.import_module("nr3d_lib.graphics.pack_ops.pack_ops"),
                occgrid_raymarch=importlib.import_module("nr3d_lib.graphics.raymarch.occgrid_raymarch"))
    for m in (mods["lotd"], mods["pack_ops"], mods["occgrid_raymarch"]):
        assert os.path.abspath(m.__file__).startswith(os.path.abspath(root)), m.__file__       # the reference's files, not our mirrors
    _refpy_cache["mods"] = mods
    return mods


def run_ref_worker(jobs, variant="G", timeout=1200):
    """Run LoTD jobs on another build of the reference's _lotd in a subprocess (tests/ref_worker.py).  jobs: list of dicts
    (name, dtype 'f32'|'f16', N, seed[, batch_mode]).  Returns {job_index: {output name: numpy array}} or None if that build is absent."""
    import json
    import subprocess
    import tempfile
    so = os.path.join(REF_DIR, "_lotd" + (("__" + variant) if variant else "") + ".so")
    if not os.path.exists(so):
        return None
    with tempfile.TemporaryDirectory() as td:
        out = os.path.join(td, "out.npz")
        r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "ref_worker.py"), variant, json.dumps(jobs), out],
                           capture_output=True, text=True, timeout=timeout)
        if r.returncode != 0 or not os.path.exists(out):
            raise RuntimeError("ref_worker failed:\n" + r.stdout[-2000:] + r.stderr[-4000:])
        z = np.load(out, allow_pickle=False)
        res = {}
        for k in z.files:
            j, key = k.split("/", 1)
            res.setdefault(int(j), {})[key] = z[k]
        return res


C4_ARGS = (3, [8, 16, 32, 64, 128, 256], [4, 4, 4, 4, 2, 2], ["Dense", "Dense", "VM", "VM", "CP", "CP"], None, False)   # BASELINE.json configs[3]


def c4_inputs(N, n_params, seed=11, B=8):
    """BASELINE.json configs[3]: B scenes, mixed Dense / VM / CP levels, per-point batch indices (1 % = -1).  CPU torch generator:
    identical in the test process and in tests/ref_worker.py."""
    g = torch.Generator().manual_seed(int(seed))
    x = torch.rand(N, 3, generator=g).clamp(1e-6, 1 - 1e-6)
    bi = torch.randint(0, B, (N,), generator=g)
    bi[torch.rand(N, generator=g) < 0.01] = -1
    params = torch.randn(B * n_params, generator=g) * 0.1
    E = sum(C4_ARGS[2])
    return dict(x=x, batch_inds=bi, params=params, dL_dy=torch.randn(N, E, generator=g), dL_ddLdx=torch.randn(N, 3, generator=g), batch_data_size=0)


EPS32 = 2.0 ** -24


def elementwise_excess(ours, f64, mag, ref=None, k=2.0, rel=1e-5, c_eps=8.0):
    """Element-wise arbiter check of an atomically accumulated fp32 table against its float64 value:

        |ours - f64|  <=  k * |ref - f64|  +  rel * |f64|  +  c_eps * eps32 * mag

    `mag` = sum of |terms| per entry: what fp32 summation ORDER may change; `ref` (optional) = the reference build's table, whose own
    distance to the float64 value is granted k-fold.  Returns (number of violating entries, worst ratio error / tolerance)."""
    o, t = np.asarray(ours, dtype=np.float64), np.asarray(f64, dtype=np.float64)
    tol = rel * np.abs(t) + c_eps * EPS32 * np.asarray(mag, dtype=np.float64) + 1e-300
    if ref is not None:
        tol = tol + k * np.abs(np.asarray(ref, dtype=np.float64) - t)
    ratio = np.abs(o - t) / tol
    return int((ratio > 1.0).sum()), float(ratio.max())
