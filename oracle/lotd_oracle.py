"""CPU oracle for the LoTD encoder -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` legs may import
this module; nothing under ``nr3d_lib_b200/`` does.

A float64, differentiable (torch autograd) restatement of the reference's LoTD maths:

  meta            csrc/lotd/src/lotd_torch_api.cu:29-230              -> ``OracleMeta``
  pos_fract       csrc/lotd/include/lotd/lotd_cuda.h:959-1077         -> ``_pos``
  index functions csrc/lotd/include/lotd/lotd_cuda.h:92-296           -> ``_idx_*``
  corner values   csrc/lotd/include/lotd/lotd_cuda.h:298-492          -> ``_level_value``
  n-linear        csrc/lotd/include/lotd/linear_interpolate.cuh:92-150
  NPlaneSum       csrc/lotd/include/lotd/lotd_encoding.h:268-351, CPfast :353-410
  2nd order       csrc/lotd/include/lotd/lotd_encoding.h:713-762,1043-1298 (through autograd of the same function;
                  the reference's coverage holes -- no d(dL/dx)/dx for CP/CPfast/NPlane* -- are reproduced by masking)

Numerics: the cell index and the fractional offset are computed in float32 with a fused multiply-add
(``v = fma(x, res-2, 0.5)``, as nvcc contracts the reference expression) so that the oracle sits on exactly the same
piece of the piecewise-polynomial as the CUDA kernels; everything downstream is float64, i.e. the oracle is the
"infinitely precise" value both fp32 implementations are compared against.

Pinning status: ``OracleMeta`` is checked against the reference's own compiled ``LoDMeta`` (oracle/_ref/_lotd.so, host
code, runs without a GPU); Dense levels are checked against the reference's pure-PyTorch ``F.grid_sample`` path
(nr3d_lib/models/grid_encodings/lotd/lotd_helpers.py:327-346); every type and derivative is checked against golden
vectors produced by the reference's own CUDA build on a B200 (tests/golden/, script tests/golden/make_golden.py).
"""
from typing import List, Optional, Sequence

import numpy as np
import torch

DENSE, VM, VECZMATXOY, CP, CPFAST, NPLANEMUL, NPLANESUM, HASH = range(8)
_TYPE_FROM_STR = {"dense": 0, "hash": 7, "nplane": 6, "nplanesum": 6, "nplanemul": 5, "vectormatrix": 1, "vm": 1,
                  "veczmatxoy": 2, "cpfast": 4, "cp": 3}
_PRIMES = [1, 2654435761, 805459861, 3674653429]
_U32 = 0xFFFFFFFF


class OracleMeta:
    """Pure-Python restatement of LoDMeta::create_meta (lotd_torch_api.cu:29-230)."""

    def __init__(self, n_input_dims: int, lod_res, lod_n_feats: Sequence[int], lod_types: Sequence[str],
                 hashmap_size: Optional[int] = None, use_smooth_step: Optional[bool] = None):
        D = int(n_input_dims)
        if D not in (2, 3, 4):
            raise RuntimeError("LoTDEncoding: `n_input_dim` must be 2/3/4.")
        res_md = [[int(r)] * D if np.isscalar(r) else [int(v) for v in r] for r in lod_res]
        if not (len(res_md) == len(lod_n_feats) == len(lod_types)):
            raise RuntimeError("LoTDEncoding: Expect los_res, lod_n_feats, lod_str_types to have the same length")
        L = len(res_md)
        if L > 32:
            raise RuntimeError("LoTDEncoding: too many levels")
        nf = [int(v) for v in lod_n_feats]
        for k in (8, 4, 2):
            if all(v % k == 0 for v in nf):
                fpl = k
                break
        else:
            raise RuntimeError("LoTDEncoding: the greatest common divisor of `lod_n_feats` must be at least 2")
        hs = int(hashmap_size or 0)
        self.n_dims_to_encode, self.n_levels, self.n_feat_per_pseudo_lvl = D, L, fpl
        self.interpolation_type = 1 if use_smooth_step else 0
        self.level_res_multidim, self.level_res, self.level_n_feats, self.level_types = [], [], [], []
        self.level_types_str = list(lod_types)
        self.level_sizes, self.level_n_params, self.level_offsets = [], [], []
        self.c_hash_only = True
        acc, accf, max_params = 0, np.float32(0), (_U32 // 2)
        for l in range(L):
            tp = _TYPE_FROM_STR.get(str(lod_types[l]).lower())
            if tp is None:
                raise RuntimeError(f"LoTDEncoding: Invalid lod type: {lod_types[l]}")
            if tp not in (DENSE, HASH):
                self.c_hash_only = False
            R = res_md[l]
            if any(r <= 2 for r in R):
                raise RuntimeError("LoTDEncoding: only support grid resolutions >= 3")
            if tp == DENSE:
                size = int(np.prod(R))
            elif tp in (NPLANEMUL, NPLANESUM):
                size = sum(int(np.prod([R[d] for d in range(D) if d != k])) for k in range(D))
            elif tp == VM:
                if D != 3:
                    raise RuntimeError("LoTDEncoding: VectorMatrix mode only support 3D encoding.")
                size = sum(int(np.prod([R[d] for d in range(D) if d != k])) + R[k] for k in range(D))
            elif tp == VECZMATXOY:
                if D != 3:
                    raise RuntimeError("LoTDEncoding: VecZMatXoY mode only support 3D encoding.")
                size = R[0] * R[1] + R[2]
            elif tp in (CP, CPFAST):
                size = sum(R)
            else:
                if not hs:
                    raise RuntimeError("LoTDEncoding: Hash mode need `hashmap_size`")
                size = hs
            accf = np.float32(accf + np.float32(size) * np.float32(nf[l]))
            if accf > np.float32(max_params):
                raise RuntimeError("LoTDEncoding: param size too large.")
            self.level_res_multidim.append(R)
            self.level_res.append(R[0] if all(r == R[0] for r in R) else 0)
            self.level_n_feats.append(nf[l])
            self.level_types.append(tp)
            self.level_sizes.append(size & _U32)
            self.level_n_params.append((size * nf[l]) & _U32)
            self.level_offsets.append(acc)
            acc = (acc + size * nf[l]) & _U32
        self.level_offsets.append(acc)
        self.n_params = acc
        self.n_encoded_dims = sum(nf)
        if self.n_encoded_dims > 1024:
            raise RuntimeError("LoTDEncoding: total number of features too large. Shoule be <= 1024.")
        self.map_levels, self.map_cnt = [], []
        for l in range(L):
            for j in range(nf[l] // fpl):
                self.map_levels.append(l)
                self.map_cnt.append(j)
        self.n_pseudo_levels = len(self.map_levels)


# ------------------------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------------------------
def _pos(x32: torch.Tensor, xd: torch.Tensor, R: List[int], smooth: bool, forest: bool = False):
    """cell (int64 [N,D]) and (p, dp/dv) in xd's dtype; value path identical to the fp32 kernels, grad path exact.
    forest=True: scale = res (lotd_forest.h:237-240), so cells run 0..res and corner 0 / res+1 belong to the neighbour blocks."""
    scale = torch.tensor([r if forest else r - 2 for r in R], dtype=torch.float64)
    v32 = (x32.double() * scale + 0.5).float()          # == fmaf(x, scale, 0.5f): the double product/sum is exact before rounding
    cell = torch.floor(v32)
    f32 = (v32 - cell).to(xd.dtype)                      # exact in float32
    scale = scale.to(xd.dtype)
    lin = xd * scale
    f = f32 + (lin - lin.detach())                       # value from the fp32 path, derivative d f / d x = scale
    cell = cell.long()
    if smooth:
        p = f * f * (3.0 - 2.0 * f)
    else:
        p = f
    return cell, p


def _corner_bits(D):
    return [[(idx >> d) & 1 for d in range(D)] for idx in range(1 << D)]


def _idx_dense(R, pos):
    D = len(R)
    stride, index = 1, torch.zeros_like(pos[..., 0])
    for d in range(D - 1, -1, -1):
        index = index + pos[..., d] * stride
        stride *= R[d]
    return index & _U32


def _idx_hash(pos, size):
    h = torch.zeros_like(pos[..., 0])
    for d in range(pos.shape[-1]):
        h = h ^ ((pos[..., d] * _PRIMES[d]) & _U32)
    return (h & _U32) % size


def _idx_nplane(R, pos, jump):
    D = len(R)
    stride, index = 1, torch.zeros_like(pos[..., 0])
    for d2 in range(D - 1):
        d3 = d2 + 1 if d2 >= jump else d2
        index = index + pos[..., D - 1 - d3] * stride
        stride *= R[D - 1 - d3]
    return index & _U32


def _idx_nplane_sub(R, pos_plane, jump):
    D = len(R)
    stride, index = 1, torch.zeros_like(pos_plane[..., 0])
    for d2 in range(D - 1):
        d3 = d2 + 1 if d2 >= jump else d2
        index = index + pos_plane[..., D - 2 - d2] * stride
        stride *= R[D - 1 - d3]
    return (jump * stride + index) & _U32


def _idx_cp_line(R, pos_line, line_dim):
    return (sum(R[:line_dim]) + pos_line) & _U32


def _idx_vm(R, pos):
    D = len(R)
    lines, planes = [], []
    acc_line = 0
    for k in range(D):
        lines.append((acc_line + pos[..., k]) & _U32)
        acc_line += R[k]
    acc_plane = 0
    for k in range(D):
        rev_jump = D - 1 - k
        stride, index = 1, torch.zeros_like(pos[..., 0])
        for d2 in range(D - 1):
            d3 = d2 + 1 if d2 >= rev_jump else d2
            index = index + pos[..., D - 1 - d3] * stride
            stride *= R[D - 1 - d3]
        planes.append((acc_line + acc_plane + index) & _U32)
        acc_plane += stride
    return planes, lines


class _Gather:
    """params[base + cell*n_feat + gfo + f] for f in range(F) -> [N, F] float64 (differentiable w.r.t. params)."""

    def __init__(self, params64, base, n_feat, gfo, F):
        self.p, self.base, self.nf, self.gfo, self.F = params64, base, n_feat, gfo, F
        self.ar = torch.arange(F, dtype=torch.int64)

    def __call__(self, cell_idx):
        flat = (self.base + cell_idx * self.nf + self.gfo).unsqueeze(-1) + self.ar
        return self.p[flat]

    def flat_index(self, cell_idx):
        return (self.base + cell_idx * self.nf + self.gfo).unsqueeze(-1) + self.ar


def _corner_value(tp, R, size, pos, G: _Gather):
    if tp == DENSE:
        return G(_idx_dense(R, pos))
    if tp == HASH:
        return G(_idx_hash(pos, size))
    if tp == VM:
        planes, lines = _idx_vm(R, pos)
        return sum(G(planes[k]) * G(lines[k]) for k in range(len(R)))
    if tp == VECZMATXOY:
        line = pos[..., 2] & _U32
        plane = (R[2] + pos[..., 1] + pos[..., 0] * R[0]) & _U32
        return G(plane) * G(line)
    if tp == NPLANEMUL:
        out = G(_idx_nplane(R, pos, 0))
        for j in range(1, len(R)):
            out = out * G(_idx_nplane(R, pos, j))
        return out
    if tp == CP:
        out = G(_idx_cp_line(R, pos[..., 0], 0))
        for k in range(1, len(R)):
            out = out * G(_idx_cp_line(R, pos[..., k], k))
        return out
    raise ValueError(tp)


def _level_value(tp, R, size, cell, p, G: _Gather):
    """[N, F] float64 value of one pseudo level."""
    D = len(R)
    N = cell.shape[0]
    if tp in (DENSE, HASH, VM, VECZMATXOY, NPLANEMUL, CP):
        out = torch.zeros(N, G.F, dtype=p.dtype)
        for bits in _corner_bits(D):
            w = torch.ones(N, dtype=p.dtype)
            pos = cell.clone()
            for d in range(D):
                if bits[d]:
                    w = w * p[:, d]
                    pos[:, d] += 1
                else:
                    w = w * (1.0 - p[:, d])
            out = out + w.unsqueeze(-1) * _corner_value(tp, R, size, pos, G)
        return out
    if tp == NPLANESUM:
        out = torch.zeros(N, G.F, dtype=p.dtype)
        for j in range(D):
            dims3 = [d for d in range(D) if d != j]
            for bits in _corner_bits(D - 1):
                w = torch.ones(N, dtype=p.dtype)
                pp = cell[:, dims3].clone()
                for d2, d3 in enumerate(dims3):
                    if bits[d2]:
                        w = w * p[:, d3]
                        pp[:, d2] += 1
                    else:
                        w = w * (1.0 - p[:, d3])
                out = out + w.unsqueeze(-1) * G(_idx_nplane_sub(R, pp, j))
        return out
    if tp == CPFAST:
        out = torch.ones(N, G.F, dtype=p.dtype)
        for k in range(D):
            L_ = G(_idx_cp_line(R, cell[:, k], k))
            R_ = G(_idx_cp_line(R, cell[:, k] + 1, k))
            out = out * ((1.0 - p[:, k]).unsqueeze(-1) * L_ + p[:, k].unsqueeze(-1) * R_)
        return out
    raise ValueError(tp)


# ------------------------------------------------------------------------------------------------------------------
# forest (multi-block) variant: csrc/lotd/include/lotd/lotd_forest.h:31-139 (forward), :319-407 (dL/dparam), csrc/forest/forest.h:25-97
# ------------------------------------------------------------------------------------------------------------------
class OracleForest:
    """What the reference's ForestMetaRef carries (forest.h:59-97): SPC octree bytes, their exclusive child-count sums, the blocks'
    integer coordinates and the octree level they live on."""

    def __init__(self, octree, exsum, block_ks, level, level_poffset, continuity_enabled=True):
        self.octree = np.asarray(octree, dtype=np.uint8)
        self.exsum = np.asarray(exsum, dtype=np.int64)
        self.block_ks = np.asarray(block_ks, dtype=np.int64)
        self.level, self.level_poffset, self.continuity_enabled = int(level), int(level_poffset), bool(continuity_enabled)

    def map_block_ind(self, k: np.ndarray) -> np.ndarray:
        """== ForestMetaRef::map_block_ind / identify (forest.h:25-57,88-95), vectorised over k [M,3]: walk the octree from the root,
        at every level test the child bit and jump to `exsum[ord] + (inclusive count of set bits up to the child)`; -1 if absent."""
        k = np.asarray(k, dtype=np.int64)
        M = k.shape[0]
        maxval = (1 << self.level) - 1
        ok = np.all((k >= 0) & (k <= maxval), axis=1)
        ord_ = np.zeros(M, dtype=np.int64)
        kk = np.where(ok[:, None], k, 0)
        for l in range(self.level):
            depth = self.level - l - 1
            child = (((kk[:, 0] >> depth) & 1) << 2) | (((kk[:, 1] >> depth) & 1) << 1) | ((kk[:, 2] >> depth) & 1)
            bits = self.octree[ord_].astype(np.int64)
            has = ((bits >> child) & 1) == 1
            masked = bits & ((2 << child) - 1)
            cnt = np.zeros(M, dtype=np.int64)
            for b in range(8):
                cnt += (masked >> b) & 1
            ok &= has
            ord_ = np.where(ok, self.exsum[ord_] + cnt, 0)
        return np.where(ok, ord_ - self.level_poffset, -1)


def _forest_level_value(tp, R, size, cell, p, params64, b, forest: OracleForest, block_base_of, lvl_off, n_feat, gfo, F):
    """n-linear value of one pseudo level with the cross-block continuity rule (lotd_forest.h:53-88): corner coordinate 0 is the
    left neighbour's res-1, res+1 the right neighbour's 0, 1..res are the block's own 0..res-1; a corner whose block is not in
    the forest (or any remapped corner when continuity is disabled) contributes zero."""
    if tp not in (DENSE, HASH, VM, NPLANEMUL, CP):
        raise RuntimeError("lotd-forest supports Dense / VM / NPlaneMul / CP / Hash levels only (lotd_forest.h:263-311)")
    D, N = 3, cell.shape[0]
    Rt = torch.tensor(R, dtype=torch.int64)
    out = torch.zeros(N, F, dtype=p.dtype)
    bk = torch.from_numpy(forest.block_ks)[b]                       # [N,3]
    for bits in _corner_bits(D):
        w = torch.ones(N, dtype=p.dtype)
        pos = cell.clone()
        for d in range(D):
            if bits[d]:
                w = w * p[:, d]
                pos[:, d] += 1
            else:
                w = w * (1.0 - p[:, d])
        left, right = pos == 0, pos == (Rt + 1)
        kloc = bk - left.long() + right.long()
        ploc = torch.where(left, Rt - 1, torch.where(right, torch.zeros_like(pos), pos - 1))
        changed = (left | right).any(dim=1)
        bl = b.clone()
        valid = torch.ones(N, dtype=torch.bool)
        if changed.any():
            if not forest.continuity_enabled:
                valid = ~changed
            else:
                nb = torch.from_numpy(forest.map_block_ind(kloc[changed].numpy()))
                bl[changed] = nb.clamp_min(0)
                valid[changed] = nb >= 0
        G = _Gather(params64, block_base_of(bl) + lvl_off, n_feat, gfo, F)
        val = _corner_value(tp, R, size, ploc, G)
        out = out + (w * valid.to(p.dtype)).unsqueeze(-1) * val
    return out


# ------------------------------------------------------------------------------------------------------------------
# public API
# ------------------------------------------------------------------------------------------------------------------
def _point_batches(N, batch_inds, batch_data_size):
    if batch_inds is not None:
        b = batch_inds.long().clone()
        valid = b >= 0
        b[~valid] = 0
        return b, valid
    if batch_data_size:
        return torch.arange(N, dtype=torch.int64) // int(batch_data_size), torch.ones(N, dtype=torch.bool)
    return torch.zeros(N, dtype=torch.int64), torch.ones(N, dtype=torch.bool)


def encode_levels(meta, x: torch.Tensor, params: torch.Tensor, batch_inds=None, batch_offsets=None, batch_data_size=0,
                  max_level=None, dtype=torch.float64, forest: Optional["OracleForest"] = None):
    """Returns a list with one [N, F_pl] float64 tensor per pseudo level (differentiable w.r.t. x / params if they
    are float64 leaves requiring grad; otherwise they are promoted)."""
    D = meta.n_dims_to_encode
    N = x.shape[0]
    x32 = x.detach().float()
    xd = x if x.dtype == dtype else x.to(dtype)
    pd = params if params.dtype == dtype else params.to(dtype)
    max_level = meta.n_levels if max_level is None else int(max_level)
    b, valid = _point_batches(N, batch_inds, batch_data_size)
    boff = batch_offsets.long()[b] if batch_offsets is not None else b * meta.n_params
    F = meta.n_feat_per_pseudo_lvl
    smooth = int(meta.interpolation_type) == 1
    outs = []
    cache = {}
    for pl in range(meta.n_pseudo_levels):
        lvl = meta.map_levels[pl]
        if lvl > max_level or max_level <= -1:
            outs.append(torch.zeros(N, F, dtype=dtype))
            continue
        R = list(meta.level_res_multidim[lvl])
        if lvl not in cache:
            cache[lvl] = _pos(x32, xd, R, smooth, forest is not None)
        cell, p = cache[lvl]
        if forest is not None:
            base_of = (lambda bl: batch_offsets.long()[bl]) if batch_offsets is not None else (lambda bl: bl * meta.n_params)
            val = _forest_level_value(int(meta.level_types[lvl]), R, meta.level_sizes[lvl], cell, p, pd, b, forest, base_of,
                                      meta.level_offsets[lvl], meta.level_n_feats[lvl], meta.map_cnt[pl] * F, F)
            outs.append(val * valid.unsqueeze(-1))
            continue
        G = _Gather(pd, boff + meta.level_offsets[lvl], meta.level_n_feats[lvl], meta.map_cnt[pl] * F, F)
        val = _level_value(int(meta.level_types[lvl]), R, meta.level_sizes[lvl], cell, p, G)
        outs.append(val * valid.unsqueeze(-1))
    return outs


def encode(meta, x, params, **kw) -> torch.Tensor:
    """y [N, n_encoded_dims] float64."""
    return torch.cat(encode_levels(meta, x, params, **kw), dim=-1)


def fwd_dydx(meta, x, params, **kw):
    """(y [N,E], dy_dx [N,E,D]) float64, no graph kept."""
    xd = x.detach().double().requires_grad_(True)
    y = encode(meta, xd, params.detach().double(), **kw)
    cols = []
    for j in range(y.shape[1]):
        (g,) = torch.autograd.grad(y[:, j].sum(), xd, retain_graph=True, allow_unused=True)
        cols.append(torch.zeros_like(xd) if g is None else g)
    return y.detach(), torch.stack(cols, 1)


def bwd(meta, dL_dy, x, params, **kw):
    """(dL_dx [N,D], dL_dparam [n_params*B]) float64 via autograd of the oracle forward."""
    xd = x.detach().double().requires_grad_(True)
    pd = params.detach().double().requires_grad_(True)
    y = encode(meta, xd, pd, **kw)
    gx, gp = torch.autograd.grad(y, [xd, pd], dL_dy.double(), allow_unused=True)
    return (torch.zeros_like(xd) if gx is None else gx), (torch.zeros_like(pd) if gp is None else gp)


_SECOND_ORDER_DX_TYPES = (DENSE, HASH, VM, VECZMATXOY)  # lotd_encoding.h:1245-1284
_SECOND_ORDER_DX_TYPES_FOREST = (DENSE, HASH, VM)       # lotd_forest.h:1022-1050


def bwd_bwd_input(meta, dL_ddLdx, dL_dy, x, params, **kw):
    """(dL_ddLdy [N,E], dL_dparam, dL_dx [N,D]) float64: gradients of <dL_dx(dL_dy, x, params), dL_ddLdx>."""
    xd = x.detach().double().requires_grad_(True)
    pd = params.detach().double().requires_grad_(True)
    gy = dL_dy.detach().double().requires_grad_(True)
    levels = encode_levels(meta, xd, pd, **kw)
    F = meta.n_feat_per_pseudo_lvl
    total_dldx = 0
    dx_supported = 0
    for pl, yl in enumerate(levels):
        if not yl.requires_grad:
            continue
        (dldx,) = torch.autograd.grad(yl, xd, gy[:, pl * F:(pl + 1) * F], create_graph=True, allow_unused=True)
        if dldx is None:
            continue
        total_dldx = total_dldx + dldx
        if int(meta.level_types[meta.map_levels[pl]]) in (_SECOND_ORDER_DX_TYPES_FOREST if kw.get("forest") is not None else _SECOND_ORDER_DX_TYPES):
            dx_supported = dx_supported + dldx
    s_all = (total_dldx * dL_ddLdx.double()).sum()
    g_gy, g_p = torch.autograd.grad(s_all, [gy, pd], retain_graph=True, allow_unused=True)
    g_x = None
    if torch.is_tensor(dx_supported):
        (g_x,) = torch.autograd.grad((dx_supported * dL_ddLdx.double()).sum(), xd, allow_unused=True)
    z = lambda g, ref: torch.zeros_like(ref) if g is None else g
    return z(g_gy, gy), z(g_p, pd), z(g_x, xd)


def grid_index(meta, x, batch_inds=None, batch_offsets=None, batch_data_size=0, max_level=None) -> torch.Tensor:
    """int64 [N, n_enc, 2^D] (lotd_encoding.h:1300-1433), uint32 wrap-around arithmetic, zeros where skipped."""
    D = meta.n_dims_to_encode
    N = x.shape[0]
    x32 = x.detach().float()
    max_level = meta.n_levels if max_level is None else int(max_level)
    b, valid = _point_batches(N, batch_inds, batch_data_size)
    boff = batch_offsets.long()[b] if batch_offsets is not None else b * meta.n_params
    F = meta.n_feat_per_pseudo_lvl
    out = torch.zeros(N, meta.n_encoded_dims, 1 << D, dtype=torch.int64)
    for pl in range(meta.n_pseudo_levels):
        lvl = meta.map_levels[pl]
        if lvl > max_level or max_level <= -1:
            continue
        R = list(meta.level_res_multidim[lvl])
        cell, _ = _pos(x32, x32.double(), R, False)
        tp = int(meta.level_types[lvl])
        for idx, bits in enumerate(_corner_bits(D)):
            pos = cell + torch.tensor(bits, dtype=torch.int64)
            ci = _idx_dense(R, pos) if tp == DENSE else _idx_hash(pos, meta.level_sizes[lvl])
            ind = (boff + meta.level_offsets[lvl] + ci * meta.level_n_feats[lvl] + meta.map_cnt[pl] * F) & _U32
            for f in range(F):
                out[:, pl * F + f, idx] = torch.where(valid, (ind + f) & _U32, torch.zeros_like(ind))
    return out


def grid_sample_dense_reference(param: torch.Tensor, x01: torch.Tensor, R: int) -> torch.Tensor:
    """The reference's own pure-PyTorch path for a cubic 3-D Dense level (lotd_helpers.py:327-346 composed with
    LoTDEncoding's [-1,1] -> [0,1] mapping, lotd_encoding.py:162): param [R,R,R,F], x01 [N,3] in [0,1] -> [N,F]."""
    import torch.nn.functional as F_
    rel = (x01 * 2.0 - 1.0) * ((R - 2.0) / (R - 1.0))
    grid = rel[..., [2, 1, 0]].view(1, 1, 1, -1, 3)
    vol = param.permute(3, 0, 1, 2).unsqueeze(0)
    out = F_.grid_sample(vol, grid, align_corners=True, padding_mode="zeros")
    return out.view(param.shape[-1], -1).t()


# ------------------------------------------------------------------------------------------------------------------
# float64 arbiter for full-size gradient tables (Dense / Hash levels): bincount instead of autograd
# ------------------------------------------------------------------------------------------------------------------
def bwd_param_hash_f64(meta, dL_dy, x, chunk=1 << 19, batch_inds=None, batch_data_size=0, n_scenes=1):
    """dL/dparam [n_scenes * n_params] in float64 for Dense / Hash-only metas plus, per entry, the sum of |terms| that met there.

    Same maths as `bwd` (reference scatter lotd_cuda.h:494-829 over the n-linear weights of linear_interpolate.cuh:92-150; cell and
    fraction from the fp32 `_pos` path), evaluated with np.bincount so that BASELINE.json's full sizes (4 Mi points) take seconds.
    The |term| sum bounds what fp32 accumulation ORDER can change: |fl32(sum) - sum| <= n * eps32 * sum|terms| -- the element-wise
    tolerance the full-size tests grant a build on top of its distance to this table."""
    assert meta.c_hash_only
    D, F = meta.n_dims_to_encode, meta.n_feat_per_pseudo_lvl
    xn = x.detach().float()
    g64 = dL_dy.detach().double().numpy()
    N = xn.shape[0]
    total = meta.n_params * n_scenes
    out, mag = np.zeros(total, dtype=np.float64), np.zeros(total, dtype=np.float64)
    smooth = meta.interpolation_type == 1
    for s in range(0, N, chunk):
        xs = xn[s:s + chunk]
        n = xs.shape[0]
        if batch_inds is not None:
            sc = np.asarray(batch_inds[s:s + chunk]).astype(np.int64)
        elif batch_data_size:
            sc = (np.arange(s, s + n) // batch_data_size).astype(np.int64)
        else:
            sc = np.zeros(n, dtype=np.int64)
        live = sc >= 0
        for pl, lvl in enumerate(meta.map_levels):
            R, size, nf = meta.level_res_multidim[lvl], meta.level_sizes[lvl], meta.level_n_feats[lvl]
            cell, p = _pos(xs, xs.double(), R, smooth)
            cell, p = cell.numpy(), p.numpy()
            lnp = meta.level_n_params[lvl]
            acc, acc_mag = np.zeros(n_scenes * lnp, dtype=np.float64), np.zeros(n_scenes * lnp, dtype=np.float64)
            for bits in _corner_bits(D):
                w = np.ones(n, dtype=np.float64)
                pos = cell.copy()
                for d in range(D):
                    if bits[d]:
                        w = w * p[:, d]
                        pos[:, d] += 1
                    else:
                        w = w * (1.0 - p[:, d])
                tpos = torch.from_numpy(pos)
                idx = (_idx_dense(R, tpos) if meta.level_types[lvl] == DENSE else _idx_hash(tpos, size)).numpy()
                flat = (sc * lnp + idx * nf + meta.map_cnt[pl] * F)[live]
                for f in range(F):
                    t = (w * g64[s:s + n, pl * F + f])[live]
                    acc += np.bincount(flat + f, weights=t, minlength=n_scenes * lnp)
                    acc_mag += np.bincount(flat + f, weights=np.abs(t), minlength=n_scenes * lnp)
            o = meta.level_offsets[lvl]
            out.reshape(n_scenes, meta.n_params)[:, o:o + lnp] += acc.reshape(n_scenes, lnp)
            mag.reshape(n_scenes, meta.n_params)[:, o:o + lnp] += acc_mag.reshape(n_scenes, lnp)
    return out, mag
