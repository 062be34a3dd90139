"""CPU oracle for pack_ops -- TEST INFRASTRUCTURE ONLY (numpy; sequential per-pack loops).

Each function restates one reference kernel of csrc/pack_ops/pack_ops_cuda.cu with the SAME sequential order of
floating-point operations (numpy scalars of the tensor's dtype are IEEE-exact), so order-sensitive decisions -- the
``T < eps`` early stop and ``alpha <= thre`` skip of the alpha composite -- are reproduced bit for bit.

Pinned by: the literal known-answer vectors of the reference's own unit test
(nr3d_lib/graphics/pack_ops/unit_test.py:547-563, exclusive cumsum / cumprod as produced by kaolin), the
torch-equivalence identities of unit_test.py:188-241, and golden vectors from the reference CUDA build (tests/golden/).
"""
import numpy as np


def _ranges(pack_infos):
    pi = np.asarray(pack_infos, dtype=np.int64)
    return [(int(b), int(max(n, 0))) for b, n in pi]


def _as2d(a):
    a = np.asarray(a)
    return (a.reshape(a.shape[0], 1), True) if a.ndim == 1 else (a, False)


def packed_sum(feats, pack_infos):
    """pack_ops_cuda.cu:798-824 (empty packs give 0; the reference reads feats[begin] there)."""
    f, squeeze = _as2d(feats)
    out = np.zeros((len(pack_infos), f.shape[1]), dtype=f.dtype)
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        for j in range(f.shape[1]):
            if n == 0:
                continue
            r = f[b, j]
            for i in range(b + 1, b + n):
                r = f.dtype.type(r + f[i, j])
            out[p, j] = r
    return out[:, 0] if squeeze else out


def _scan(feats, pack_infos, exclusive, reverse, op, ident):
    f, squeeze = _as2d(feats)
    out = np.zeros_like(f)
    T = f.dtype.type
    for (b, n) in _ranges(pack_infos):
        if n == 0:
            continue
        order = range(b + n - 1, b - 1, -1) if reverse else range(b, b + n)
        for j in range(f.shape[1]):
            acc = T(ident)
            for i in order:
                if exclusive:
                    out[i, j] = acc
                    acc = T(op(acc, f[i, j]))
                else:
                    acc = T(op(acc, f[i, j]))
                    out[i, j] = acc
    return out[:, 0] if squeeze else out


def packed_cumsum(feats, pack_infos, exclusive=False, reverse=False):
    """pack_ops_cuda.cu:981-1046: inclusive, or right-shifted with a leading 0."""
    return _scan(feats, pack_infos, exclusive, reverse, lambda a, b: a + b, 0)


def packed_cumprod(feats, pack_infos, exclusive=False, reverse=False, bug_compat=False):
    """pack_ops_cuda.cu:864-929.  ``exclusive`` follows the documented semantics (leading 1, pack_ops.py:149);
    ``bug_compat`` reproduces the reference CUDA output for exclusive=True: all zeros."""
    if exclusive and bug_compat:
        return np.zeros_like(np.asarray(feats))
    return _scan(feats, pack_infos, exclusive, reverse, lambda a, b: a * b, 1)


def packed_diff(feats, pack_infos, appends=None, last_fill=None):
    """pack_ops_cuda.cu:1098-1140."""
    f, squeeze = _as2d(feats)
    out = np.zeros_like(f)
    ap = None if appends is None else _as2d(appends)[0]
    lf = None if last_fill is None else _as2d(last_fill)[0]
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        if n == 0:
            continue
        out[b:b + n - 1] = f[b + 1:b + n] - f[b:b + n - 1]
        if ap is not None:
            out[b + n - 1] = ap[p] - f[b + n - 1]
        elif lf is not None:
            out[b + n - 1] = lf[p]
    return out[:, 0] if squeeze else out


def packed_backward_diff(feats, pack_infos, prepends=None, first_fill=None):
    """pack_ops_cuda.cu:1142-1184."""
    f, squeeze = _as2d(feats)
    out = np.zeros_like(f)
    pp = None if prepends is None else _as2d(prepends)[0]
    ff = None if first_fill is None else _as2d(first_fill)[0]
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        if n == 0:
            continue
        out[b + 1:b + n] = f[b + 1:b + n] - f[b:b + n - 1]
        if pp is not None:
            out[b] = f[b] - pp[p]
        elif ff is not None:
            out[b] = ff[p]
    return out[:, 0] if squeeze else out


_OPS = {0: np.add, 1: np.subtract, 2: np.multiply, 3: np.divide, 5: np.greater, 6: np.greater_equal, 7: np.less,
        8: np.less_equal, 9: np.equal, 10: np.not_equal}


def packed_binary(op, feats, other, pack_infos):
    """pack_ops_cuda.cu:1960-2249: broadcast one operand per pack."""
    f, squeeze = _as2d(feats)
    o = _as2d(other)[0]
    out = np.zeros(f.shape, dtype=f.dtype if op < 4 else np.bool_)
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        if n:
            with np.errstate(all="ignore"):
                out[b:b + n] = _OPS[op](f[b:b + n], o[p])
    return out[:, 0] if squeeze else out


def alpha_to_vw_forward(alphas, pack_infos, early_stop_eps, alpha_thre):
    """pack_ops_cuda.cu:1735-1793 -> (weights, num_steps int64 [P], selector bool [S])."""
    a = np.asarray(alphas)
    T_ = a.dtype.type
    w = np.zeros_like(a)
    sel = np.zeros(a.shape, dtype=np.bool_)
    cnt = np.zeros(len(pack_infos), dtype=np.int64)
    eps, thre, one = T_(early_stop_eps), T_(alpha_thre), np.float32(1.0)
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        T = T_(1.0)
        for j in range(n):
            if T < eps:
                break
            al = a[b + j]
            if al <= thre:
                continue
            w[b + j] = T_(al * T)
            T = T_(T * T_(one - al)) if a.dtype != np.float16 else T_(np.float32(T) * (one - np.float32(al)))
            sel[b + j] = True
            cnt[p] += 1
    return w, cnt, sel


def alpha_to_vw_backward(weights, grad_weights, alphas, pack_infos, early_stop_eps, alpha_thre):
    """pack_ops_cuda.cu:1795-1848 (note ``alpha < thre`` here vs ``<=`` in the forward)."""
    a, w, gw = np.asarray(alphas), np.asarray(weights), np.asarray(grad_weights)
    T_ = a.dtype.type
    ga = np.zeros_like(a)
    eps, thre = T_(early_stop_eps), T_(alpha_thre)
    for (b, n) in _ranges(pack_infos):
        accum = T_(0)
        for j in range(n):
            accum = T_(accum + gw[b + j] * w[b + j])
        T = T_(1.0)
        for j in range(n):
            if T < eps:
                break
            al = a[b + j]
            if al < thre:
                continue
            denom = np.float32(max(np.float32(np.float32(1.0) - np.float32(al)) if a.dtype != np.float64 else np.float32(1.0 - al), np.float32(1e-10)))
            ga[b + j] = T_((gw[b + j] * T - accum) / T_(denom))
            accum = T_(accum - gw[b + j] * w[b + j])
            T = T_(T * T_(np.float32(1.0) - al))
    return ga


def interleave_linstep(starts, num_steps, steps, dtype=None):
    """pack_ops_cuda.cu:47-83 -> (out, nidx)."""
    starts = np.asarray(starts)
    T_ = starts.dtype.type
    steps = np.broadcast_to(np.asarray(steps, dtype=starts.dtype), starts.shape)
    out, nidx = [], []
    for p, n in enumerate(np.asarray(num_steps, dtype=np.int64)):
        for j in range(int(n)):
            if starts.dtype == np.float32:
                # nvcc contracts `start + j*step` into one FMA (single rounding): evaluate exactly in float64, round once
                out.append(T_(np.float64(starts[p]) + np.float64(T_(j)) * np.float64(steps[p])))
            else:
                out.append(T_(starts[p] + T_(j) * steps[p]))
            nidx.append(p)
    return np.asarray(out, dtype=starts.dtype), np.asarray(nidx, dtype=np.int64)


def sample_step_wrt_depth_clamped(near, far, max_steps, dt_gamma, min_step, max_step):
    """pack_ops_cuda.cu:480-545 -> (t_samples, deltas, nidx, pack_infos)."""
    near, far = np.asarray(near), np.asarray(far)
    T_ = near.dtype.type
    g, lo, hi = T_(dt_gamma), T_(min_step), T_(max_step)
    clamp = lambda v: lo if v < lo else (hi if hi < v else v)
    ts, ds, ni, counts = [], [], [], []
    for p in range(near.shape[0]):
        t, n = near[p], 0
        while t <= far[p] and n < max_steps:
            dt = clamp(T_(t * g))
            ts.append(t); ds.append(dt); ni.append(p)
            t = T_(t + dt)
            n += 1
        counts.append(n)
    counts = np.asarray(counts, dtype=np.int64)
    pack_infos = np.stack([np.cumsum(counts) - counts, counts], 1)
    return (np.asarray(ts, dtype=near.dtype), np.asarray(ds, dtype=near.dtype), np.asarray(ni, dtype=np.int64), pack_infos)


def mark_pack_boundaries(ids):
    """pack_ops_cuda.cu:2765-2782."""
    ids = np.asarray(ids)
    out = np.ones(ids.shape[0], dtype=np.int32)
    out[1:] = (ids[1:] != ids[:-1]).astype(np.int32)
    return out


def _lower_bound(val, data):
    """binary_search_unsafe (pack_ops_cuda.cu:1336-1362): first index with data[idx] >= val."""
    first, count = 0, len(data)
    while count > 0:
        step = count // 2
        it = first + step
        if data[it] < val:
            first, count = it + 1, count - step - 1
        else:
            count = step
    return first


def packed_searchsorted(bins, vals, pack_infos, val_pack_infos=None):
    """pack_ops_cuda.cu:1374-1407: begin + min(lower_bound, len-1); vals [P, n] or packed by val_pack_infos."""
    bins, vals = np.asarray(bins), np.asarray(vals)
    out = np.full(vals.shape, -1, dtype=np.int64)
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        if val_pack_infos is None:
            qs = [(p, i) for i in range(vals.shape[1])]
        else:
            qb, qn = _ranges(val_pack_infos)[p]
            qs = [(qb + i,) for i in range(qn)]
        for q in qs:
            pos = min(_lower_bound(vals[q], bins[b:b + n]), n - 1) if n else 0
            out[q] = b + pos
    return out


def packed_invert_cdf(bins, cdfs, u_vals, pack_infos):
    """pack_ops_cuda.cu:1633-1681 -> (samples, bin_idx); `a + b*c` evaluated with one rounding (nvcc FMA contraction)."""
    bins, cdfs, u = np.asarray(bins), np.asarray(cdfs), np.asarray(u_vals)
    T_ = bins.dtype.type
    samples = np.zeros_like(u)
    idx = np.full(u.shape, -1, dtype=np.int64)
    eps = T_(1.0e-5)
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        bn, cd = bins[b:b + n], cdfs[b:b + n]
        for i in range(u.shape[1]):
            pos = min(_lower_bound(u[p, i], cd), n - 1) if n else 0
            idx[p, i] = pos + b
            if pos == 0:
                samples[p, i] = bn[0]
            else:
                pmf = T_(cd[pos] - cd[pos - 1])
                if pmf < eps:
                    samples[p, i] = bn[pos - 1]
                else:
                    r = T_(T_(u[p, i] - cd[pos - 1]) / pmf)
                    d = T_(bn[pos] - bn[pos - 1])
                    samples[p, i] = T_(np.float64(bn[pos - 1]) + np.float64(r) * np.float64(d)) if bins.dtype == np.float32 else T_(bn[pos - 1] + r * d)
    return samples, idx


def try_merge_two_packs_sorted_aligned(vals_a, pack_infos_a, vals_b, pack_infos_b):
    """pack_ops_cuda.cu:1505-1571 -> (pidx_a, pidx_b, merged pack_infos)."""
    va, vb = np.asarray(vals_a), np.asarray(vals_b)
    n = np.asarray(pack_infos_a)[:, 1] + np.asarray(pack_infos_b)[:, 1]
    pm = pack_infos_from_counts(n)
    pa = np.zeros(va.shape[0], dtype=np.int64)
    pb = np.zeros(vb.shape[0], dtype=np.int64)
    for p, ((ab, an), (bb, bn)) in enumerate(zip(_ranges(pack_infos_a), _ranges(pack_infos_b))):
        out_begin = int(pm[p, 0])
        lb = [_lower_bound(vb[bb + j], va[ab:ab + an]) for j in range(bn)]
        cnt = np.zeros(an, dtype=np.int64)
        for i in lb:
            if i < an:
                cnt[i] += 1
        if an:
            pa[ab:ab + an] = out_begin + np.arange(an) + np.cumsum(cnt)
        acc, last = 1, -1
        for j, i in enumerate(lb):
            acc = acc + 1 if i == last else 0
            pb[bb + j] = acc + (out_begin if i == 0 else pa[ab + i - 1] + 1)
            last = i
    return pa, pb, pm


def packed_sort(vals, pack_infos):
    """ascending per-pack sort -> (sorted vals, permuted global indices); equal keys: any order is acceptable."""
    v = np.asarray(vals).copy()
    idx = np.arange(v.shape[0], dtype=np.int64)
    for (b, n) in _ranges(pack_infos):
        o = np.argsort(v[b:b + n], kind="stable")
        v[b:b + n] = v[b:b + n][o]
        idx[b:b + n] = idx[b:b + n][o]
    return v, idx


def packed_matmul(feats, other, pack_infos):
    """pack_ops_cuda.cu:2060-2085: sequential dot products in the tensor's dtype."""
    f, o = np.asarray(feats), np.asarray(other)
    out = np.zeros((f.shape[0], o.shape[1]), dtype=f.dtype)
    T_ = f.dtype.type
    for p, (b, n) in enumerate(_ranges(pack_infos)):
        for i in range(b, b + n):
            for j in range(o.shape[1]):
                acc = T_(0)
                for k in range(f.shape[1]):
                    acc = T_(np.float64(acc) + np.float64(f[i, k]) * np.float64(o[p, j, k])) if f.dtype == np.float32 else T_(acc + f[i, k] * o[p, j, k])
                out[i, j] = acc
    return out


def sample_step_in_packed_segments(near, far, entry, exit_, seg_pack_infos, max_steps, dt_gamma, min_step, max_step):
    """pack_ops_cuda.cu:606-721 -> (t_samples, deltas, sidx, nidx, pack_infos)."""
    near, far, entry, exit_ = (np.asarray(a) for a in (near, far, entry, exit_))
    T_ = near.dtype.type
    g, lo, hi = T_(dt_gamma), T_(min_step), T_(max_step)
    clamp = lambda v: lo if v < lo else (hi if hi < v else v)
    ts, ds, si, ni, counts = [], [], [], [], []
    for p, (sb, sn) in enumerate(_ranges(seg_pack_infos)):
        t, n = near[p], 0
        for i in range(sb, sb + sn):
            if entry[i] >= far[p] or exit_[i] <= near[p]:
                break
            while True:                       # do { t += min_step } while (t < entry)
                t = T_(t + lo)
                if not (t < entry[i]):
                    break
            while t <= exit_[i] and t <= far[p] and n < max_steps:
                dt = clamp(T_(t * g))
                ts.append(t); ds.append(dt); si.append(i); ni.append(p)
                t = T_(t + dt)
                n += 1
        counts.append(n)
    counts = np.asarray(counts, dtype=np.int64)
    pack_infos = np.stack([np.cumsum(counts) - counts, counts], 1)
    return (np.asarray(ts, dtype=near.dtype), np.asarray(ds, dtype=near.dtype), np.asarray(si, dtype=np.int64),
            np.asarray(ni, dtype=np.int64), pack_infos)


def octree_mark_consecutive_segments(pidx, pack_infos, points, offset_fix=False):
    """pack_ops_cuda.cu:2807-2841; offset_fix=False keeps the reference's un-offset walk over `point_indices`."""
    pidx, points = np.asarray(pidx), np.asarray(points).astype(np.int64)
    ms = np.zeros(pidx.shape[0], dtype=bool)
    me = np.zeros(pidx.shape[0], dtype=bool)
    for (b, n) in _ranges(pack_infos):
        if n == 0:
            continue
        pi = pidx[b:] if offset_fix else pidx
        ms[b] = True
        for j in range(1, n):
            if np.abs(points[pi[j]] - points[pi[j - 1]]).sum() > 1:
                me[b + j - 1] = True
                ms[b + j] = True
        me[b + n - 1] = True
    return ms, me


def pack_infos_from_counts(counts):
    c = np.asarray(counts, dtype=np.int64)
    return np.stack([np.cumsum(c) - c, c], 1)
