"""Packet model of the headline backward with CTA-level shared-memory tiles (CPU-only, numpy) -- companion of sim_sectors.py.

Question: how many L2 reduction packets per point are left if every CTA first sums the contributions of its points in a shared-memory
tile per level (indexed by the corner's cell coordinates relative to the CTA's bounding box) and then flushes each touched tile entry with
ONE reduction?  A level is tiled when its corner bounding box fits the CTA's tile budget; other levels keep the warp-level run merge of the
shipped kernel (cost model of sim_sectors.py: lanes of one instruction that hit different entries of a 32-byte sector share a packet).
The flush sweeps a tile in z-fastest order with 32 lanes per instruction, so entries of one sector leave in one packet.

    python scripts/sim_tiles.py [cta_points] [cap_entries] > profiles/r2_tile_simulation.txt
"""
import sys
import numpy as np

N = 1 << 22
T = 1 << 19
CTA = int(sys.argv[1]) if len(sys.argv) > 1 else 128
CAP = int(sys.argv[2]) if len(sys.argv) > 2 else 3072
rs = np.random.RandomState(0)
x = rs.rand(N, 3).astype(np.float32)
res_list = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
b = np.minimum((x * 128).astype(np.int64), 127)


def order_key(name):
    bx, by, bz = b[:, 0], b[:, 1], b[:, 2]
    if name == "xfast":
        return (bz * 128 + by) * 128 + bx
    if name.startswith("brick"):
        sx, sy, sz = [int(v) for v in name[5:].split("x")]
        nbx, nby = 128 // sx, 128 // sy
        brick = ((bz // sz) * nby + (by // sy)) * nbx + (bx // sx)
        inner = ((bz % sz) * sy + (by % sy)) * sx + (bx % sx)
        return brick * (sx * sy * sz) + inner
    if name == "morton":
        k = np.zeros(N, np.int64)
        for i in range(7):
            k |= ((bx >> i) & 1) << (3 * i) | ((by >> i) & 1) << (3 * i + 1) | ((bz >> i) & 1) << (3 * i + 2)
        return k
    raise ValueError(name)


def warp_packets(c, R, dense):
    """all-runs-merged packets of the shipped kernel (see sim_sectors.py), per level, summed over all warps."""
    W = N // 16
    key = (c[:, 0] | (c[:, 1] << 10) | (c[:, 2] << 20)).reshape(W, 16)
    head = np.ones((W, 16), bool)
    head[:, 1:] = key[:, 1:] != key[:, :-1]
    valid = np.repeat(head, 2, axis=1)
    idx = np.arange(32)[None, :].repeat(W, 0)
    total = 0
    for q in range(4):
        ents = []
        for side in range(2):
            if dense:
                dx, dy = q & 1, q >> 1
                e = ((c[:, 0] + dx) * R + (c[:, 1] + dy)) * R + c[:, 2] + side
            else:
                dy, dz = q & 1, q >> 1
                hyz = (((c[:, 1] + dy) * 2654435761) ^ ((c[:, 2] + dz) * 805459861)) & 0xFFFFFFFF
                e = ((c[:, 0] + side) ^ hyz) & (T - 1)
            ents.append(e.reshape(W, 16))
        ent = np.stack(ents, 2).reshape(W, 32)
        e = np.where(valid, ent, (np.int64(1) << 40) + idx)
        o = np.argsort(e, axis=1, kind="stable")
        es, vs = np.take_along_axis(e, o, 1), np.take_along_axis(valid, o, 1)
        newa = np.ones((W, 32), bool)
        newa[:, 1:] = es[:, 1:] != es[:, :-1]
        rank = idx - np.maximum.accumulate(np.where(newa, idx, 0), axis=1)
        for r in range(int(rank[vs].max()) + 1):
            m = (rank == r) & vs
            if not m.any():
                break
            ss = np.sort(np.where(m, es >> 2, -1 - idx), axis=1)
            total += ((np.diff(ss, axis=1) != 0).sum(1) + 1 - (~m).sum(1)).sum()
    return total


def tile_model(c, R, dense):
    """per CTA: corner bounding-box size (entries) and flush packets (distinct sectors among the touched entries)."""
    G = N // CTA
    cc = c.reshape(G, CTA, 3)
    lo, hi = cc.min(1), cc.max(1) + 1
    ext = hi - lo + 1
    size = ext.prod(1)
    # touched entries: 8 corners per point -> global entry -> sector; distinct per CTA
    secs = []
    for dx in range(2):
        for dy in range(2):
            for dz in range(2):
                if dense:
                    e = ((c[:, 0] + dx) * R + (c[:, 1] + dy)) * R + c[:, 2] + dz
                else:
                    hyz = (((c[:, 1] + dy) * 2654435761) ^ ((c[:, 2] + dz) * 805459861)) & 0xFFFFFFFF
                    e = ((c[:, 0] + dx) ^ hyz) & (T - 1)
                secs.append((e >> 2).reshape(G, CTA))
    s = np.sort(np.concatenate(secs, 1), axis=1)
    pk = (np.diff(s, axis=1) != 0).sum(1) + 1
    return size, pk


print(f"# {N} uniform points, 16-level NGP LoTD (T = 2^19, F = 2); CTA = {CTA} consecutive sorted points, tile budget {CAP} entries (8 bytes each)")
for name in (sys.argv[3].split(",") if len(sys.argv) > 3 else ["xfast", "brick8x4x2", "brick8x4x4", "brick8x2x2", "brick16x2x2", "brick4x4x4", "morton"]):
    xs = x[np.argsort(order_key(name), kind="stable")]
    rows, tot_warp, tot_mix = [], 0.0, 0.0
    used = np.zeros(N // CTA, np.int64)
    for li, R in enumerate(res_list):
        dense = R ** 3 <= T
        c = np.floor(xs * np.float32(R - 2) + np.float32(0.5)).astype(np.int64)
        wp = warp_packets(c, R, dense) / N
        size, pk = tile_model(c, R, dense)
        # greedy in level order: tile while the budget lasts AND the flush is cheaper than the warp path would be on average
        fits = (used + size <= CAP)
        worth = pk / CTA < wp
        take = fits & worth
        used += np.where(take, size, 0)
        mixed = (np.where(take, pk, 0).sum() + (~take).mean() * wp * N) / N
        rows.append((li, R, "Dense" if dense else "Hash ", wp, pk.mean() / CTA, size.mean(), take.mean(), mixed))
        tot_warp += wp
        tot_mix += mixed
    print(f"## order {name}")
    print("# level   res  type | warp-merge pkts/pt | tile flush pkts/pt | mean bbox entries | CTAs tiled | resulting pkts/pt")
    for r in rows:
        print(f"  L{r[0]:<2d}  {r[1]:5d}  {r[2]} | {r[3]:18.3f} | {r[4]:18.3f} | {r[5]:17.0f} | {r[6]:10.2f} | {r[7]:17.3f}")
    print(f"  total               | {tot_warp:18.3f} |                    |                   |            | {tot_mix:17.3f}   (tile memory used: mean {used.mean():.0f}, max {used.max()} entries)")
