"""How many L2 reduction packets does the headline backward HAVE to issue?  (CPU-only, numpy.)

Re-enacts the kernel's addressing on 4 Mi uniform points: counting sort by a 128^3 bin key (x fastest), 16 points per warp, lane pair =
the two x-neighbour corners of a Hash level (z-neighbours of a Dense level), one reduction instruction per corner group q = 0..3.
Cost model (scripts/ubench_redgroup.cu, profiles/r1_s2_ubench_redgroup.txt): lanes of one instruction that hit DIFFERENT entries of a
32-byte sector share one packet, lanes that hit the SAME entry need one packet each.  Columns, in packets per point and level:
  no merge          every lane issues its reduction
  runs >= 4 merged  the shared-memory run merge (consecutive points in one cell) for runs of four and more  -- the build that ncu saw
  all runs merged   + runs of two and three points through shuffles                                        -- the build shipped until late round 2
  any lanes merged  every point of the warp in the same cell merged, neighbours or not (match.any + pointer jumping) -- the shipped build on the
                    levels it merges (res^3 <= 4 N); a cell boundary that cuts a sort bin leaves the bin's points interleaved, which "runs" cannot see
  + x neighbours    additionally, on Hash levels, the side-0 head of cell (X, y, z) hands its sums to the side-1 head of cell (X - 1, y, z): this
                    reaches the floor of the Hash levels and still loses on the GPU (a second match.any, profiles/r2_ab_merge.txt)
  distinct sectors  the floor if every repeated entry inside an instruction were merged for free
Measured (ncu, profiles/r1_s2_pair_ncu_summary.txt): 260.8 M packets / 4 194 304 points = 62.2 for the "runs >= 4" build.

    python scripts/sim_sectors.py > profiles/r2_sector_simulation.txt     (round 1: profiles/r1_s2_sector_simulation.txt)
"""
import numpy as np

N = 1 << 22
T = 1 << 19
W = N // 16
rs = np.random.RandomState(0)
x = rs.rand(N, 3).astype(np.float32)
res_list = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
b = np.minimum((x * 128).astype(np.int64), 127)
xs = x[np.argsort((b[:, 2] * 128 + b[:, 1]) * 128 + b[:, 0], kind="stable")]


def packets(ent, valid):
    """ent, valid: [W, 32].  Sum over duplicate ranks of the distinct sectors among the lanes of that rank."""
    L = ent.shape[1]
    idx = np.arange(L)[None, :].repeat(W, 0)
    e = np.where(valid, ent, (np.int64(1) << 40) + idx)
    o = np.argsort(e, axis=1, kind="stable")
    es, vs = np.take_along_axis(e, o, 1), np.take_along_axis(valid, o, 1)
    newa = np.ones((W, L), bool)
    newa[:, 1:] = es[:, 1:] != es[:, :-1]
    rank = idx - np.maximum.accumulate(np.where(newa, idx, 0), axis=1)
    total = 0
    for r in range(int(rank[vs].max()) + 1):
        m = (rank == r) & vs
        if not m.any():
            break
        ss = np.sort(np.where(m, es >> 2, -1 - idx), axis=1)
        total += ((np.diff(ss, axis=1) != 0).sum(1) + 1 - (~m).sum(1)).sum()
    return total


print(f"# {N} uniform points, 16-level NGP LoTD (T = 2^19, F = 2), sorted by 128^3 bins (x fastest); L2 reduction packets per point")
print("# level   res  type | no merge | runs >= 4 merged | all runs merged | any lanes merged | + x neighbours | distinct sectors")
tot = np.zeros(6)
for li, R in enumerate(res_list):
    dense = R ** 3 <= T
    c = np.floor(xs * np.float32(R - 2) + np.float32(0.5)).astype(np.int64)
    key = (c[:, 0] | (c[:, 1] << 10) | (c[:, 2] << 20)).reshape(W, 16)
    head = np.ones((W, 16), bool)
    head[:, 1:] = key[:, 1:] != key[:, :-1]
    run_id = np.cumsum(head, axis=1)
    rl = np.zeros((W, 16), np.int64)
    for r in range(1, 17):
        m = run_id == r
        rl += m * m.sum(1, keepdims=True)
    head_any = np.ones((W, 16), bool)      # first point of its cell in the warp
    for k in range(1, 16):
        head_any[:, k] = ~(key[:, :k] == key[:, k:k + 1]).any(1)
    valid_any = np.repeat(head_any, 2, axis=1)
    valid_nb = valid_any.copy()
    if not dense:
        keym1 = ((c[:, 0] - 1) | (c[:, 1] << 10) | (c[:, 2] << 20)).reshape(W, 16)
        for k in range(16):
            has = ((key == keym1[:, k:k + 1]) & head_any).any(1) & head_any[:, k]
            valid_nb[:, 2 * k] &= ~has
    res = np.zeros(6)
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
        res[0] += packets(ent, np.ones((W, 32), bool))
        res[1] += packets(ent, np.repeat(head | (rl < 4), 2, axis=1))
        res[2] += packets(ent, np.repeat(head, 2, axis=1))
        res[3] += packets(ent, valid_any)
        res[4] += packets(ent, valid_nb)
        s = np.sort(ent >> 2, axis=1)
        res[5] += (np.diff(s, axis=1) != 0).sum() + W
    res /= N
    tot += res
    print(f"  L{li:<2d}  {R:5d}  {'Dense' if dense else 'Hash '} | {res[0]:8.3f} | {res[1]:16.3f} | {res[2]:15.3f} | {res[3]:16.3f} | {res[4]:14.3f} | {res[5]:16.3f}", flush=True)
print(f"  total               | {tot[0]:8.3f} | {tot[1]:16.3f} | {tot[2]:15.3f} | {tot[3]:16.3f} | {tot[4]:14.3f} | {tot[5]:16.3f}")
print("# measured on B200: 62.2 packets per point at 212 G/s for the 'runs >= 4' build (backward 1.268 ms); the 'all runs' build runs the backward")
print("# in 1.148 ms -- the time followed the packet count (61.6 -> 55.6 = -9.7 %, time -9.5 %).  'any lanes' on the levels up to res 406 (shipped,")
print("# profiles/r2_ab_merge.txt): 1.200 -> 1.169 ms; by now the kernel also sits on its issue slots, and every match.any costs about as much as")
print("# it finds distinct values, so the last 3 packets ('+ x neighbours') cost more than they save.")
