"""How many 32-byte reduction sectors does the sorted pair layout HAVE to touch?  (CPU-only, numpy.)

Re-enacts the headline backward kernel's addressing on 4 Mi uniform points: counting sort by a 128^3 bin key (x fastest), 16 points per
warp, lane pair = the two x-neighbour corners of a Hash level (z-neighbours of a Dense level), one reduction instruction per corner group
q = 0..3.  Reports distinct sectors per point for every level if the hardware merged lanes over the whole warp / half / quarter warp,
and what a software pre-reduction over 16 / 64 / 256 / 1024 consecutive points could reach.  Compare with the measured
l1tex__t_sectors_pipe_lsu_mem_global_op_red / N of profiles/r1_s2_pair_ncu_summary.txt (62 sectors per point).

    python scripts/sim_sectors.py > profiles/r1_s2_sector_simulation.txt
"""
import numpy as np

N = 1 << 22
T = 1 << 19
rs = np.random.RandomState(0)
x = rs.rand(N, 3).astype(np.float32)
res_list = (16 * 1.382 ** np.arange(16)).astype(int).tolist()
b = np.minimum((x * 128).astype(np.int64), 127)
xs = x[np.argsort((b[:, 2] * 128 + b[:, 1]) * 128 + b[:, 0], kind="stable")]


def level_sectors(R):
    """[N, 8] sector ids (8-byte entries, 4 per sector) in instruction order q * 2 + side."""
    dense = R ** 3 <= T
    c = np.floor(xs * np.float32(R - 2) + np.float32(0.5)).astype(np.int64)
    out = []
    for q in range(4):
        for side in range(2):
            if dense:
                dx, dy = q & 1, q >> 1
                e = ((c[:, 0] + dx) * R + (c[:, 1] + dy)) * R + c[:, 2] + side
            else:
                dy, dz = q & 1, q >> 1
                hyz = (((c[:, 1] + dy) * 2654435761) ^ ((c[:, 2] + dz) * 805459861)) & 0xFFFFFFFF
                e = ((c[:, 0] + side) ^ hyz) & (T - 1)
            out.append(e >> 2)
    return np.stack(out, 1)


def distinct(a):
    a = np.sort(a, axis=1)
    return (np.diff(a, axis=1) != 0).sum() + a.shape[0]


print(f"# {N} uniform points, 16-level NGP LoTD (T = 2^19, F = 2), sorted by 128^3 bins (x fastest); distinct 32-byte sectors per point")
print("# level  res  type | per instruction, lanes merged over: warp(32)  half(16)  quarter(8) | software pre-reduction over: 16 pts   64    256   1024")
tot = np.zeros(7)
for li, R in enumerate(res_list):
    S = level_sectors(R)
    row = []
    for lanes in (32, 16, 8):
        pts = lanes // 2
        row.append(sum(distinct(S[:, 2 * q:2 * q + 2].reshape(N // pts, lanes)) for q in range(4)) / N)
    for P in (16, 64, 256, 1024):
        row.append(distinct(S.reshape(N // P, P * 8)) / N)
    tot += np.array(row)
    print(f"  L{li:<2d} {R:5d}  {'Dense' if R ** 3 <= T else 'Hash '} | " + "  ".join(f"{v:7.3f}" for v in row[:3]) + "   | " + "  ".join(f"{v:6.3f}" for v in row[3:]))
print("  total              | " + "  ".join(f"{v:7.3f}" for v in tot[:3]) + "   | " + "  ".join(f"{v:6.3f}" for v in tot[3:]))
print("# measured on B200 (ncu, shipped kernel): 260.8 M sectors / 4 194 304 points = 62.2 per point at 212 G sectors/s (L2 reduction unit: 231 G/s)")
