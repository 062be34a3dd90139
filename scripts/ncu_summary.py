#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --page raw --csv) and launch lists into small text files for profiles/."""
import collections, csv, subprocess, sys

KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__block_size", "launch__grid_size",
        "smsp__inst_executed.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_op_red.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]


def rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        print(f"--- {path}")
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"{k:72s} {vals[i]} {units[i]}")


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        try:
            t = float(d["Metric Value"])
        except ValueError:
            continue
        agg[d["Kernel Name"][:90]][0] += 1
        agg[d["Kernel Name"][:90]][1] += t
    tot = sum(v[1] for v in agg.values())
    print(f"--- {path}: launch list (gpu__time_duration.sum, cold-cache serialised: compare SHARES)")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print(f"{v[1] / 1e6:10.3f} ms {v[0]:4d}x {100 * v[1] / tot:5.1f}%  {k}")


for p in sys.argv[1:]:
    (rep if p.endswith(".ncu-rep") else launches)(p)
