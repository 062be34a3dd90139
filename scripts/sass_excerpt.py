#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the shipped library (cuobjdump -sass nr3d_lib_b200/lib/libnr3d_b200.so): which kernels carry
REDG (L2 reductions), UTCHMMA / LDTM / UTCBAR (tcgen05 MMA, TMEM loads, MMA barriers), UBLKCP / UTMALDG (bulk / tensor TMA copies),
ATOMS (shared-memory atomics), MATCH (match.any of the backward's warp-level merge).  Output goes to profiles/ next to the ncu summaries.

    python scripts/sass_excerpt.py > profiles/r2_sass_excerpt.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "nr3d_lib_b200", "lib", "libnr3d_b200.so")
MNEMONICS = ("REDG", "ATOMG", "ATOMS", "UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "MATCH", "LDG", "STG", "LDS", "STS", "SHFL")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None or "/*" not in line:
            continue
        for mn in MNEMONICS:
            if re.search(r"\b" + mn + r"\b", line):
                cnt[fn][mn] += 1
    names = subprocess.run(["c++filt"], input="\n".join(cnt), capture_output=True, text=True).stdout.splitlines()
    want = sys.argv[1:] or ["lotd_pair", "lotd_fused", "lotd_dec", "sort_", "alpha_to_vw", "pack_sum", "march_kernel", "march_fill", "march_compact", "lotd_tma"]
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: instruction counts per kernel (static code, all template instances that match {want})")
    rows = []
    for mangled, name in zip(cnt, names):
        if any(w in name for w in want):
            short = re.sub(r"\(.*", "", name)
            rows.append(f"{short[:100]:100s} " + " ".join(f"{k}={v}" for k, v in sorted(cnt[mangled].items())))
    print("\n".join(sorted(rows)))


if __name__ == "__main__":
    main()
