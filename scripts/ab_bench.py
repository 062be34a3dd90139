"""A/B the fast-path tunables: builds lib variants here (no GPU), runs bench.py per variant on the GPU box.
    python scripts/ab_bench.py build      # on the build box
    python scripts/ab_bench.py run        # on the GPU box (via gpurun)"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VARIANTS = {
    "base": [],                                   # x-fastest bins, no CTA tiles, 256-thread backward CTAs compiled for 1536 threads / SM
    "bwd128": ["NR3D_BWD_THREADS=128"],
    "bwd256_nocap": ["NR3D_BWD_OCC=0"],
    "bwd128_nocap": ["NR3D_BWD_THREADS=128", "NR3D_BWD_OCC=0"],
    "bwd128_occ2048": ["NR3D_BWD_THREADS=128", "NR3D_BWD_OCC=2048"],
    "fwd128": ["NR3D_FWD_THREADS=128"],
    # round-2 experiments that lost (profiles/r2_ab_tiles.txt): CTA tiles through shared-memory CAS, brick order of the sort bins
    "tiles_brick": ["NR3D_BWD_TILES=1", "NR3D_BIN_ORDER=2", "NR3D_BWD_OCC=0"],
    "brick": ["NR3D_BIN_ORDER=2"],
    # round-2 A/B of TMA staging for the Dense levels of the forward (profiles/r2_ab_fwd_tma.txt)
    "fwd_occ1536": ["NR3D_FWD_OCC=1536"],
    "fwd_occ1280": ["NR3D_FWD_OCC=1280"],
    "fwd_occ1536_u2": ["NR3D_FWD_OCC=1536", "NR3D_FWD_UNROLL=2"],
    "fwd_tma": ["NR3D_FWD_TMA=1"],
    "fwd_tma_12k": ["NR3D_FWD_TMA=1", "NR3D_FWD_TMA_FLOATS=3072"],
    "fwd_tma_brick": ["NR3D_FWD_TMA=1", "NR3D_BIN_ORDER=2"],
    # round-2 (late) A/B of the warp-level merge in the backward (profiles/r2_ab_merge.txt): contiguous runs (round 1) vs any lanes of the warp
    # (match.any / shuffle links + pointer jumping), Hash-level neighbour hand-over, and up to which level merging is attempted
    "runs_all": ["NR3D_BWD_MERGE=0", "NR3D_MERGE_DENSITY=0"],          # the round-1 / early round-2 kernel
    "any_d1": ["NR3D_BWD_MERGE=1", "NR3D_BWD_NEIGH=0", "NR3D_MERGE_DENSITY=1"],
    "any_d2": ["NR3D_BWD_MERGE=1", "NR3D_BWD_NEIGH=0", "NR3D_MERGE_DENSITY=2"],
    "any_d4": ["NR3D_BWD_MERGE=1", "NR3D_BWD_NEIGH=0", "NR3D_MERGE_DENSITY=4"],
    "any_d8": ["NR3D_BWD_MERGE=1", "NR3D_BWD_NEIGH=0", "NR3D_MERGE_DENSITY=8"],
    "any_d16": ["NR3D_BWD_MERGE=1", "NR3D_BWD_NEIGH=0", "NR3D_MERGE_DENSITY=16"],
    "any_neigh_d4": ["NR3D_BWD_MERGE=1", "NR3D_BWD_NEIGH=1", "NR3D_MERGE_DENSITY=4"],
    "link3_d8": ["NR3D_BWD_MERGE=2", "NR3D_BWD_NEIGH=0", "NR3D_MERGE_DENSITY=8"],
}
if sys.argv[1] == "build":
    from nr3d_lib_b200.csrc import build as B
    B.build()
    for n, d in VARIANTS.items():
        if len(sys.argv) > 2 and n not in sys.argv[2].split(","):
            continue
        print(n, B.build_variant(n, d))
else:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    names = sys.argv[2].split(",") if len(sys.argv) > 2 else list(VARIANTS)
    for n in names:
        env = dict(os.environ, NR3D_B200_LIB=os.path.join(root, "nr3d_lib_b200", "lib", "variants", n + ".so"))
        for extra in ([], ["--half"]) if not os.environ.get("NR3D_AB_FP32_ONLY") else ([],):
            r = subprocess.run([sys.executable, os.path.join(root, "scripts", "step_probe.py"), "--tag", n] + extra, env=env, capture_output=True, text=True)
            line = r.stdout.strip().splitlines()[-1] if r.stdout.strip() else ("FAILED " + r.stderr[-400:])
            print(line, flush=True)
