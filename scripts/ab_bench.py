"""A/B the fast-path tunables: builds lib variants here (no GPU), runs bench.py per variant on the GPU box.
    python scripts/ab_bench.py build      # on the build box
    python scripts/ab_bench.py run        # on the GPU box (via gpurun)"""
import json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VARIANTS = {
    "base": [],                                   # pair layout (two lanes per point)
    "unroll1": ["NR3D_FWD_UNROLL=1"],
    "unroll4": ["NR3D_FWD_UNROLL=4"],
    "fwd128": ["NR3D_FWD_THREADS=128"],
    "bwd256": ["NR3D_BWD_THREADS=256"],
    "bwd64": ["NR3D_BWD_THREADS=64"],
}
if sys.argv[1] == "build":
    from nr3d_lib_b200.csrc import build as B
    B.build()
    for n, d in VARIANTS.items():
        print(n, B.build_variant(n, d))
else:
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for n in VARIANTS:
        env = dict(os.environ, NR3D_B200_LIB=os.path.join(root, "nr3d_lib_b200", "lib", "variants", n + ".so"))
        out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline", "--no-m2"],
                             env=env, capture_output=True, text=True).stdout.strip().splitlines()
        try:
            d = json.loads(out[-1])
            print(f"{n:10s} {d['value']:8.1f} Msamples/s  step {d['ms_per_step']:.3f} ms  fwd+sort {d['roofline']['ms']['lod_fwd']:.3f}  bwd {d['roofline']['ms']['lod_bwd']:.3f}", flush=True)
        except Exception as e:
            print(n, "FAILED", e, out[-3:])
