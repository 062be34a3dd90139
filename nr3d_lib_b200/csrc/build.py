#!/usr/bin/env python
"""Build libnr3d_b200.so (sm_100a only) in-tree with plain nvcc.  No torch headers are involved: the library is a
pure C-ABI / CUDA-runtime shared object (see include/nr3d_b200.h)."""
import hashlib
import os
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
LIB_DIR = os.path.join(PKG, "lib")
OBJ_DIR = os.path.join(HERE, "build")
LIB = os.path.join(LIB_DIR, "libnr3d_b200.so")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--use_fast_math=false",
         "-Xcompiler", "-fPIC", "-DNR3D_BUILDING", "-cudart", "static"]
FLAGS = [f for f in FLAGS if f != "--use_fast_math=false"]

SOURCES = ["lotd_api.cu", "lotd_d2.cu", "lotd_d3.cu", "lotd_d4.cu", "lotd_fast.cu", "lotd_sort.cu", "lotd_fused.cu", "lotd_fused_bwd.cu", "lotd_forest.cu", "pack_ops.cu", "pack_staged.cu", "pack_next.cu", "march.cu", "occ_update.cu", "pipeline_ops.cu"]


def _deps():
    return [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith((".cuh", ".h"))] + \
           [os.path.join(os.path.dirname(PKG), "include", "nr3d_b200.h")]


def _stamp(src):
    h = hashlib.sha1()
    for p in [src] + sorted(_deps()):
        with open(p, "rb") as f:
            h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OBJ_DIR, os.path.basename(src) + ".o")
    stamp_file = obj + ".stamp"
    stamp = _stamp(src)
    if os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return obj, 0.0, ""
    t0 = time.time()
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if p.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}:\n{p.stdout}")
    with open(stamp_file, "w") as f:
        f.write(stamp)
    return obj, time.time() - t0, p.stdout


def build_variant(name, defines):
    """A/B build of the library with extra -D flags into lib/variants/<name>.so (used by scripts/ab_bench.py)."""
    out_dir = os.path.join(LIB_DIR, "variants")
    obj_dir = os.path.join(OBJ_DIR, "variant_" + name)
    os.makedirs(out_dir, exist_ok=True)
    os.makedirs(obj_dir, exist_ok=True)
    objs = []
    for s in SOURCES:
        src = os.path.join(HERE, s)
        if s in ("lotd_fast.cu", "lotd_sort.cu"):
            obj = os.path.join(obj_dir, s + ".o")
            subprocess.check_call([NVCC] + FLAGS + ["-D" + d for d in defines] + ["-c", src, "-o", obj])
        else:
            obj = os.path.join(OBJ_DIR, s + ".o")
        objs.append(obj)
    lib = os.path.join(out_dir, name + ".so")
    subprocess.check_call([NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib] + objs)
    return lib


def build(verbose=False, force=False):
    os.makedirs(LIB_DIR, exist_ok=True)
    os.makedirs(OBJ_DIR, exist_ok=True)
    srcs = [os.path.join(HERE, s) for s in SOURCES if os.path.exists(os.path.join(HERE, s))]
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        res = list(ex.map(lambda s: _compile(s, verbose), srcs))
    objs = [r[0] for r in res]
    rebuilt = any(r[1] > 0 for r in res)
    for (obj, dt, out), s in zip(res, srcs):
        if dt > 0:
            print(f"[nr3d_b200 build] {os.path.basename(s):18s} {dt:6.1f}s")
            if verbose:
                print(out)
    if rebuilt or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a", "-o", LIB] + objs
        p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if p.returncode != 0:
            raise RuntimeError("link failed:\n" + p.stdout)
        print(f"[nr3d_b200 build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force="-f" in sys.argv)
