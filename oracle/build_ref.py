#!/usr/bin/env python
"""Build the UNMODIFIED reference CUDA extensions (csrc/lotd, csrc/pack_ops, csrc/occ_grid) for sm_100
into ``oracle/_ref/`` so that GPU tests / bench can compare against "the reference's own CUDA path".

TEST INFRASTRUCTURE ONLY.  Nothing under ``nr3d_lib_b200/`` may import what this script produces.

* Sources are compiled **where they lie** under /root/reference (read-only); no reference source is
  copied into this repository.  The single file that needs a torch-2.11 compatibility edit
  (``csrc/pack_ops/pack_ops_cuda.cu``: ``pack_ids.type()`` -> ``.scalar_type()`` and the removed
  ``AT_DISPATCH_ALL_TYPES_AND_HALF`` macro) is patched on a scratch copy under /tmp.
* Flags follow the reference's setup.py (setup.py:85-142 lotd, :144-190 pack_ops, :478-522 occ_grid),
  with the arch fixed to sm_100 because there is no GPU in the build container.
* Outputs: ``oracle/_ref/_lotd.so``, ``_pack_ops.so``, ``_occ_grid.so`` (git-ignored, shipped by gpurun).

Usage:  python oracle/build_ref.py [--jobs N] [--only lotd,pack_ops,occ_grid]
"""
import argparse
import os
import shutil
import subprocess
import sys
import sysconfig
import tempfile
import time
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NR3D_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")
CSRC = os.path.join(REF, "csrc")

ARCH = ["-gencode=arch=compute_100,code=sm_100"]


def torch_flags():
    import torch
    from torch.utils import cpp_extension as ce
    inc = []
    for p in ce.include_paths(device_type="cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(cuda=True):
        inc += ["-I" + p]
    inc += ["-I" + sysconfig.get_paths()["include"]]
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    defs = ["-DTORCH_API_INCLUDE_EXTENSION_H", f"-D_GLIBCXX_USE_CXX11_ABI={abi}",
            '-DPYBIND11_COMPILER_TYPE="_gcc"', '-DPYBIND11_STDLIB="_libstdcpp"', '-DPYBIND11_BUILD_ABI="_cxxabi1018"']
    # read the ABI tag torch was actually built with if available
    try:
        defs[-1] = '-DPYBIND11_BUILD_ABI="%s"' % torch._C._PYBIND11_BUILD_ABI
        defs[-2] = '-DPYBIND11_STDLIB="%s"' % torch._C._PYBIND11_STDLIB
        defs[-3] = '-DPYBIND11_COMPILER_TYPE="%s"' % torch._C._PYBIND11_COMPILER_TYPE
    except AttributeError:
        pass
    link = ["-L" + libdir, "-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda",
            "-L/usr/local/cuda/lib64", "-lcudart", "-Wl,-rpath," + libdir]
    return inc, defs, link


POSIX_NVCC = ["-Xcompiler=-mf16c", "-Xcompiler=-Wno-float-conversion", "-Xcompiler=-fno-strict-aliasing",
              "-Xcudafe=--diag_suppress=unrecognized_gcc_pragma", "-Xcompiler=-fPIC", "-w"]


def ext_specs(scratch):
    """name -> dict(sources, includes, nvcc_flags, cxx_flags)."""
    lotd_nv = ["-std=c++17", "--extended-lambda", "--expt-relaxed-constexpr",
               "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF2_OPERATORS__",
               "-U__CUDA_NO_HALF_CONVERSIONS__", "-U__CUDA_NO_BFLOAT16_CONVERSIONS__",
               "-O3", "-DNDEBUG"] + ARCH + POSIX_NVCC
    # torch's CUDAExtension always appends --expt-relaxed-constexpr (torch/utils/cpp_extension.py COMMON_NVCC_FLAGS)
    other_nv = ["-O3", "-std=c++17", "--expt-relaxed-constexpr", "-D__CUDA_NO_HALF_OPERATORS__", "-D__CUDA_NO_HALF_CONVERSIONS__",
                "-D__CUDA_NO_HALF2_OPERATORS__", "-D__CUDA_NO_BFLOAT16_CONVERSIONS__"] + ARCH + POSIX_NVCC
    # --- pack_ops compat patch on a scratch copy (never inside the repo)
    pk_src = os.path.join(CSRC, "pack_ops", "pack_ops_cuda.cu")
    pk_dst = os.path.join(scratch, "pack_ops_cuda_patched.cu")
    with open(pk_src) as f:
        txt = f.read()
    txt = txt.replace("AT_DISPATCH_INTEGRAL_TYPES(pack_ids.type()", "AT_DISPATCH_INTEGRAL_TYPES(pack_ids.scalar_type()")
    with open(pk_dst, "w") as f:
        f.write(txt)
    shim = os.path.join(scratch, "compat_shim.h")
    with open(shim, "w") as f:
        f.write("#pragma once\n#include <ATen/Dispatch.h>\n#ifndef AT_DISPATCH_ALL_TYPES_AND_HALF\n"
                "#define AT_DISPATCH_ALL_TYPES_AND_HALF(TYPE, NAME, ...) "
                "AT_DISPATCH_ALL_TYPES_AND(at::ScalarType::Half, TYPE, NAME, __VA_ARGS__)\n#endif\n")
    # forest_marching.cu may not compile against the forest headers without kaolin; provide a stub then.
    return {
        "_lotd": dict(
            sources=[os.path.join(CSRC, "lotd", "src", f) for f in
                     ("compile_split_1.cu", "compile_split_2.cu", "compile_split_3.cu", "lotd_torch_api.cu", "lotd.cpp")],
            includes=[os.path.join(CSRC, "lotd", "include"), os.path.join(CSRC, "forest")],
            nvcc=lotd_nv, cxx=["-std=c++17", "-O3", "-DNDEBUG", "-fPIC", "-w"]),
        "_pack_ops": dict(
            sources=[pk_dst, os.path.join(CSRC, "pack_ops", "pack_ops.cpp")],
            includes=[os.path.join(CSRC, "pack_ops")],
            nvcc=other_nv + ["-include", shim], cxx=["-std=c++17", "-O3", "-fPIC", "-w"]),
        "_occ_grid": dict(
            sources=[os.path.join(CSRC, "occ_grid", "src", f) for f in
                     ("ray_marching.cu", "batched_marching.cu", "forest_marching.cu", "occ_grid.cpp")],
            includes=[os.path.join(CSRC, "occ_grid", "include"), os.path.join(CSRC, "forest")],
            nvcc=other_nv, cxx=["-std=c++17", "-O3", "-fPIC", "-w"]),
        # our own 20-line pybind registration of the reference's ForestMeta struct (the reference's forest.cpp needs kaolin)
        "_forest": dict(
            sources=[os.path.join(HERE, "ref_forest_meta.cpp")],
            includes=[os.path.join(CSRC, "forest")],
            nvcc=other_nv, cxx=["-std=c++17", "-O3", "-fPIC", "-w"]),
    }


# The reference's own PYTHON wrappers of the hot path (autograd Functions, pack helpers, marcher post-processing) and the two pure-torch
# baselines BASELINE.json names.  They are staged VERBATIM into the git-ignored oracle/_ref/pyref/ (like the compiled .so files: out of
# history, shipped to the GPU box by gpurun) so that the GPU tests can run the UNMODIFIED reference files on top of the injected
# nr3d_lib.bindings.* shims (tests/util.py:reference_wrappers) -- /root/reference itself does not exist on the GPU box.
PY_STAGE = ["nr3d_lib/profile.py", "nr3d_lib/fmt.py", "nr3d_lib/distributed.py",
            "nr3d_lib/models/grid_encodings/lotd/lotd.py", "nr3d_lib/models/grid_encodings/lotd/lotd_helpers.py",
            "nr3d_lib/graphics/pack_ops/__init__.py", "nr3d_lib/graphics/pack_ops/pack_ops.py",
            "nr3d_lib/graphics/raymarch/__init__.py", "nr3d_lib/graphics/raymarch/occgrid_raymarch.py",
            "nr3d_lib/graphics/nerf/nerf_utils.py"]


def stage_python():
    dst_root = os.path.join(OUT, "pyref")
    n = 0
    for rel in PY_STAGE:
        src = os.path.join(REF, rel)
        if not os.path.exists(src):
            print(f"[build_ref] stage-python: {src} missing")
            continue
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        n += 1
    print(f"[build_ref] staged {n} reference python files (verbatim) under {dst_root}")
    return 0


def run(cmd, log):
    t0 = time.time()
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + p.stdout)
    return p.returncode, time.time() - t0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--jobs", type=int, default=min(8, os.cpu_count() or 4))
    ap.add_argument("--only", default="lotd,pack_ops,occ_grid,forest")
    ap.add_argument("--variant", default="", help="checker build of _lotd with extra flags for the generic-kernel TUs (compile_split_*.cu): "
                    "'O1' = -Xptxas -O1, 'O0' = -Xptxas -O0, 'G' = -G; written to oracle/_ref/_lotd__<variant>.so (module name stays _lotd)")
    args = ap.parse_args()
    variant_flags = {"": [], "O1": ["-Xptxas", "-O1"], "O0": ["-Xptxas", "-O0"], "G": ["-G"], "O2": ["-Xptxas", "-O2"]}[args.variant]
    ap_stage_only = args.only == "python"
    if args.variant:
        args.only = "lotd"
    if not os.path.isdir(CSRC):
        print(f"[build_ref] {CSRC} not present - nothing to build (GPU box uses the prebuilt oracle/_ref/*.so)")
        return 0
    os.makedirs(OUT, exist_ok=True)
    stage_python()
    if ap_stage_only:
        return 0
    scratch = tempfile.mkdtemp(prefix="nr3d_ref_build_")
    keep = os.path.join(tempfile.gettempdir(), "nr3d_ref_objs")      # objects of the stock build are kept for the variant links
    os.makedirs(keep, exist_ok=True)
    inc, defs, link = torch_flags()
    specs = ext_specs(scratch)
    want = {"_" + s.strip() for s in args.only.split(",") if s.strip()}
    jobs = []
    for name, sp in specs.items():
        if name not in want:
            continue
        for src in sp["sources"]:
            is_generic_tu = args.variant and os.path.basename(src).startswith("compile_split_")
            obj = os.path.join(keep, name + "__" + os.path.basename(src) + (("." + args.variant) if is_generic_tu else "") + ".o")
            incs = ["-I" + d for d in sp["includes"]]
            if os.path.exists(obj) and os.path.getmtime(obj) > os.path.getmtime(src) and not src.startswith(scratch):
                jobs.append((name, src, obj, ["true"]))
                continue
            if src.endswith(".cu"):
                extra = variant_flags if is_generic_tu else []
                cmd = ["nvcc"] + sp["nvcc"] + extra + incs + inc + defs + [f"-DTORCH_EXTENSION_NAME={name}", "-c", src, "-o", obj]
            else:
                cmd = ["g++"] + sp["cxx"] + incs + inc + defs + [f"-DTORCH_EXTENSION_NAME={name}", "-c", src, "-o", obj]
            jobs.append((name, src, obj, cmd))
    print(f"[build_ref] compiling {len(jobs)} translation units with {args.jobs} jobs (scratch={scratch})", flush=True)
    results = {}
    with ThreadPoolExecutor(max_workers=args.jobs) as ex:
        futs = {ex.submit(run, cmd, obj + ".log"): (name, src, obj) for name, src, obj, cmd in jobs}
        for fut, key in futs.items():
            rc, dt = fut.result()
            results[key] = rc
            print(f"[build_ref] {'ok ' if rc == 0 else 'FAIL'} {key[0]:10s} {os.path.basename(key[1]):28s} {dt:6.0f}s", flush=True)
    status = 0
    for name in specs:
        if name not in want:
            continue
        objs, failed = [], []
        for (n, src, obj), rc in results.items():
            if n != name:
                continue
            (objs if rc == 0 else failed).append((src, obj))
        if failed and name == "_occ_grid" and all("forest_marching" in s for s, _ in failed):
            # forest marcher is out of scope; link a stub that raises so the module imports.
            stub = os.path.join(HERE, "ref_forest_stub.cu")
            sobj = os.path.join(scratch, "ref_forest_stub.o")
            sp = specs[name]
            cmd = ["nvcc"] + sp["nvcc"] + ["-I" + d for d in sp["includes"]] + inc + defs + ["-c", stub, "-o", sobj]
            rc, _ = run(cmd, sobj + ".log")
            if rc == 0:
                objs.append((stub, sobj)); failed = []
                print("[build_ref] forest_marching.cu did not compile; linked oracle/ref_forest_stub.cu instead")
        if failed:
            print(f"[build_ref] {name}: NOT linked, failed units: {[os.path.basename(s) for s, _ in failed]} (logs in {scratch})")
            status = 1
            continue
        so = os.path.join(OUT, name + (("__" + args.variant) if args.variant else "") + ".so")
        rc, _ = run(["g++", "-shared", "-o", so] + [o for _, o in objs] + link, so + ".link.log")
        print(f"[build_ref] link {name}: {'ok' if rc == 0 else 'FAIL'} -> {so}")
        status |= rc
    if status == 0:
        shutil.rmtree(scratch, ignore_errors=True)
    return status


if __name__ == "__main__":
    sys.exit(main())
