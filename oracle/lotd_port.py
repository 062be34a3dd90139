"""ctypes front-end of oracle/lotd_port.c (plain-C, fp32, OpenMP port of the reference's Dense/Hash LoTD kernels).
TEST / BASELINE INFRASTRUCTURE ONLY: imported by tests/ and by bench.py's cpu_baseline / `--impl reference` legs."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "lotd_port.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "liblotd_port.so")
_lib = None


def build(force=False):
    """gcc -O3 -fopenmp, baseline x86-64 ISA (the .so may be built on one machine and run on another)."""
    os.makedirs(_OUT_DIR, exist_ok=True)
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O3", "-fopenmp", "-fno-fast-math", "-shared", "-fPIC", "-o", _SO, _SRC, "-lm"])
    return _SO


def _get():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.lotd_port_max_threads.restype = ctypes.c_int
    return _lib


def max_threads():
    return int(_get().lotd_port_max_threads())


def _meta_arrays(meta):
    """(n_levels, res, types, n_feats, sizes, offsets, smooth) from an OracleMeta / LoDMeta with cubic Dense / Hash levels."""
    res = []
    for r in meta.level_res_multidim:
        r = list(r)[:3]
        if len(set(r)) != 1:
            raise ValueError("lotd_port: cubic levels only")
        res.append(r[0])
    types = [int(t) for t in meta.level_types]
    if any(t not in (0, 7) for t in types) or meta.n_dims_to_encode != 3:
        raise ValueError("lotd_port: Dense / Hash levels with D = 3 only")
    u32 = lambda a: np.ascontiguousarray(a, dtype=np.uint32)
    return (len(res), u32(res), u32(types), u32(list(meta.level_n_feats)), u32(list(meta.level_sizes)), u32(list(meta.level_offsets)[:len(res)]),
            int(int(meta.interpolation_type) == 1))


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def fwd(meta, x, params, n_threads=0):
    """y [N, n_enc] float32."""
    L, res, types, nf, sizes, offs, smooth = _meta_arrays(meta)
    x = np.ascontiguousarray(x, dtype=np.float32)
    params = np.ascontiguousarray(params, dtype=np.float32)
    y = np.empty((x.shape[0], int(meta.n_encoded_dims)), dtype=np.float32)
    _get().lotd_port_fwd(ctypes.c_int(L), _p(res), _p(types), _p(nf), _p(sizes), _p(offs), ctypes.c_int(smooth), ctypes.c_uint64(x.shape[0]),
                         _p(x), _p(params), _p(y), ctypes.c_int(int(n_threads)))
    return y


def bwd_param(meta, dL_dy, x, n_threads=0, out=None):
    """dL/dparam [n_params] float32 (accumulated into `out` if given)."""
    L, res, types, nf, sizes, offs, smooth = _meta_arrays(meta)
    x = np.ascontiguousarray(x, dtype=np.float32)
    dL_dy = np.ascontiguousarray(dL_dy, dtype=np.float32)
    grad = np.zeros(int(meta.n_params), dtype=np.float32) if out is None else out
    _get().lotd_port_bwd(ctypes.c_int(L), _p(res), _p(types), _p(nf), _p(sizes), _p(offs), ctypes.c_int(smooth), ctypes.c_uint64(x.shape[0]),
                         _p(x), _p(dL_dy), _p(grad), ctypes.c_int(int(n_threads)))
    return grad
