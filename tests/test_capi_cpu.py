"""CPU tests of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol that
include/nr3d_b200.h declares with the declared arity, the Python shims expose the reference's module surface, and the
unmodified reference wrappers import on top of them (when /root/reference is available)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

from tests.util import ROOT

HEADER = os.path.join(ROOT, "include", "nr3d_b200.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    out = {}
    for m in re.finditer(r"\b(?:int|uint64_t|const char\*)\s+(nr3d_\w+)\s*\(([^;{]*)\)\s*;", src):
        name, args = m.group(1), m.group(2).strip()
        n = 0 if args in ("void", "") else len([a for a in args.split(",") if a.strip()])
        out[name] = n
    return out


def test_library_exports_every_declared_symbol():
    from nr3d_lib_b200 import _lib
    lib = _lib.get_lib()
    decl = _declared_functions()
    assert len(decl) >= 25, decl
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/nr3d_b200.h but not exported"
    # ctypes signatures agree with the header's arity
    for name, argtypes in _lib.SIGNATURES.items():
        assert name in decl, f"{name} bound in _lib.py but not declared in the header"
        assert len(argtypes) == decl[name], (name, len(argtypes), decl[name])
    assert set(decl) - set(_lib.SIGNATURES) == {"nr3d_last_error", "nr3d_version", "nr3d_launch_count"}
    assert lib.nr3d_version() >= 100
    nm = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (nr3d_\w+)", nm))
    assert set(decl) <= exported


def test_struct_layout_matches_header():
    from nr3d_lib_b200 import _lib
    # nr3d_lotd_meta: 8 scalars + 32*4 + 4*32 + 33 + 2*256 uint32
    assert ctypes.sizeof(_lib.LotdMetaStruct) == 4 * (8 + 32 * 4 + 4 * 32 + 33 + 2 * 256)
    # nr3d_forest_meta: three device pointers + four uint32 (forest_cpp_api.h:16-36 as device views)
    assert ctypes.sizeof(_lib.ForestMetaStruct) == 3 * ctypes.sizeof(ctypes.c_void_p) + 4 * 4
    assert [f[0] for f in _lib.ForestMetaStruct._fields_] == ["octree", "exsum", "block_ks", "n_trees", "level", "level_poffset", "continuity_enabled"]


def test_errors_are_reported_without_gpu():
    from nr3d_lib_b200 import _lib
    from nr3d_lib_b200.bindings import _lotd
    with pytest.raises(RuntimeError, match="resolutions >= 3"):
        _lotd.LoDMeta(3, [2], [2], ["Dense"])
    meta = _lotd.LoDMeta(3, [8, 16], [2, 2], ["Dense", "Hash"], 1024)
    # product path refuses CPU tensors loudly (no fallback)
    with pytest.raises(RuntimeError, match="CUDA"):
        _lotd.lod_fwd(meta, torch.rand(4, 3), torch.zeros(meta.n_params))
    assert isinstance(_lib.launch_count(), int)


def test_shim_surface_matches_reference_pybind_modules():
    """Names exported by csrc/{lotd/src/lotd.cpp, pack_ops/pack_ops.cpp, occ_grid/src/occ_grid.cpp}."""
    from nr3d_lib_b200.bindings import _lotd, _occ_grid, _pack_ops
    for n in ("lod_fwd lod_bwd lod_bwd_bwd_input lod_get_grid_index LoDType InterpolationType LoDMeta Dense VectorMatrix CP CPfast "
              "NPlaneMul NPlaneSum Hash Linear Smoothstep").split():
        assert hasattr(_lotd, n), n
    assert [int(_lotd.LoDType[k]) for k in ("Dense", "VectorMatrix", "CP", "CPfast", "NPlaneMul", "NPlaneSum", "Hash")] == [0, 1, 3, 4, 5, 6, 7]
    m = _lotd.LoDMeta(3, [8], [2], ["Dense"])
    for a in ("level_res level_res_multidim level_n_params level_n_feats level_types level_sizes level_types_str level_offsets map_levels "
              "map_cnt n_levels n_pseudo_levels n_feat_per_pseudo_lvl n_dims_to_encode n_encoded_dims n_params interpolation_type "
              "c_hash_only c_profile c_bmm_backend c_prefetch c_permute_dydx").split():
        assert hasattr(m, a), a
    m.c_permute_dydx = False
    with pytest.raises(AttributeError):
        m.n_params = 3
    for n in ("interleave_arange interleave_linstep interleave_sample_step_wrt_depth_clamp_deprecated interleave_sample_step_wrt_depth_clamped "
              "interleave_sample_step_wrt_depth_in_packed_segments packed_add packed_sub packed_mul packed_div packed_matmul packed_gt packed_geq "
              "packed_lt packed_leq packed_eq packed_neq packed_sum packed_diff packed_backward_diff packed_cumsum packed_cumprod packed_sort_qsort "
              "packed_sort_thrust packed_searchsorted packed_searchsorted_packed_vals try_merge_two_packs_sorted_aligned packed_invert_cdf "
              "packed_alpha_to_vw_forward packed_alpha_to_vw_backward mark_pack_boundaries_cuda octree_mark_consecutive_segments").split():
        assert callable(getattr(_pack_ops, n)), n
    for n in ("ray_marching", "batched_ray_marching", "forest_ray_marching", "ContractionType"):
        assert hasattr(_occ_grid, n), n
    assert [int(_occ_grid.ContractionType[k]) for k in ("AABB", "UN_BOUNDED_TANH", "UN_BOUNDED_SPHERE")] == [0, 1, 2]
    import pickle
    m2 = pickle.loads(pickle.dumps(_lotd.LoDMeta(3, [8, 9], [2, 4], ["Dense", "VM"])))
    assert m2.n_params == _lotd.LoDMeta(3, [8, 9], [2, 4], ["Dense", "VM"]).n_params


@pytest.mark.skipif(not os.path.isdir("/root/reference/nr3d_lib"), reason="reference checkout not present")
def test_unmodified_reference_wrappers_import_on_the_shims():
    """Drop-in: the reference's own lotd.py / pack_ops.py / occgrid_raymarch.py import against the injected modules."""
    code = r"""
import sys, types
sys.path.insert(0, %r)
from nr3d_lib_b200.install import install
# namespace shells so that only the three boundary files (and fmt/profile) of the reference are imported
import importlib.util, os
REF = "/root/reference/nr3d_lib"
def shell(name, path=None):
    m = types.ModuleType(name); m.__path__ = [path] if path else []; sys.modules[name] = m; return m
shell("nr3d_lib", REF)
# parent packages as shells (their __init__ pull in the whole model zoo and its third-party deps, all out of scope)
shell("nr3d_lib.models", REF + "/models")
shell("nr3d_lib.models.grid_encodings", REF + "/models/grid_encodings")
shell("nr3d_lib.models.grid_encodings.lotd", REF + "/models/grid_encodings/lotd")
shell("nr3d_lib.graphics", REF + "/graphics")
install()
import nr3d_lib.bindings._lotd as b
from nr3d_lib.models.grid_encodings.lotd import lotd as ref_lotd      # unmodified reference file
enc = ref_lotd.LoTD(3, [8, 16, 32], [2, 2, 2], ["Dense", "Hash", "VM"], hashmap_size=1024, dtype=__import__("torch").float)
assert enc.n_params == b.LoDMeta(3, [8, 16, 32], [2, 2, 2], ["Dense", "Hash", "VM"], 1024).n_params
assert [t.name for t in enc.level_types] == ["Dense", "Hash", "VectorMatrix"]
from nr3d_lib.graphics.pack_ops import pack_ops as ref_pack              # unmodified reference file
assert ref_pack._backend is sys.modules["nr3d_lib.bindings._pack_ops"]
from nr3d_lib.graphics.raymarch import occgrid_raymarch as ref_march   # unmodified reference file
assert ref_march._backend is sys.modules["nr3d_lib.bindings._occ_grid"] and ref_march.ContractionType.AABB.value == 0
print("DROPIN_OK")
""" % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    if "DROPIN_OK" not in r.stdout:
        # the reference package __init__ chain may need optional deps that are absent here: report, do not hide
        missing = re.findall(r"No module named '([\w\.]+)'", r.stderr)
        if missing and all(not m.startswith("nr3d_lib_b200") for m in missing):
            pytest.skip(f"reference import chain needs unavailable third-party modules: {sorted(set(missing))}")
        raise AssertionError(r.stderr[-2000:])


def test_sort_points_flag_default_and_env(monkeypatch):
    """`LoDMeta.c_sort_points` (B200-only knob) is ON by default -- a drop-in user gets the fast path -- and NR3D_B200_SORT_POINTS=0 (or
    `meta.c_sort_points = False`) restores the reference's strides / the generic kernels."""
    from nr3d_lib_b200.bindings import _lotd
    args = (3, [16, 32], [2, 2], ["Dense", "Hash"], 2 ** 10, False)
    monkeypatch.delenv("NR3D_B200_SORT_POINTS", raising=False)
    assert _lotd.LoDMeta(*args).c_sort_points is True
    monkeypatch.setenv("NR3D_B200_SORT_POINTS", "0")
    m = _lotd.LoDMeta(*args)
    assert m.c_sort_points is False
    m.c_sort_points = 1
    assert m.c_sort_points is True
    monkeypatch.setenv("NR3D_B200_SORT_POINTS", "1")
    assert _lotd.LoDMeta(*args).c_sort_points is True


def test_sort_workspace_sizes_and_scratch_validation_without_gpu():
    """Host-side logic of the two scratch buffers added late in round 2, exercised through the C-ABI without a GPU (argument checks precede every
    CUDA call): the sort workspace layout (size query; two-level scratch only for large single-scene calls without batch indices) and the
    single-pass marcher's record buffer."""
    from nr3d_lib_b200 import _lib
    lib = _lib.get_lib()

    def need(N, has_inds=False, bds=0, ns=1):
        nb = ctypes.c_uint64(0)
        fake = ctypes.c_void_p(256) if has_inds else None          # only tested for NULL-ness by the size query
        _lib.check(lib.nr3d_lotd_sort_points(N, None, fake, bds, ns, 1, None, None, None, ctypes.byref(nb), None))
        return nb.value

    n4, n30 = need(4 << 20), need(30 << 20)
    head = 256 + 8192 * 8
    cnt = -(-(128 ** 3 + 1) * 4 // 256) * 256
    assert n4 == head + 2 * cnt + (4 << 20) * 4                            # header + scan status + 2 counter tables (128^3 bins + 1) + rank
    assert n30 - ((30 << 20) * 16) > head + (30 << 20) * 4                 # two-level: + one 16-byte record per point (+ 256^3 counters)
    assert need(30 << 20, has_inds=True) < n30 - (30 << 20) * 16 + 4096    # batch indices -> one-level sort, no record scratch
    assert need(30 << 20, bds=1 << 20, ns=30) < n30                        # batched by data size: one-level as well
    assert need(8 << 20) > n4 and need(1000) < n4                          # monotone in the point count
    # ws_reset refuses a workspace that is too small for the configuration (before touching the device)
    rc = lib.nr3d_lotd_sort_ws_reset(4 << 20, 0, 0, 1, ctypes.c_void_p(256), 1024, None)
    assert rc != 0 and b"workspace too small" in lib.nr3d_last_error()
    rc = lib.nr3d_lotd_sort_ws_reset(4 << 20, 0, 0, 1, ctypes.c_void_p(128), n4, None)
    assert rc != 0 and b"256-byte aligned" in lib.nr3d_last_error()
    # marcher record pass: scratch smaller than n_rays * max_steps * 16 bytes is an error, so is a misaligned one
    p = ctypes.c_void_p(4096)
    common = (1000, p, p, p, p, None, 0, 1, p, p, 8, 8, 8, 0, 0.01, 1e10, 0.0, 64)
    rc = lib.nr3d_march_record(*common, p, p, 1000 * 64 * 16 - 16, None)
    assert rc != 0 and b"records buffer too small" in lib.nr3d_last_error()
    rc = lib.nr3d_march_record(*common, p, ctypes.c_void_p(4104), 1000 * 64 * 16, None)
    assert rc != 0 and b"16-byte aligned" in lib.nr3d_last_error()
    assert lib.nr3d_march_compact(0, None, 0, None, None, None, None, None, None, None, None) == 0      # nothing to do for zero rays
