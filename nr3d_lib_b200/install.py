"""Make an unmodified nr3d_lib checkout use the B200 kernels.

The reference imports its native code by module name (``import nr3d_lib.bindings._lotd as _backend``,
nr3d_lib/models/grid_encodings/lotd/lotd.py:29; pack_ops.py:14; occgrid_raymarch.py:18).  ``install()`` registers this
package's shims under those names in ``sys.modules`` -- call it before importing the reference's wrappers.
"""
import importlib
import sys
import types


def install(force: bool = True) -> None:
    from .bindings import _lotd, _pack_ops, _occ_grid
    try:
        pkg = importlib.import_module("nr3d_lib")
    except Exception:  # nr3d_lib not importable as a whole (missing optional deps): provide a namespace shell
        pkg = sys.modules.get("nr3d_lib")
        if pkg is None:
            pkg = types.ModuleType("nr3d_lib")
            pkg.__path__ = []  # type: ignore[attr-defined]
            sys.modules["nr3d_lib"] = pkg
    bind = sys.modules.get("nr3d_lib.bindings")
    if bind is None:
        bind = types.ModuleType("nr3d_lib.bindings")
        bind.__path__ = []  # type: ignore[attr-defined]
        sys.modules["nr3d_lib.bindings"] = bind
        setattr(pkg, "bindings", bind)
    for name, mod in (("_lotd", _lotd), ("_pack_ops", _pack_ops), ("_occ_grid", _occ_grid)):
        full = "nr3d_lib.bindings." + name
        if force or full not in sys.modules:
            sys.modules[full] = mod
            setattr(bind, name, mod)
