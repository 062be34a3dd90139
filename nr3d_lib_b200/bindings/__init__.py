"""Replacements for ``nr3d_lib.bindings.{_lotd,_pack_ops,_occ_grid}`` (see ``nr3d_lib_b200.install``)."""
from . import _lotd, _pack_ops, _occ_grid  # noqa: F401
