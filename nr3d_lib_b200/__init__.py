"""nr3d_lib_b200 -- B200-native (sm_100a) LoTD encoder, occupancy-grid ray marcher and pack_ops.

Only the hot path of PJLab-ADG/nr3d_lib lives here:

* ``nr3d_lib_b200.bindings._lotd / _pack_ops / _occ_grid``  drop-in replacements of the reference's three pybind
  extensions, implemented on the C-ABI library ``lib/libnr3d_b200.so`` (``include/nr3d_b200.h``);
* ``nr3d_lib_b200.lotd / pack_ops / occgrid_raymarch``       host-side mirrors of the reference's autograd wrappers;
* ``nr3d_lib_b200.install.install()``                        registers the bindings as ``nr3d_lib.bindings.*`` so that
  an unmodified nr3d_lib checkout runs on top of them;
* ``nr3d_lib_b200.dist``                                     ray / point sharding + the one NCCL all-reduce of dL/dparams.

There is no CPU path: tensors must live on a CUDA device and the shared library must have been built.
"""
__version__ = "0.1.0"
