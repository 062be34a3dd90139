// LoTD kernel instantiations for n_dims_to_encode = 3 (split per dimension to compile in parallel,
// as the reference does with compile_split_{1,2,3}.cu).
#include "lotd_kernels.cuh"
namespace nr3d {
NR3D_LOTD_DEFINE_DIM(3)
}
