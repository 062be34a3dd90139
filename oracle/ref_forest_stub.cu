// Test infrastructure only: stub for the reference's out-of-scope forest marcher so that the reference's
// `_occ_grid` pybind module (csrc/occ_grid/src/occ_grid.cpp:32) links when forest_marching.cu cannot be
// compiled without the kaolin/forest dependencies.  Written for this repository; not reference code.
#include <occ_grid/cpp_api.h>
#include <stdexcept>

std::vector<at::Tensor> forest_ray_marching(
    const ForestMeta&, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
    const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor, const at::Tensor,
    const float, const float, const float, const uint32_t, const bool) {
    throw std::runtime_error("forest_ray_marching: not built in oracle/_ref (out of scope)");
}
