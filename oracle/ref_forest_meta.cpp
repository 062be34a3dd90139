// Test infrastructure only: registers the reference's `ForestMeta` struct (csrc/forest/forest_cpp_api.h:16-36, included from
// where it lies under /root/reference) with pybind11 so that the reference's own compiled `forest_ray_marching`
// (oracle/_ref/_occ_grid.so) can be called from tests/golden/make_golden_forest.py.  The reference registers the class in
// csrc/forest/forest.cpp, which cannot be built here (it includes a kaolin header that is not in the tree); this file binds
// the same ten fields and nothing else.  Written for this repository; not reference code.
#include <torch/extension.h>
#include "forest_cpp_api.h"

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    py::class_<ForestMeta>(m, "ForestMeta")
        .def(py::init<>())
        .def_readwrite("octree", &ForestMeta::octree)
        .def_readwrite("exsum", &ForestMeta::exsum)
        .def_readwrite("block_ks", &ForestMeta::block_ks)
        .def_readwrite("world_block_size", &ForestMeta::world_block_size)
        .def_readwrite("world_origin", &ForestMeta::world_origin)
        .def_readwrite("resolution", &ForestMeta::resolution)
        .def_readwrite("n_trees", &ForestMeta::n_trees)
        .def_readwrite("level", &ForestMeta::level)
        .def_readwrite("level_poffset", &ForestMeta::level_poffset)
        .def_readwrite("continuity_enabled", &ForestMeta::continuity_enabled);
}
