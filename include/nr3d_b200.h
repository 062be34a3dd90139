/* nr3d_b200.h -- C ABI of libnr3d_b200.so: the B200-native (sm_100a) replacement for the three
 * native extensions on nr3d_lib's LoTD / occupancy-march / pack_ops hot path.
 *
 * Each entry point names the reference interface it replaces (file:line under the reference tree).
 * Conventions
 *   - plain C types only; every `const void*` / `void*` is a DEVICE pointer unless marked (host);
 *   - the callee never allocates device memory: outputs are caller-allocated (the Python shim
 *     allocates them through torch's caching allocator), two-pass ops expose a count + fill pair;
 *   - every call enqueues on `stream` (a cudaStream_t passed as void*) and returns immediately;
 *   - return value 0 = ok, negative = error; nr3d_last_error() returns a thread-local message.
 *     The Python shim turns a non-zero status into RuntimeError (the reference throws
 *     std::runtime_error / c10::Error, e.g. csrc/lotd/src/lotd_torch_api.cu:41,257).
 *   - thread-safe and re-entrant: no global mutable state besides the thread-local error string.
 */
#ifndef NR3D_B200_H
#define NR3D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NR3D_MAX_LEVELS 32          /* csrc/lotd/include/lotd/lotd_cuda.h:26 */
#define NR3D_MAX_DIMS 4             /* csrc/lotd/include/lotd/lotd_cuda.h:27 */
#define NR3D_MAX_PSEUDO_LEVELS 256  /* csrc/lotd/include/lotd/lotd_cuda.h:37 (MAX_N_LEVELS * 8) */

/* dtype codes used by every entry point */
enum { NR3D_F32 = 0, NR3D_F16 = 1, NR3D_F64 = 2, NR3D_I32 = 3, NR3D_I64 = 4, NR3D_I16 = 5, NR3D_I8 = 6, NR3D_U8 = 7 };

/* LoDType values follow the reference's C++ enum incl. the un-exported VecZMatXoY
 * (csrc/lotd/include/lotd/lotd_types.h:16-26). */
enum { NR3D_LOD_DENSE = 0, NR3D_LOD_VM = 1, NR3D_LOD_VECZMATXOY = 2, NR3D_LOD_CP = 3, NR3D_LOD_CPFAST = 4,
       NR3D_LOD_NPLANEMUL = 5, NR3D_LOD_NPLANESUM = 6, NR3D_LOD_HASH = 7 };
enum { NR3D_INTERP_LINEAR = 0, NR3D_INTERP_SMOOTHSTEP = 1 }; /* lotd_types.h:78-82 */

/* Host-side POD mirror of LoDMeta / LoDMetaRef
 * (csrc/lotd/include/lotd/lotd_torch_api.h:81-135, csrc/lotd/include/lotd/lotd_cuda.h:29-76). */
typedef struct nr3d_lotd_meta {
    uint32_t n_levels, n_pseudo_levels, n_feat_per_pseudo_lvl, n_dims_to_encode, n_encoded_dims, n_params;
    uint32_t interpolation_type; /* NR3D_INTERP_* */
    uint32_t hash_only;          /* 1 when every level is Dense or Hash (c_hash_only) */
    uint32_t level_res[NR3D_MAX_LEVELS][NR3D_MAX_DIMS];
    uint32_t level_n_feats[NR3D_MAX_LEVELS];
    uint32_t level_types[NR3D_MAX_LEVELS];
    uint32_t level_n_params[NR3D_MAX_LEVELS];
    uint32_t level_sizes[NR3D_MAX_LEVELS];
    uint32_t level_offsets[NR3D_MAX_LEVELS + 1];
    uint32_t map_levels[NR3D_MAX_PSEUDO_LEVELS];
    uint32_t map_cnt[NR3D_MAX_PSEUDO_LEVELS];
} nr3d_lotd_meta;

const char* nr3d_last_error(void);
int nr3d_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t nr3d_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * LoTD  (replaces nr3d_lib.bindings._lotd, csrc/lotd/src/lotd.cpp:23-110)
 * ---------------------------------------------------------------------------------------------- */

/* == LoDMeta::create_meta, csrc/lotd/src/lotd_torch_api.cu:29-230.  Host only, no CUDA call.
 * res: [n_levels * n_dims] (host), n_feats/types: [n_levels] (host). hashmap_size 0 = not given. */
int nr3d_lotd_meta_create(int32_t n_dims, int32_t n_levels, const int32_t* res, const int32_t* n_feats,
                          const int32_t* types, uint32_t hashmap_size, int32_t use_smooth_step,
                          nr3d_lotd_meta* out /* host */);

/* == lod_fwd, csrc/lotd/src/lotd_torch_api.cu:232-395 (kernels lotd_hash_only.h:15-378, lotd_encoding.h:113-428).
 * x: [N, D] input_dtype contiguous.  params: [n_batches * n_params] param_dtype.
 * batch_inds: int64 [N] or NULL; batch_offsets: int64 [B] or NULL; batch_data_size: 0 = unused.
 * y element (n, j) is written at y + n*y_stride_n + j*y_stride_f (elements of param_dtype); EVERY element is
 * written (zeros for skipped points / levels) so the caller may pass uninitialised memory.
 * dy_dx (nullable): element (n, j, d) at dy_dx + n*dydx_stride_n + j*dydx_stride_f + d (input_dtype). */
int nr3d_lotd_fwd(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                  const void* x, const void* params, const int64_t* batch_inds, const int64_t* batch_offsets,
                  uint32_t batch_data_size, int32_t max_level,
                  void* y, int64_t y_stride_n, int64_t y_stride_f,
                  void* dy_dx, int64_t dydx_stride_n, int64_t dydx_stride_f, void* stream);

/* == lod_bwd (dL/dparam part), csrc/lotd/src/lotd_torch_api.cu:397-573
 * (kernels lotd_hash_only.h:380-470, lotd_encoding.h:467-711).
 * dL_dy element (n, j) read at dL_dy + n*s_n + j*s_f (param_dtype, any strides).
 * dL_dparam: [n_batches*n_params] param_dtype, MUST be zero-filled by the caller; gradients are accumulated. */
int nr3d_lotd_bwd_param(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                        const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f,
                        const void* x, const void* params, const int64_t* batch_inds, const int64_t* batch_offsets,
                        uint32_t batch_data_size, int32_t max_level, void* dL_dparam, void* stream);

/* == lod_bwd (dL/dx part): dL_dx[n,d] = sum_j dL_dy[n,j] * dy_dx[n,j,d]
 * (reference: at::mul + at::sum_out, lotd_hash_only.h:839-863, lotd_encoding.h:1562-1586), fused in one kernel.
 * dL_dx: [N, D] input_dtype contiguous, fully written. */
int nr3d_lotd_bwd_input(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                        const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f,
                        const void* dy_dx, int64_t dydx_stride_n, int64_t dydx_stride_f,
                        void* dL_dx, void* stream);

/* == lod_bwd_bwd_input, csrc/lotd/src/lotd_torch_api.cu:575-769 -- three independent outputs, each nullable:
 *   dL_ddLdy [N, n_enc] param_dtype contiguous = sum_d dL_ddLdx[n,d]*dy_dx[n,j,d]   (lotd_hash_only.h:982-1006)
 *   dL_dparam (zero-filled by caller) += d(dL/dx)/dparam . dL_ddLdx                (lotd_encoding.h:713-1041)
 *   dL_dx [N, D] (zero-filled by caller) += d(dL/dx)/dx . dL_ddLdx                  (lotd_encoding.h:1043-1298) */
int nr3d_lotd_bwd_bwd_input(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                            const void* dL_ddLdx, const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f,
                            const void* x, const void* params,
                            const void* dy_dx, int64_t dydx_stride_n, int64_t dydx_stride_f,
                            const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size,
                            int32_t max_level, void* dL_ddLdy, void* dL_dparam, void* dL_dx, void* stream);

/* dL/dparam (dL_ddLdx == NULL) or the second-order d(dL/dx)/dparam . dL_ddLdx (dL_ddLdx: f32 [N, D]) with the number of scenes
 * behind `params` made explicit (n_batches = params.numel / n_params; 0 = unknown).  Same results as nr3d_lotd_bwd_param /
 * the dL_dparam output of nr3d_lotd_bwd_bwd_input; knowing the scene count lets small Dense / CP tables and the lines of VM levels
 * be accumulated in shared memory (one copy per scene and CTA) before they are added to dL_dparam -- the reference scatters every
 * corner straight into these same-address hot spots (lotd_cuda.h:494-829).  dL_dparam is accumulated into: zero it first. */
int nr3d_lotd_bwd_param_scenes(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N, const void* dL_dy,
                               int64_t dLdy_stride_n, int64_t dLdy_stride_f, const void* dL_ddLdx, const void* x, const void* params,
                               const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size, uint32_t n_batches,
                               int32_t max_level, void* dL_dparam, void* stream);

/* == lod_get_grid_index, csrc/lotd/src/lotd_torch_api.cu:771-855 (kernel lotd_encoding.h:1300-1433).
 * out: int64 [N, n_enc, 2^D], MUST be zero-filled by the caller (skipped entries stay 0). Dense/Hash only. */
int nr3d_lotd_grid_index(const nr3d_lotd_meta* meta, int32_t input_dtype, uint64_t N, const void* x,
                         const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size,
                         int32_t max_level, int64_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * LoTD over a forest of blocks.  Replaces the `metas=(LoDMeta, ForestMeta)` overloads of
 * lod_fwd / lod_bwd / lod_bwd_bwd_input (csrc/lotd/src/lotd.cpp:45-58 -> lod_forest_*, lotd_torch_api.cu:367-381,539-556,750-769;
 * kernels csrc/lotd/include/lotd/lotd_forest.h).  `batch_inds` are the points' BLOCK indices (int64 [N], < 0 = skip),
 * `x` their block-local coordinates in [0,1]^3; every block owns n_params parameters at batch_offsets[b] (or b * n_params).
 * Corners on a block face belong to the neighbour block found through the octree (csrc/forest/forest.h:25-57); absent
 * neighbours contribute zero.  D = 3; level types Dense / VM / NPlaneMul / CP / Hash.  Outputs are row-major.
 * ---------------------------------------------------------------------------------------------- */
typedef struct nr3d_forest_meta {          /* device views of ForestMeta's tensors, csrc/forest/forest_cpp_api.h:16-36 */
    const uint8_t* octree;                 /* uint8 [n_nodes]   SPC octree bytes (kaolin layout) */
    const int32_t* exsum;                  /* int32 [n_nodes+1] exclusive sum of the child counts */
    const int16_t* block_ks;               /* int16 [n_trees,3] integer block coordinates */
    uint32_t n_trees, level, level_poffset, continuity_enabled;
} nr3d_forest_meta;

/* == lod_forest_fwd.  y: [N, n_enc] param dtype; dy_dx: f32 [N, n_enc, 3] or NULL; both fully written. */
int nr3d_lotd_forest_fwd(const nr3d_lotd_meta* meta, const nr3d_forest_meta* forest, int32_t input_dtype, int32_t param_dtype,
                         uint64_t N, const void* x, const void* params, const int64_t* batch_inds, const int64_t* batch_offsets,
                         uint32_t batch_data_size, int32_t max_level, void* y, void* dy_dx, void* stream);
/* == lod_forest_bwd (dL/dparam) when dL_ddLdx == NULL, else the second-order d(dL/dx)/dparam . dL_ddLdx of
 * lod_forest_bwd_bwd_input.  dL_dparam (param dtype, [n_blocks * n_params]) is ACCUMULATED into: zero it first.
 * (dL/dx and dL/d(dL/dy) are contractions with dy_dx: use nr3d_lotd_bwd_input / nr3d_lotd_bwd_bwd_input with dL_dparam = dL_dx = NULL.) */
int nr3d_lotd_forest_bwd_param(const nr3d_lotd_meta* meta, const nr3d_forest_meta* forest, int32_t input_dtype, int32_t param_dtype,
                               uint64_t N, const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f, const void* dL_ddLdx,
                               const void* x, const void* params, const int64_t* batch_inds, const int64_t* batch_offsets,
                               uint32_t batch_data_size, int32_t max_level, void* dL_dparam, void* stream);
/* == the d(dL/dx)/dx . dL_ddLdx part of lod_forest_bwd_bwd_input (Dense / VM / Hash levels, lotd_forest.h:1022-1050).
 * dL_dx: f32 [N,3], accumulated into (zero it first). */
int nr3d_lotd_forest_bwd_bwd_dx(const nr3d_lotd_meta* meta, const nr3d_forest_meta* forest, int32_t input_dtype, int32_t param_dtype,
                                uint64_t N, const void* dL_ddLdx, const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f,
                                const void* x, const void* params, const int64_t* batch_inds, const int64_t* batch_offsets,
                                uint32_t batch_data_size, int32_t max_level, void* dL_dx, void* stream);

/* B200 fast path for Dense/Hash-only metas (D = 3, 2 / 4 / 8 features per pseudo level, fp32 or fp16 params, one scene or many).
 * No reference counterpart: the shim uses it for lod_fwd / lod_bwd / lod_bwd_bwd_input whenever LoDMeta.c_sort_points is set (default)
 * and the call is eligible; the reference's hash-only kernels (lotd_hash_only.h:15-695) serve the same set of configurations.
 *   sort_points : counting sort of the points by (scene, cell bin) into `xs`, float4 [N] records (x, y, z, original index as raw uint32
 *                 bits) and -- for batched calls -- `scenes`, uint16 [N] (0xffff = skipped point: batch_inds < 0).  batch_inds (int64 [N])
 *                 / batch_data_size / n_scenes as in nr3d_lotd_fwd.  Query the workspace size with ws == NULL.
 *                 The workspace is STATEFUL: zero-fill it before its first use and keep it with `xs`; every call fingerprints the points it
 *                 is given (on the device, in stream order) and re-sorts only when they differ from the ones the records were built from
 *                 (`force` != 0: always sort, the fingerprint is taken inside the histogram pass -- what a forward call with new points wants) --
 *                 lod_fwd and lod_bwd of one step share one sort without trusting host-side tensor identity.
 *   fwd_sorted  : same values as nr3d_lotd_fwd; y element (n, j) at y + n*y_stride_n + j*y_stride_f (row-major is fastest).
 *   bwd_param_sorted : same values as nr3d_lotd_bwd_param (up to fp32 summation order); dL_dparam zero-filled by caller.  Only the pseudo
 *                 levels [pl_begin, pl_end) are scattered (0, UINT32_MAX: all) -- multi-GPU callers launch the fine levels first and start
 *                 the all-reduce of their part of the table while the coarse levels are still being scattered.
 * params / dL_dparam must be 16-byte aligned; n_scenes * n_params < 2^32. */
int nr3d_lotd_sort_points(uint64_t N, const float* x, const int64_t* batch_inds, uint32_t batch_data_size, uint32_t n_scenes, int32_t force,
                          void* xs, uint16_t* scenes, void* ws, uint64_t* ws_bytes, void* stream);
/* Same, with a coordinate map applied on the fly: the records hold x' = fma(x, scale, shift), clamped to [1e-6, 1 - 1e-6] when clamp01 != 0
 * (ray samples in [-1, 1]^3 -> scale = shift = 0.5: replaces the `x * 0.5 + 0.5` and `clamp` passes of LoTDEncoding.forward /
 * LoTD.forward, lotd_encoding.py:162, lotd.py:211).  The map is part of the records' fingerprint. */
int nr3d_lotd_sort_points_mapped(uint64_t N, const float* x, const int64_t* batch_inds, uint32_t batch_data_size, uint32_t n_scenes, int32_t force,
                                 float scale, float shift, int32_t clamp01, void* xs, uint16_t* scenes, void* ws, uint64_t* ws_bytes, void* stream);
/* Puts the stateful parts of a sort workspace (header, histogram counters, coarse bucket sizes: a few tens of MB) into their initial all-zero
 * state for calls with this configuration, on `stream`; has_batch_inds = whether those calls pass batch_inds.  Callers whose point count changes
 * from call to call (ray samples) keep ONE workspace of sufficient size and call this whenever (N, batching) changes, instead of allocating
 * and zero-filling a new one (the rank / record scratch behind the counters needs no initialisation). */
int nr3d_lotd_sort_ws_reset(uint64_t N, int32_t has_batch_inds, uint32_t batch_data_size, uint32_t n_scenes, void* ws, uint64_t ws_bytes, void* stream);
/* Test / A-B knob: point count from which single-scene calls take the two-level (coarse partition + fine) sort; 0 restores the default (12 Mi).
 * The workspace size of a given N depends on it: query it again afterwards. */
int nr3d_lotd_sort_set_two_level_min(uint64_t n_points);
int nr3d_lotd_fwd_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                         const void* params, int32_t max_level, void* y, int64_t y_stride_n, int64_t y_stride_f, void* stream);
int nr3d_lotd_bwd_param_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                               const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f, int32_t max_level, uint32_t pl_begin, uint32_t pl_end,
                               void* dL_dparam, void* stream);

/* The same fast path for callers that need the nablas (NeuS-style eikonal terms): forward with dy/dx and the second-order
 * d(dL/dx)/dparam . dL_ddLdx scatter (reference kernel_lod_hash_only_with_dydx / kernel_lod_hashonly_backward_input_backward_grid,
 * lotd_hash_only.h:164-378, 472-695).  y: [N, n_enc] param dtype, dy_dx: f32 [N, n_enc, 3], both row-major at the points'
 * original indices and fully written.  dL_ddLdx: f32 [N, 3]; dL_dparam is accumulated into (zero it first). */
int nr3d_lotd_fwd_dydx_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                              const void* params, int32_t max_level, void* y, void* dy_dx, void* stream);
int nr3d_lotd_bwd_param2_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                                const void* dL_dy, int64_t dLdy_stride_n, int64_t dLdy_stride_f, const float* dL_ddLdx, int32_t max_level,
                                void* dL_dparam, void* stream);

/* Sorted fast path with the M2 workload's density head fused in (no reference counterpart; replaces the composition
 * encoding(x) -> softplus(gain * sum_c h) -> alpha = 1 - exp(-sigma * delta) of the render step, nr3d_lib/graphics/nerf/nerf_ray_query.py:182,
 * and its autograd chain).  Forward: sigma, alpha f32 [N] at the points' original indices; the [N, n_enc] features are never written.
 * Backward: dL/dalpha f32 [N] (+ the saved sigma / alpha / deltas) -> dL/dparam, accumulated (zero it first); dL/dy is never materialised. */
int nr3d_lotd_density_head_fwd_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                                      const void* params, int32_t max_level, const float* deltas, float gain, float* sigma, float* alpha, void* stream);
int nr3d_lotd_density_head_bwd_sorted(const nr3d_lotd_meta* meta, int32_t param_dtype, uint64_t N, const void* xs, const uint16_t* scenes, uint32_t n_scenes,
                                      const float* d_alpha, const float* sigma, const float* alpha, const float* deltas, float gain, int32_t max_level,
                                      void* dL_dparam, void* stream);

/* Fused LoTD encode + density decoder, forward only.  Replaces the composition
 * LoTDNeRF.query_density (nr3d_lib/models/fields/nerf/lotd_nerf.py:169-178): encoding(x) -> Linear(32,64) -> ReLU ->
 * Linear(64, <=16) -> activation(out[...,0]), without writing the [N,32] features to HBM.  tcgen05 (bf16 operands, f32
 * accumulation).  xs: sorted records from nr3d_lotd_sort_points; params: fp32 table; w1_packed / w2_packed: bf16 weights
 * in K-major core-matrix order ([K/8][rows/8][8][8], rows = 64 resp. 16); b1 [64] / b2 [16] f32 or NULL;
 * activation: 0 identity, 1 exp, 2 softplus, 3 relu.  sigma [N] f32 and out16 ([N,16] f32, nullable) are written at the
 * points' original indices. */
int nr3d_lotd_fused_density_fwd(const nr3d_lotd_meta* meta, uint64_t N, const void* xs, const void* params, int32_t max_level,
                                const void* w1_packed, const float* b1, const void* w2_packed, const float* b2, int32_t activation,
                                float* sigma, float* out16, void* stream);

/* Backward of the fused encoder + decoder (training through LoTDNeRF.forward_density, lotd_nerf.py:136-178, and models/blocks/mlp.py):
 * one persistent kernel re-gathers the features, recomputes the hidden layer and runs  dH = G W2,  dF = (dH . relu') W1,
 * dW2 += G^T H,  [dW1 | db1] += (dH . relu')^T [F | 1]  on tcgen05 (weight-gradient accumulators stay in TMEM for the whole kernel), then
 * scatters dF into dL_dparam with the run-merged reductions of the sorted fast path.  G = dL/d(decoder output) [N, 16]:
 * column 0 = d_sigma * activation'(out0) (derived from the forward's `sigma`) + d_out16[:, 0], the other columns d_out16 (either may be NULL).
 * w2t_packed = W2^T [64, 16], w1t_packed = W1^T [32, 64], both bf16 K-major core-matrix order like w1_packed.
 * Outputs are ACCUMULATED (zero them first): dL_dparam [n_params] f32, dW1 [64, 32], db1 [64], dW2 [16, 64], db2 [16] f32.  No dL/dx. */
int nr3d_lotd_fused_density_bwd(const nr3d_lotd_meta* meta, uint64_t N, const void* xs, const void* params, int32_t max_level,
                                const void* w1_packed, const void* w2t_packed, const void* w1t_packed, const float* b1, int32_t activation,
                                const float* sigma, const float* d_sigma, const float* d_out16, float* dL_dparam, float* dW1, float* db1,
                                float* dW2, float* db2, void* stream);


/* ------------------------------------------------------------------------------------------------
 * occupancy-grid ray marching (replaces nr3d_lib.bindings._occ_grid, csrc/occ_grid/src/occ_grid.cpp:22-33)
 * ---------------------------------------------------------------------------------------------- */
enum { NR3D_CONTRACTION_AABB = 0, NR3D_CONTRACTION_TANH = 1, NR3D_CONTRACTION_SPHERE = 2 }; /* cpp_api.h:14-19 */

/* Pass 1 of ray_marching / batched_ray_marching (csrc/occ_grid/src/ray_marching.cu:17-134,179-203,
 * batched_marching.cu:18-148): writes num_steps[R] (int32) for every ray (0 for batch_inds<0).
 * batch_inds (int32 [R]) and batch_data_size select the batched variant; pass NULL / 0 with n_batches=1 for single.
 * roi: [n_batches, 6] float; grid: bool/uint8 [n_batches, rx, ry, rz]. */
int nr3d_march_count(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                     const int32_t* batch_inds, uint32_t batch_data_size, int32_t n_batches,
                     const float* roi, const uint8_t* grid, int32_t rx, int32_t ry, int32_t rz, int32_t contraction,
                     float step_size, float max_step_size, float dt_gamma, uint32_t max_steps,
                     int32_t* num_steps, void* stream);
/* On-device replacement of `cumsum` + `stack` + `.item()` (ray_marching.cu:205-209): packed_info[R,2] int32 =
 * (exclusive offset, count); total[0] (int64, device) = number of samples. ws: scratch, query size with ws==NULL. */
int nr3d_march_pack(uint64_t n_rays, const int32_t* num_steps, int32_t* packed_info, int64_t* total,
                    void* ws, uint64_t* ws_bytes, void* stream);
/* Pass 2 (ray_marching.cu:218-241): fills t_starts/t_ends [S] f32, ridx [S] i32, bidx (nullable), gidx (nullable). */
int nr3d_march_fill(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                    const int32_t* batch_inds, uint32_t batch_data_size, int32_t n_batches,
                    const float* roi, const uint8_t* grid, int32_t rx, int32_t ry, int32_t rz, int32_t contraction,
                    float step_size, float max_step_size, float dt_gamma, uint32_t max_steps,
                    const int32_t* packed_info, float* t_starts, float* t_ends, int32_t* ridx, int32_t* bidx,
                    int32_t* gidx, void* stream);

/* Single-pass variant of the two passes above (same outputs, the ray loop runs ONCE): nr3d_march_record counts like nr3d_march_count and
 * parks every sample as a 16-byte record (t_start, t_end, voxel id) in records[j * n_rays + ray] (16-byte aligned scratch of at least
 * n_rays * max_steps * 16 bytes); after nr3d_march_pack, nr3d_march_compact moves the records to their packed positions (ridx / bidx are
 * regenerated from the ray index and batch_inds / batch_data_size).  Callers choose it when the scratch fits their budget. */
int nr3d_march_record(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                      const int32_t* batch_inds, uint32_t batch_data_size, int32_t n_batches,
                      const float* roi, const uint8_t* grid, int32_t rx, int32_t ry, int32_t rz, int32_t contraction,
                      float step_size, float max_step_size, float dt_gamma, uint32_t max_steps,
                      int32_t* num_steps, void* records, uint64_t records_bytes, void* stream);
int nr3d_march_compact(uint64_t n_rays, const int32_t* batch_inds, uint32_t batch_data_size, const void* records,
                       const int32_t* packed_info, float* t_starts, float* t_ends, int32_t* ridx, int32_t* bidx, int32_t* gidx,
                       void* stream);

/* Forest (multi-block) marcher, SURVEY.md section 8f row n4: == forest_ray_marching (csrc/occ_grid/src/forest_marching.cu:27-303,
 * bound as nr3d_lib.bindings._occ_grid.forest_ray_marching, occ_grid.cpp:32).  Every ray owns a pack of block segments
 * (seg_pack_infos int32 [R,2] -> seg_block_inds int32 / seg_entries f32 / seg_exits f32 [n_segments]) from the octree ray
 * trace; block b occupies [world_origin + block_ks[b] * world_block_size, + world_block_size] (block_ks int16 [n_trees,3];
 * world_origin / world_block_size: 3 HOST floats each = ForestMeta.world_origin / world_block_size, forest_cpp_api.h:27-28)
 * and owns grid[b] (bool/uint8 [n_trees, rx, ry, rz]).  Count / pack (nr3d_march_pack) / fill like the single-grid marcher;
 * fill additionally writes blidx int32 [S]; gidx (nullable) = voxel index + b * rx*ry*rz. */
int nr3d_forest_march_count(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                            const int32_t* seg_block_inds, const float* seg_entries, const float* seg_exits,
                            const int32_t* seg_pack_infos, const int16_t* block_ks, const float* world_origin,
                            const float* world_block_size, const uint8_t* grid, int32_t rx, int32_t ry, int32_t rz,
                            float step_size, float max_step_size, float dt_gamma, uint32_t max_steps, int32_t* num_steps, void* stream);
int nr3d_forest_march_fill(uint64_t n_rays, const float* rays_o, const float* rays_d, const float* t_min, const float* t_max,
                           const int32_t* seg_block_inds, const float* seg_entries, const float* seg_exits,
                           const int32_t* seg_pack_infos, const int16_t* block_ks, const float* world_origin,
                           const float* world_block_size, const uint8_t* grid, int32_t rx, int32_t ry, int32_t rz,
                           float step_size, float max_step_size, float dt_gamma, uint32_t max_steps, const int32_t* packed_info,
                           float* t_starts, float* t_ends, int32_t* ridx, int32_t* blidx, int32_t* gidx, void* stream);

/* Post-processing of the marcher's output in one pass (replaces index_select x2 + addcmul + sub of
 * nr3d_lib/graphics/raymarch/occgrid_raymarch.py:96-107): samples[i] = rays_o[ridx[i]] + rays_d[ridx[i]] * t_starts[i]
 * (one FMA per component -- bit-identical to torch.addcmul on CUDA), deltas[i] = t_ends[i] - t_starts[i] (deltas / t_ends nullable).
 * ridx: int32 or int64 [S] (ridx_dtype = NR3D_I32 / NR3D_I64). */
int nr3d_march_samples(uint64_t S, const float* rays_o, const float* rays_d, const float* t_starts, const float* t_ends,
                       const void* ridx, int32_t ridx_dtype, float* samples, float* deltas, void* stream);

/* Density head + opacity in one pass each way (M2 workload): sigma = softplus(gain * sum_c h[i,c]),
 * alpha = 1 - exp(-sigma * deltas[i]) (nr3d_lib/graphics/nerf/nerf_ray_query.py:182).  The reference composes these from torch
 * element-wise / reduce kernels; there is no reference entry point -- these two serve nr3d_lib_b200.pipeline only.
 * h: f32 [S, C] with row stride h_stride (elements) and unit column stride.  Backward: d_h [S, C] contiguous =
 * (d_alpha * deltas * (1 - alpha) + d_sigma_extra) * gain * sigmoid(gain * sum) broadcast over C (d_sigma_extra nullable). */
int nr3d_density_alpha_fwd(uint64_t S, uint32_t C, const float* h, int64_t h_stride, const float* deltas, float gain, float* sigma,
                           float* alpha, void* stream);
int nr3d_density_alpha_bwd(uint64_t S, uint32_t C, const float* d_alpha, const float* d_sigma_extra, const float* sigma,
                           const float* alpha, const float* deltas, float gain, float* d_h, void* stream);

/* ------------------------------------------------------------------------------------------------
 * occupancy value-grid maintenance.  Replaces the torch + torch_scatter compositions of
 * nr3d_lib/models/accelerations/occgrid/utils.py:18-133 and ema_single.py:186-218.  Grids are [B, rx, ry, rz] (z fastest).
 * ---------------------------------------------------------------------------------------------- */
/* Pass 1 of update_[batched_]occ_val_grid[_idx]_ (utils.py:93-133): scratch[cell(i)] = max(scratch, ordered_bits(vals[i])).
 * Exactly one of pts ([N,3] f32 in [-1,1], cell = ((p/2+0.5)*res).long().clamp(0,res-1)) / gidx ([N,3] i64).
 * bidx: [N] i64 or NULL; with NULL and batch_data_size > 0 the batch of point i is i / batch_data_size.
 * scratch: u32 [B*rx*ry*rz], all zero on entry (nr3d_occ_apply leaves it all zero again). */
int nr3d_occ_scatter_max(uint64_t N, const float* pts, const int64_t* gidx, const int64_t* bidx, uint64_t batch_data_size,
                         const float* vals, uint32_t B, const int32_t* res, uint32_t* scratch, void* stream);
/* Pass 2: for every touched cell grid = max(ema_decay * grid, scattered max) (utils.py:101-102), scratch reset to zero;
 * write_occ != 0 also writes occ = grid > occ_thre (binarize, utils.py:84-87); sum != NULL accumulates the sum of the updated
 * grid (f64, must be zero on entry) for the mean-relative threshold. */
int nr3d_occ_apply(uint64_t n_cells, float* grid, uint32_t* scratch, float ema_decay, int32_t write_occ, float occ_thre,
                   uint8_t* occ, double* sum, void* stream);
/* == binarize (utils.py:84-87): occ = grid > (consider_mean ? min(sum/n - eps, occ_thre) : occ_thre). */
int nr3d_occ_binarize(uint64_t n_cells, const float* grid, float occ_thre, int32_t consider_mean, float eps, const double* sum,
                      uint8_t* occ, void* stream);
/* == arithmetic of sample_pts_in_voxels (utils.py:18-39): pts[i] = ((gidx[v] + offsets[i]) / res) * 2 - 1 with v = vidx[i]
 *    (random voxel per point) or i / n_per_vox (vidx == NULL); the random offsets / voxel picks stay the caller's RNG. */
int nr3d_occ_sample_in_voxels(uint64_t n_pts, const int64_t* gidx, const int64_t* vidx, uint64_t n_per_vox, const float* offsets,
                              const int32_t* res, float* pts, int64_t* vidx_out, void* stream);
/* == query (ema_single.py:214-218, ema_batched.py): out[i] = occ[b(i), cell(pts[i])]; out-of-range batch -> 0. */
int nr3d_occ_query(uint64_t N, const float* pts, const int64_t* bidx, uint64_t batch_data_size, uint32_t B, const int32_t* res,
                   const uint8_t* occ, uint8_t* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * pack_ops (replaces nr3d_lib.bindings._pack_ops, csrc/pack_ops/pack_ops.cpp:20-58)
 * pack_infos: int64 [P, 2] = (first index, length).  feats: [S, C] contiguous (C = 1 for 1-D tensors).
 * ---------------------------------------------------------------------------------------------- */
/* == packed_sum, pack_ops_cuda.cu:798-861.  out [P, C] fully written (0 for empty packs).  S = rows of `feats` (the tensor's size, which
 * the reference's pybind entry knows too): lets the staged kernels (TMA bulk copies of whole sample windows, pack_staged.cu) stay inside the
 * array; S = 0 selects the warp-per-pack kernels. */
int nr3d_pack_sum(int32_t dtype, uint64_t P, uint32_t C, uint64_t S, const void* feats, const int64_t* pack_infos, void* out, void* stream);
/* == packed_cumsum / packed_cumprod, pack_ops_cuda.cu:864-1095.  out [S, C] must be zero-filled where elements
 * are not covered by any pack. exclusive cumprod implements the DOCUMENTED semantics (leading 1,
 * nr3d_lib/graphics/pack_ops/pack_ops.py:149); bug_compat=1 reproduces the reference CUDA output (all zeros).
 * S = rows of `feats` (see nr3d_pack_sum): one-channel fp32 / fp64 calls run the staged sequential kernels (bit-identical to the reference's
 * per-thread loops for sums and inclusive products); other calls use warp scans (fp32 reassociation, tolerance 3e-5 on long packs). */
int nr3d_pack_cumsum(int32_t dtype, uint64_t P, uint32_t C, uint64_t S, const void* feats, const int64_t* pack_infos,
                     int32_t exclusive, int32_t reverse, void* out, void* stream);
int nr3d_pack_cumprod(int32_t dtype, uint64_t P, uint32_t C, uint64_t S, const void* feats, const int64_t* pack_infos,
                      int32_t exclusive, int32_t reverse, int32_t bug_compat, void* out, void* stream);
/* == packed_diff / packed_backward_diff, pack_ops_cuda.cu:1098-1334. edge / fill: [P, C] nullable, at most one. */
int nr3d_pack_diff(int32_t dtype, uint64_t P, uint32_t C, const void* feats, const int64_t* pack_infos,
                   const void* appends, const void* last_fill, void* out, void* stream);
int nr3d_pack_backward_diff(int32_t dtype, uint64_t P, uint32_t C, const void* feats, const int64_t* pack_infos,
                            const void* prepends, const void* first_fill, void* out, void* stream);
/* == packed_{add,sub,mul,div,gt,geq,lt,leq,eq,neq}, pack_ops_cuda.cu:1960-2540. op: 0 add 1 sub 2 mul 3 div
 * (out dtype = dtype), 5 gt 6 geq 7 lt 8 leq 9 eq 10 neq (out = uint8/bool). other: [P, C]. */
int nr3d_pack_binary(int32_t op, int32_t dtype, uint64_t P, uint32_t C, const void* feats, const void* other,
                     const int64_t* pack_infos, void* out, void* stream);
/* == packed_alpha_to_vw_forward, pack_ops_cuda.cu:1735-1793,1850-1911.  weights (nullable) must be zero-filled;
 * num_steps (nullable, int64 [P]) fully written; selector (nullable, bool [S]) must be zero-filled.  S = number of samples in `alphas`
 * (see nr3d_pack_sum).  Weights, selectors and counts are bit-identical to the reference's for every dtype. */
int nr3d_pack_alpha_to_vw_fwd(int32_t dtype, uint64_t P, uint64_t S, const void* alphas, const int64_t* pack_infos,
                              float early_stop_eps, float alpha_thre, void* weights, int64_t* num_steps,
                              uint8_t* selector, void* stream);
/* == packed_alpha_to_vw_backward, pack_ops_cuda.cu:1795-1848,1914-1958. grad_alphas must be zero-filled.  fp32 / fp64 results are
 * bit-identical to the reference build's (same sequential order, same fused multiply-adds). */
int nr3d_pack_alpha_to_vw_bwd(int32_t dtype, uint64_t P, uint64_t S, const void* weights, const void* grad_weights,
                              const void* alphas, const int64_t* pack_infos, float early_stop_eps, float alpha_thre,
                              void* grad_alphas, void* stream);
/* Accumulated opacity and expected depth of every pack in one pass (no reference export; replaces acc = packed_sum(w), depth = packed_sum(w * t)
 * of the render step, nr3d_lib/graphics/nerf/nerf_ray_query.py:182-188, and their autograd chain).  Values are bit-identical to that
 * composition (sequential sums, product rounded before the add).  S = samples in `weights` / `depths` (f32 [S], 16-byte aligned).
 * Backward: grad_weights [S] (elements outside every pack keep the caller's zeros) = grad_depth[p] * depths + grad_acc[p]; either grad may be NULL. */
int nr3d_pack_weighted_sums_fwd(uint64_t P, uint64_t S, const float* weights, const float* depths, const int64_t* pack_infos, float* acc, float* depth,
                                void* stream);
int nr3d_pack_weighted_sums_bwd(uint64_t P, uint64_t S, const float* depths, const int64_t* pack_infos, const float* grad_acc, const float* grad_depth,
                                float* grad_weights, void* stream);
/* int64 exclusive scan helper: pack_infos[P,2] = (exclusive cumsum(n), n); total[0] = sum.  Replaces the
 * `cumsum` + `stack` + `.item()` idiom (pack_ops_cuda.cu:579-581,1874-1875). */
int nr3d_pack_infos_from_counts(uint64_t P, const int64_t* counts, int64_t* pack_infos, int64_t* total,
                                void* ws, uint64_t* ws_bytes, void* stream);
/* == interleave_arange / interleave_linstep, pack_ops_cuda.cu:47-218: out[begin_p + j] = start_p + j*step_p;
 * nidx (nullable int64) = pack id.  starts / steps nullable -> scalar start / step (as double). */
int nr3d_pack_interleave_linstep(int32_t dtype, uint64_t P, const int64_t* pack_infos, const void* starts,
                                 const void* steps, double start, double step, void* out, int64_t* nidx, void* stream);
/* == interleave_sample_step_wrt_depth_clamped, pack_ops_cuda.cu:480-604 (two passes). */
int nr3d_pack_sample_step_count(int32_t dtype, uint64_t P, const void* nears, const void* fars, uint32_t max_steps,
                                double dt_gamma, double min_step, double max_step, int64_t* n_per_pack, void* stream);
int nr3d_pack_sample_step_fill(int32_t dtype, uint64_t P, const void* nears, const int64_t* pack_infos,
                               double dt_gamma, double min_step, double max_step, void* t_samples, void* deltas,
                               int64_t* nidx, void* stream);
/* == mark_pack_boundaries_cuda, pack_ops_cuda.cu:2765-2805. ids dtype in {U8,I8,I16,I32,I64}; out int32 [S]. */
int nr3d_pack_mark_boundaries(int32_t dtype, uint64_t S, const void* pack_ids, int32_t* boundaries, void* stream);

/* ---- hierarchical-sampling ops ("next" row n2 of SURVEY.md 8f) ---- */
/* == packed_searchsorted / packed_searchsorted_packed_vals, pack_ops_cuda.cu:1374-1503.  val_pack_infos == NULL: vals is
 * [P, num_to_search]; otherwise vals is packed by val_pack_infos.  pidx (int64, same shape as vals) fully written for every
 * listed query: begin + min(lower_bound, len-1). */
int nr3d_pack_searchsorted(int32_t dtype, uint64_t P, const void* bins, const int64_t* pack_infos, const void* vals,
                           uint32_t num_to_search, const int64_t* val_pack_infos, int64_t* pidx, void* stream);
/* == packed_invert_cdf, pack_ops_cuda.cu:1633-1733.  u_vals / samples / bin_idx: [P, num_to_sample]. */
int nr3d_pack_invert_cdf(int32_t dtype, uint64_t P, const void* bins, const void* cdfs, const int64_t* pack_infos, const void* u_vals,
                         uint32_t num_to_sample, void* samples, int64_t* bin_idx, void* stream);
/* == try_merge_two_packs_sorted_aligned, pack_ops_cuda.cu:1505-1631.  pack_infos_merged = pack infos of n_a + n_b;
 * pidx_a [S_a] must be zero-filled; pidx_b [S_b]. */
int nr3d_pack_merge_sorted_aligned(int32_t dtype, uint64_t P, const void* vals_a, const int64_t* pack_infos_a, const void* vals_b,
                                   const int64_t* pack_infos_b, const int64_t* pack_infos_merged, int64_t* pidx_a, int64_t* pidx_b,
                                   void* stream);
/* == packed_sort_qsort / packed_sort_thrust, pack_ops_cuda.cu:2556-2763: ascending in-place sort of every pack; idx (nullable,
 * int64 [S], pre-filled with arange) receives the same permutation. */
int nr3d_pack_sort(int32_t dtype, uint64_t P, void* vals, const int64_t* pack_infos, int64_t* idx, void* stream);
/* == packed_matmul, pack_ops_cuda.cu:2060-2085: out[i, o] = sum_k feats[i, k] * other[p(i), o, k]; feats [S, C], other [P, C_out, C]. */
int nr3d_pack_matmul(int32_t dtype, uint64_t P, uint32_t C, uint32_t C_out, const void* feats, const void* other, const int64_t* pack_infos,
                     void* out, void* stream);

/* == interleave_sample_step_wrt_depth_in_packed_segments, pack_ops_cuda.cu:606-795 (two passes: per-ray step counts, then
 *    t_samples / deltas / ray index / segment index written at the offsets of `pack_infos`). */
int nr3d_pack_seg_sample_count(int32_t dtype, uint64_t P, const void* nears, const void* fars, const void* entries, const void* exits,
                               const int64_t* seg_pack_infos, uint32_t max_steps, double dt_gamma, double min_step, double max_step,
                               int64_t* n_per_pack, void* stream);
int nr3d_pack_seg_sample_fill(int32_t dtype, uint64_t P, const void* nears, const void* fars, const void* entries, const void* exits,
                              const int64_t* seg_pack_infos, const int64_t* pack_infos, double dt_gamma, double min_step, double max_step,
                              void* t_samples, void* deltas, int64_t* nidx, int64_t* sidx, void* stream);
/* == octree_mark_consecutive_segments, pack_ops_cuda.cu:2807-2885.  point_hierarchies: int16 [n_points, 3]; marks: bool (1 byte),
 *    zero-initialised by the caller.  offset_fix = 0 reproduces the reference's un-offset `point_indices` walk. */
int nr3d_pack_mark_consecutive_segments(uint64_t P, const int64_t* pack_infos, const int32_t* pidx, const int16_t* point_hierarchies,
                                        int32_t offset_fix, uint8_t* mark_start, uint8_t* mark_end, void* stream);


#ifdef __cplusplus
}
#endif
#endif /* NR3D_B200_H */
