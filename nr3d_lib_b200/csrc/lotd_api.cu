// lotd_api.cu -- C-ABI entry points of the LoTD encoder + host-side meta construction.
#include "lotd_kernels.cuh"
#include <limits>
#include <string.h>

namespace nr3d {

static thread_local char g_err[1024];
char* tls_error_buffer() { return g_err; }
std::atomic<uint64_t> g_launch_count{0};

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return -1;
}

static bool all_divisible(const int32_t* v, int n, int k) {
    for (int i = 0; i < n; ++i)
        if (v[i] % k != 0) return false;
    return true;
}

int build_launch(const nr3d_lotd_meta* m, int32_t input_dtype, int32_t param_dtype, uint64_t N, const void* x,
                        const void* params, const int64_t* batch_inds, const int64_t* batch_offsets,
                        uint32_t batch_data_size, int32_t max_level, void* stream, LotdLaunch& L) {
    NR3D_CHECK(m != nullptr, "LoTDEncoding: null meta");
    NR3D_CHECK((input_dtype == NR3D_F32 && (param_dtype == NR3D_F32 || param_dtype == NR3D_F16)) || (input_dtype == NR3D_F16 && param_dtype == NR3D_F16),
               "LoTDEncoding: Input type combination not supported. Supported types are: "
               "<input,param> -> (half, half), (float, half), (float, float)");
    NR3D_CHECK(m->n_levels <= NR3D_MAX_LEVELS && m->n_pseudo_levels <= NR3D_MAX_PSEUDO_LEVELS, "LoTDEncoding: corrupt meta");
    NR3D_CHECK(N < (1ull << 32), "LoTDEncoding: batch_size must be < 2^32");
    memset(&L.tab, 0, sizeof(L.tab));
    for (uint32_t l = 0; l < m->n_levels; ++l) {
        LevelDesc& d = L.tab.lv[l];
        for (int k = 0; k < 4; ++k) d.res[k] = m->level_res[l][k];
        d.type = m->level_types[l];
        d.n_feat = m->level_n_feats[l];
        d.size = m->level_sizes[l];
        d.offset = m->level_offsets[l];
    }
    for (uint32_t p = 0; p < m->n_pseudo_levels; ++p) {
        L.tab.map_level[p] = (uint8_t)m->map_levels[p];
        L.tab.map_cnt[p] = (uint8_t)m->map_cnt[p];
    }
    L.tab.n_levels = m->n_levels;
    L.tab.n_pseudo = m->n_pseudo_levels;
    L.tab.n_enc = m->n_encoded_dims;
    L.tab.n_params = m->n_params;
    L.tab.interp = m->interpolation_type;
    L.tab.fpl = m->n_feat_per_pseudo_lvl;
    L.in.N = N;
    L.in.x = (const float*)x;
    L.in.x_half = input_dtype == NR3D_F16 ? 1u : 0u;
    L.in.params = params;
    L.in.batch_inds = batch_inds;
    L.in.batch_offsets = batch_offsets;
    L.in.batch_data_size = batch_data_size;
    L.in.max_level = max_level;
    // 8-byte vector access needs (params base + batch offset + level offset) 8-byte aligned: level offsets and
    // n_params are even (feature widths are even), user batch_offsets may be anything.
    L.in.vec_ok = (batch_offsets == nullptr) && ((reinterpret_cast<uintptr_t>(params) & 7u) == 0);
    L.fpl = (int)m->n_feat_per_pseudo_lvl;
    L.half = param_dtype == NR3D_F16;
    L.stream = (cudaStream_t)stream;
    return 0;
}

#define NR3D_DISPATCH_DIM(meta, CALL)                                                        \
    switch ((meta)->n_dims_to_encode) {                                                      \
    case 2: { constexpr int D = 2; return CALL; }                                            \
    case 3: { constexpr int D = 3; return CALL; }                                            \
    case 4: { constexpr int D = 4; return CALL; }                                            \
    default: return fail("LoTDEncoding: `n_dims_to_encode` must be 2, 3 or 4."); }

}  // namespace nr3d

using namespace nr3d;

extern "C" {

const char* nr3d_last_error(void) { return tls_error_buffer(); }
int nr3d_version(void) { return 100; }
uint64_t nr3d_launch_count(void) { return g_launch_count.load(); }

int nr3d_lotd_meta_create(int32_t n_dims, int32_t n_levels, const int32_t* res, const int32_t* n_feats,
                          const int32_t* types, uint32_t hashmap_size, int32_t use_smooth_step, nr3d_lotd_meta* out) {
    NR3D_CHECK(out && res && n_feats && types, "LoTDEncoding: null argument");
    NR3D_CHECK(n_dims == 2 || n_dims == 3 || n_dims == 4, "LoTDEncoding: `n_input_dim` must be 2/3/4.");
    NR3D_CHECK(n_levels >= 0 && n_levels <= NR3D_MAX_LEVELS, "LoTDEncoding:` num_level`=%d exceeds maximum level=%d", n_levels, NR3D_MAX_LEVELS);
    memset(out, 0, sizeof(*out));
    out->interpolation_type = use_smooth_step ? NR3D_INTERP_SMOOTHSTEP : NR3D_INTERP_LINEAR;
    out->n_dims_to_encode = (uint32_t)n_dims;
    out->n_levels = (uint32_t)n_levels;
    if (all_divisible(n_feats, n_levels, 8)) out->n_feat_per_pseudo_lvl = 8;
    else if (all_divisible(n_feats, n_levels, 4)) out->n_feat_per_pseudo_lvl = 4;
    else if (all_divisible(n_feats, n_levels, 2)) out->n_feat_per_pseudo_lvl = 2;
    else return fail("LoTDEncoding: the greatest common divisor of `lod_n_feats` must be at least 2");
    const uint32_t fpl = out->n_feat_per_pseudo_lvl;
    const uint32_t max_params = std::numeric_limits<uint32_t>::max() / 2;
    uint32_t acc = 0;
    float acc_f = 0.f;
    out->hash_only = 1;
    uint32_t n_pseudo = 0;
    for (int l = 0; l < n_levels; ++l) {
        const int32_t nf = n_feats[l];
        const int32_t tp = types[l];
        NR3D_CHECK(nf > 0, "LoTDEncoding: feature width must be positive");
        NR3D_CHECK(tp >= 0 && tp <= 7, "LoTDEncoding: Invalid lod type");
        if (tp != NR3D_LOD_DENSE && tp != NR3D_LOD_HASH) out->hash_only = 0;
        out->level_n_feats[l] = (uint32_t)nf;
        out->level_types[l] = (uint32_t)tp;
        n_pseudo += (uint32_t)nf / fpl;
        out->n_encoded_dims += (uint32_t)nf;
        uint32_t R[NR3D_MAX_DIMS] = {0, 0, 0, 0};
        for (int d = 0; d < n_dims; ++d) {
            const int32_t r = res[l * n_dims + d];
            NR3D_CHECK(r > 2, "LoTDEncoding: only support grid resolutions >= 3");
            R[d] = (uint32_t)r;
            out->level_res[l][d] = R[d];
        }
        uint32_t size = 0;
        float size_f = 0.f;
        switch (tp) {
        case NR3D_LOD_DENSE:
            size = 1; size_f = 1.f;
            for (int d = 0; d < n_dims; ++d) { size *= R[d]; size_f *= (float)R[d]; }
            break;
        case NR3D_LOD_NPLANEMUL:
        case NR3D_LOD_NPLANESUM:
        case NR3D_LOD_VM:
            if (tp == NR3D_LOD_VM) NR3D_CHECK(n_dims == 3, "LoTDEncoding: VectorMatrix mode only support 3D encoding.");
            for (int k = 0; k < n_dims; ++k) {
                uint32_t ps = 1; float psf = 1.f;
                for (int d = 0; d < n_dims; ++d) if (d != k) { ps *= R[d]; psf *= (float)R[d]; }
                if (tp == NR3D_LOD_VM) { size += ps + R[k]; size_f += (float)(ps + R[k]); }
                else { size += ps; size_f += psf; }
            }
            break;
        case NR3D_LOD_VECZMATXOY:
            NR3D_CHECK(n_dims == 3, "LoTDEncoding: VecZMatXoY mode only support 3D encoding.");
            size = R[0] * R[1] + R[2];
            size_f = (float)R[0] * (float)R[1] + (float)R[2];
            break;
        case NR3D_LOD_CP:
        case NR3D_LOD_CPFAST:
            for (int d = 0; d < n_dims; ++d) { size += R[d]; size_f += (float)R[d]; }
            break;
        case NR3D_LOD_HASH:
            NR3D_CHECK(hashmap_size != 0, "LoTDEncoding: Hash mode need `hashmap_size`");
            size = hashmap_size; size_f = (float)hashmap_size;
            break;
        }
        acc_f += size_f * (float)nf;
        NR3D_CHECK(!(acc_f > (float)max_params), "LoTDEncoding: param size too large.");
        out->level_sizes[l] = size;
        out->level_n_params[l] = size * (uint32_t)nf;
        out->level_offsets[l] = acc;
        acc += size * (uint32_t)nf;
    }
    out->level_offsets[n_levels] = acc;
    out->n_params = acc;
    NR3D_CHECK(n_pseudo <= NR3D_MAX_PSEUDO_LEVELS, "LoTDEncoding: too many pseudo levels (%u > %d)", n_pseudo, NR3D_MAX_PSEUDO_LEVELS);
    out->n_pseudo_levels = n_pseudo;
    uint32_t p = 0;
    for (int l = 0; l < n_levels; ++l) {
        const uint32_t n = (uint32_t)n_feats[l] / fpl;
        for (uint32_t j = 0; j < n; ++j) { out->map_levels[p + j] = (uint32_t)l; out->map_cnt[p + j] = j; }
        p += n;
    }
    NR3D_CHECK(out->n_encoded_dims <= 1024, "LoTDEncoding: total number of features too large. Shoule be <= 1024.");
    return 0;
}

int nr3d_lotd_fwd(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N, const void* x,
                  const void* params, const int64_t* batch_inds, const int64_t* batch_offsets, uint32_t batch_data_size,
                  int32_t max_level, void* y, int64_t y_stride_n, int64_t y_stride_f, void* dy_dx, int64_t dydx_stride_n,
                  int64_t dydx_stride_f, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    if (int rc = build_launch(meta, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L)) return rc;
    NR3D_CHECK(y != nullptr, "LoTDEncoding::fwd: null output");
    NR3D_DISPATCH_DIM(meta, lotd_launch_fwd<D>(L, y, y_stride_n, y_stride_f, (float*)dy_dx, dydx_stride_n, dydx_stride_f));
}

int nr3d_lotd_bwd_param(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N, const void* dL_dy,
                        int64_t s_n, int64_t s_f, const void* x, const void* params, const int64_t* batch_inds,
                        const int64_t* batch_offsets, uint32_t batch_data_size, int32_t max_level, void* dL_dparam, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    if (int rc = build_launch(meta, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L)) return rc;
    NR3D_CHECK(dL_dy && dL_dparam, "LoTDEncoding::bwd: null argument");
    L.in.vec_ok = L.in.vec_ok && ((reinterpret_cast<uintptr_t>(dL_dparam) & 7u) == 0);
    NR3D_DISPATCH_DIM(meta, lotd_launch_bwd_param<D>(L, dL_dy, s_n, s_f, nullptr, dL_dparam));
}

int nr3d_lotd_bwd_param_scenes(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N, const void* dL_dy,
                               int64_t s_n, int64_t s_f, const void* dL_ddLdx, const void* x, const void* params, const int64_t* batch_inds,
                               const int64_t* batch_offsets, uint32_t batch_data_size, uint32_t n_batches, int32_t max_level,
                               void* dL_dparam, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    if (int rc = build_launch(meta, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L)) return rc;
    NR3D_CHECK(dL_dy && dL_dparam, "LoTDEncoding::bwd: null argument");
    L.in.vec_ok = L.in.vec_ok && ((reinterpret_cast<uintptr_t>(dL_dparam) & 7u) == 0);
    L.n_batches = n_batches;
    NR3D_DISPATCH_DIM(meta, lotd_launch_bwd_param<D>(L, dL_dy, s_n, s_f, (const float*)dL_ddLdx, dL_dparam));
}

int nr3d_lotd_bwd_input(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N, const void* dL_dy,
                        int64_t s_n, int64_t s_f, const void* dy_dx, int64_t ds_n, int64_t ds_f, void* dL_dx, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    if (int rc = build_launch(meta, input_dtype, param_dtype, N, nullptr, nullptr, nullptr, nullptr, 0, 0, stream, L)) return rc;
    NR3D_CHECK(dL_dy && dy_dx && dL_dx, "LoTDEncoding::bwd: need `dy_dx` to comput `dL_dx`.");
    NR3D_DISPATCH_DIM(meta, lotd_launch_bwd_input<D>(L, dL_dy, s_n, s_f, (const float*)dy_dx, ds_n, ds_f, (float*)dL_dx));
}

int nr3d_lotd_bwd_bwd_input(const nr3d_lotd_meta* meta, int32_t input_dtype, int32_t param_dtype, uint64_t N,
                            const void* dL_ddLdx, const void* dL_dy, int64_t s_n, int64_t s_f, const void* x,
                            const void* params, const void* dy_dx, int64_t ds_n, int64_t ds_f, const int64_t* batch_inds,
                            const int64_t* batch_offsets, uint32_t batch_data_size, int32_t max_level, void* dL_ddLdy,
                            void* dL_dparam, void* dL_dx, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    if (int rc = build_launch(meta, input_dtype, param_dtype, N, x, params, batch_inds, batch_offsets, batch_data_size, max_level, stream, L)) return rc;
    NR3D_CHECK(dL_ddLdx && dL_dy, "LoTDEncoding::bwd_bwd_input: null argument");
    if (dL_ddLdy) {
        NR3D_CHECK(dy_dx != nullptr, "LoTDEncoding::bwd_bwd_input: need `dy_dx` to compute `dL_d(dLdy)`.");
        int rc = 0;
        switch (meta->n_dims_to_encode) {
        case 2: rc = lotd_launch_ddLdy<2>(L, (const float*)dL_ddLdx, (const float*)dy_dx, ds_n, ds_f, dL_ddLdy); break;
        case 3: rc = lotd_launch_ddLdy<3>(L, (const float*)dL_ddLdx, (const float*)dy_dx, ds_n, ds_f, dL_ddLdy); break;
        case 4: rc = lotd_launch_ddLdy<4>(L, (const float*)dL_ddLdx, (const float*)dy_dx, ds_n, ds_f, dL_ddLdy); break;
        default: return fail("LoTDEncoding: `n_dims_to_encode` must be 2, 3 or 4.");
        }
        if (rc) return rc;
    }
    if (dL_dx) {
        int rc = 0;
        switch (meta->n_dims_to_encode) {
        case 2: rc = lotd_launch_bwdbwd_input<2>(L, dL_dy, s_n, s_f, (const float*)dL_ddLdx, (float*)dL_dx); break;
        case 3: rc = lotd_launch_bwdbwd_input<3>(L, dL_dy, s_n, s_f, (const float*)dL_ddLdx, (float*)dL_dx); break;
        case 4: rc = lotd_launch_bwdbwd_input<4>(L, dL_dy, s_n, s_f, (const float*)dL_ddLdx, (float*)dL_dx); break;
        default: return fail("LoTDEncoding: `n_dims_to_encode` must be 2, 3 or 4.");
        }
        if (rc) return rc;
    }
    if (dL_dparam) {
        L.in.vec_ok = L.in.vec_ok && ((reinterpret_cast<uintptr_t>(dL_dparam) & 7u) == 0);
        NR3D_DISPATCH_DIM(meta, lotd_launch_bwd_param<D>(L, dL_dy, s_n, s_f, (const float*)dL_ddLdx, dL_dparam));
    }
    return 0;
}

int nr3d_lotd_grid_index(const nr3d_lotd_meta* meta, int32_t input_dtype, uint64_t N, const void* x, const int64_t* batch_inds,
                         const int64_t* batch_offsets, uint32_t batch_data_size, int32_t max_level, int64_t* out, void* stream) {
    if (N == 0) return 0;
    LotdLaunch L;
    if (int rc = build_launch(meta, input_dtype, input_dtype == NR3D_F16 ? NR3D_F16 : NR3D_F32, N, x, nullptr, batch_inds, batch_offsets, batch_data_size, max_level, stream, L)) return rc;
    for (uint32_t l = 0; l < meta->n_levels; ++l)
        NR3D_CHECK(meta->level_types[l] == NR3D_LOD_DENSE || meta->level_types[l] == NR3D_LOD_HASH,
                   "LoTDEncoding::get_grid_index: Only support Dense/Hash type.");
    NR3D_CHECK(out != nullptr, "LoTDEncoding::get_grid_index: null output");
    NR3D_DISPATCH_DIM(meta, lotd_launch_grid_index<D>(L, out));
}

}  // extern "C"
