// pack_ops.cu -- segment ("pack") primitives of libnr3d_b200, written for sm_100a.
//
// The reference runs ONE THREAD PER PACK with a sequential loop (csrc/pack_ops/pack_ops_cuda.cu:798-1095,
// 1735-1848): neighbouring threads touch addresses one whole pack apart, so no access coalesces.  Here one WARP
// owns a pack: lanes read 32 consecutive elements (one 128-byte line), scans run through register shuffles, and the
// order-sensitive transmittance chain of the alpha composite is carried through the warp with a shuffle chain
// that reproduces the sequential float product bit-for-bit (so early-stop decisions and compaction counts match).
#include "common.cuh"
#include <type_traits>

namespace nr3d {

constexpr int kPackThreads = 256;
constexpr int kWarpsPerBlock = kPackThreads / 32;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }

// accumulate type: half accumulates in float (documented: the reference accumulates in half sequentially)
template <typename T> struct Acc { using type = T; };
template <> struct Acc<__half> { using type = float; };

template <typename T> __device__ __forceinline__ typename Acc<T>::type ld(const T* p) { return (typename Acc<T>::type)(*p); }
template <> __device__ __forceinline__ float ld<__half>(const __half* p) { return __half2float(*p); }
template <typename T, typename A> __device__ __forceinline__ void st(T* p, A v) { *p = (T)v; }
template <> __device__ __forceinline__ void st<__half, float>(__half* p, float v) { *p = __float2half_rn(v); }

template <typename A> __device__ __forceinline__ A shfl_up(A v, int d) { return __shfl_up_sync(kFull, v, d); }
template <typename A> __device__ __forceinline__ A shfl_idx(A v, int s) { return __shfl_sync(kFull, v, s); }
template <typename A> __device__ __forceinline__ A shfl_xor(A v, int m) { return __shfl_xor_sync(kFull, v, m); }

struct PackRange { uint64_t begin, len; };
__device__ __forceinline__ PackRange pack_range(const int64_t* __restrict__ pack_infos, uint64_t p) {
    const longlong2 pi = __ldg(reinterpret_cast<const longlong2*>(pack_infos) + p);
    PackRange r;
    r.begin = (uint64_t)pi.x;
    r.len = pi.y > 0 ? (uint64_t)pi.y : 0;
    return r;
}

// ------------------------------------------------------------------------------------------------
// packed_sum (pack_ops_cuda.cu:798-824)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kPackThreads)
pack_sum_kernel(uint64_t P, uint32_t C, const T* __restrict__ in, const int64_t* __restrict__ pack_infos, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const uint64_t p = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = lane_id();
    const PackRange r = pack_range(pack_infos, p);
    const T* src = in + r.begin * C;
    if (C == 1) {
        A acc = A(0);
        for (uint64_t j = lane; j < r.len; j += 32) acc += ld<T>(src + j);
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) acc += shfl_xor(acc, m);
        if (lane == 0) st<T, A>(out + p, acc);
    } else {
        for (uint32_t c = 0; c < C; ++c) {
            A acc = A(0);
            for (uint64_t j = lane; j < r.len; j += 32) acc += ld<T>(src + j * C + c);
#pragma unroll
            for (int m = 16; m > 0; m >>= 1) acc += shfl_xor(acc, m);
            if (lane == 0) st<T, A>(out + p * C + c, acc);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packed_cumsum / packed_cumprod (pack_ops_cuda.cu:864-1095): warp scan over one (pack, channel)
// ------------------------------------------------------------------------------------------------
template <typename T, bool PROD>
__global__ void __launch_bounds__(kPackThreads)
pack_scan_kernel(uint64_t P, uint32_t C, const T* __restrict__ in, const int64_t* __restrict__ pack_infos, bool exclusive,
                 bool reverse, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const uint64_t w = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (w >= P * C) return;
    const uint64_t p = w / C;
    const uint32_t c = (uint32_t)(w - p * C);
    const int lane = lane_id();
    const PackRange r = pack_range(pack_infos, p);
    const A ident = PROD ? A(1) : A(0);
    A carry = ident;
    for (uint64_t base = 0; base < r.len; base += 32) {
        const uint64_t j = base + lane;
        const bool ok = j < r.len;
        const uint64_t e = reverse ? (r.begin + r.len - 1 - j) : (r.begin + j);
        A x = ok ? ld<T>(in + e * C + c) : ident;
        A s = x;  // inclusive warp scan
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const A t = shfl_up(s, d);
            if (lane >= d) s = PROD ? (A)(s * t) : (A)(s + t);
        }
        const A total = shfl_idx(s, 31);
        A res;
        if (exclusive) {
            A prev = shfl_up(s, 1);
            if (lane == 0) prev = ident;
            res = PROD ? (A)(carry * prev) : (A)(carry + prev);
        } else {
            res = PROD ? (A)(carry * s) : (A)(carry + s);
        }
        if (ok) st<T, A>(out + e * C + c, res);
        carry = PROD ? (A)(carry * total) : (A)(carry + total);
    }
}

// ------------------------------------------------------------------------------------------------
// packed_diff / packed_backward_diff (pack_ops_cuda.cu:1098-1184)
// ------------------------------------------------------------------------------------------------
template <typename T, bool BACKWARD>
__global__ void __launch_bounds__(kPackThreads)
pack_diff_kernel(uint64_t P, uint32_t C, const T* __restrict__ in, const int64_t* __restrict__ pack_infos,
                 const T* __restrict__ edge, const T* __restrict__ fill, T* __restrict__ out) {
    using A = typename Acc<T>::type;
    const uint64_t p = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = lane_id();
    const PackRange r = pack_range(pack_infos, p);
    if (r.len == 0) return;
    const uint64_t n = r.len * C;
    const T* src = in + r.begin * C;
    T* dst = out + r.begin * C;
    for (uint64_t k = lane; k < n; k += 32) {
        const uint64_t j = k / C;
        const uint32_t c = (uint32_t)(k - j * C);
        if (!BACKWARD) {
            if (j + 1 < r.len) st<T, A>(dst + k, ld<T>(src + k + C) - ld<T>(src + k));
            else if (edge) st<T, A>(dst + k, ld<T>(edge + p * C + c) - ld<T>(src + k));
            else if (fill) dst[k] = fill[p * C + c];
        } else {
            if (j > 0) st<T, A>(dst + k, ld<T>(src + k) - ld<T>(src + k - C));
            else if (edge) st<T, A>(dst + k, ld<T>(src + k) - ld<T>(edge + p * C + c));
            else if (fill) dst[k] = fill[p * C + c];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packed binary ops (pack_ops_cuda.cu:1960-2249)
// ------------------------------------------------------------------------------------------------
template <typename T, int OP>
__global__ void __launch_bounds__(kPackThreads)
pack_binary_kernel(uint64_t P, uint32_t C, const T* __restrict__ in, const T* __restrict__ other,
                   const int64_t* __restrict__ pack_infos, void* __restrict__ out_) {
    using A = typename Acc<T>::type;
    const uint64_t p = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = lane_id();
    const PackRange r = pack_range(pack_infos, p);
    const uint64_t n = r.len * C;
    const T* src = in + r.begin * C;
    for (uint64_t k = lane; k < n; k += 32) {
        const uint32_t c = (uint32_t)(k % C);
        const A a = ld<T>(src + k), b = ld<T>(other + p * C + c);
        if (OP < 4) {
            T* dst = reinterpret_cast<T*>(out_) + r.begin * C;
            A v = OP == 0 ? (A)(a + b) : OP == 1 ? (A)(a - b) : OP == 2 ? (A)(a * b) : (A)(a / b);
            st<T, A>(dst + k, v);
        } else {
            uint8_t* dst = reinterpret_cast<uint8_t*>(out_) + r.begin * C;
            bool v = OP == 5 ? a > b : OP == 6 ? a >= b : OP == 7 ? a < b : OP == 8 ? a <= b : OP == 9 ? a == b : a != b;
            dst[k] = v ? 1 : 0;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packed_alpha_to_vw forward / backward (pack_ops_cuda.cu:1735-1848)
// Arithmetic traits reproduce the reference's per-step rounding for each scalar type.
// ------------------------------------------------------------------------------------------------
template <typename T> struct AlphaArith;
template <> struct AlphaArith<float> {
    using W = float;  // working type carried through the warp
    static __device__ __forceinline__ W load(const float* p) { return *p; }
    static __device__ __forceinline__ void store(float* p, W v) { *p = v; }
    static __device__ __forceinline__ W one_minus(W a) { return 1.f - a; }
    static __device__ __forceinline__ W mul(W a, W b) { return a * b; }
    static __device__ __forceinline__ W cast(float v) { return v; }
};
template <> struct AlphaArith<double> {
    using W = double;
    static __device__ __forceinline__ W load(const double* p) { return *p; }
    static __device__ __forceinline__ void store(double* p, W v) { *p = v; }
    static __device__ __forceinline__ W one_minus(W a) { return (double)(1.f) - a; }
    static __device__ __forceinline__ W mul(W a, W b) { return a * b; }
    static __device__ __forceinline__ W cast(float v) { return (double)v; }
};
template <> struct AlphaArith<__half> {  // values held as float but rounded to half after every operation
    using W = float;
    static __device__ __forceinline__ W load(const __half* p) { return __half2float(*p); }
    static __device__ __forceinline__ void store(__half* p, W v) { *p = __float2half_rn(v); }
    static __device__ __forceinline__ W one_minus(W a) { return 1.f - a; }                       // float, not rounded (1.f - Half -> float)
    static __device__ __forceinline__ W mul(W a, W b) { return __half2float(__float2half_rn(a * b)); }
    static __device__ __forceinline__ W cast(float v) { return __half2float(__float2half_rn(v)); }
};

template <typename T>
__global__ void __launch_bounds__(kPackThreads)
alpha_to_vw_fwd_kernel(uint64_t P, const T* __restrict__ alphas, const int64_t* __restrict__ pack_infos, float eps_, float thre_,
                       T* __restrict__ weights, int64_t* __restrict__ num_steps, uint8_t* __restrict__ selector) {
    using AR = AlphaArith<T>;
    using W = typename AR::W;
    const uint64_t p = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = lane_id();
    const PackRange r = pack_range(pack_infos, p);
    const W eps = AR::cast(eps_), thre = AR::cast(thre_);
    W T_run = AR::cast(1.f);
    int cnt = 0;
    bool stopped = false;
    for (uint64_t base = 0; base < r.len && !stopped; base += 32) {
        const uint64_t j = base + lane;
        const bool ok = j < r.len;
        const W a = ok ? AR::load(alphas + r.begin + j) : AR::cast(0.f);
        const bool skip = !ok || (a <= thre);
        const W f = skip ? AR::cast(1.f) : AR::one_minus(a);
        // sequential-order transmittance: T_j = ((T_run * f_0) * f_1) ... * f_{j-1}; multiplying by 1 is exact
        W Tj = T_run;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const W fk = shfl_idx(f, k);
            if (k < lane) Tj = AR::mul(Tj, fk);
        }
        const W T_end = AR::mul(shfl_idx(Tj, 31), shfl_idx(f, 31));
        // early stop: the first position (in pack order) whose transmittance fell below eps ends the ray
        const unsigned below = __ballot_sync(kFull, ok && (Tj < eps));
        const int first_below = below ? (__ffs(below) - 1) : 32;
        const bool live = ok && !skip && lane < first_below;
        if (live) {
            if (weights) AR::store(weights + r.begin + j, AR::mul(a, Tj));
            if (selector) selector[r.begin + j] = 1;
        }
        cnt += __popc(__ballot_sync(kFull, live));
        stopped = below != 0;
        T_run = T_end;
    }
    if (num_steps && lane == 0) num_steps[p] = cnt;
}

template <typename T>
__global__ void __launch_bounds__(kPackThreads)
alpha_to_vw_bwd_kernel(uint64_t P, const T* __restrict__ alphas, const T* __restrict__ weights, const T* __restrict__ grad_weights,
                       const int64_t* __restrict__ pack_infos, float eps_, float thre_, T* __restrict__ grad_alphas) {
    using AR = AlphaArith<T>;
    using W = typename AR::W;
    const uint64_t p = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = lane_id();
    const PackRange r = pack_range(pack_infos, p);
    const W eps = AR::cast(eps_), thre = AR::cast(thre_);
    // accum = sum_j grad_w[j] * w[j] over the whole pack
    W total = AR::cast(0.f);
    for (uint64_t j = lane; j < r.len; j += 32) total += AR::load(grad_weights + r.begin + j) * AR::load(weights + r.begin + j);
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) total += shfl_xor(total, m);
    W T_run = AR::cast(1.f);
    W done = AR::cast(0.f);  // sum of grad_w*w over the already processed samples
    bool stopped = false;
    for (uint64_t base = 0; base < r.len && !stopped; base += 32) {
        const uint64_t j = base + lane;
        const bool ok = j < r.len;
        const W a = ok ? AR::load(alphas + r.begin + j) : AR::cast(0.f);
        const bool skip = !ok || (a < thre);  // NOTE `<` here, `<=` in the forward pass (reference quirk)
        const W f = skip ? AR::cast(1.f) : AR::one_minus(a);
        W Tj = T_run;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const W fk = shfl_idx(f, k);
            if (k < lane) Tj = AR::mul(Tj, fk);
        }
        const W T_end = AR::mul(shfl_idx(Tj, 31), shfl_idx(f, 31));
        const unsigned below = __ballot_sync(kFull, ok && (Tj < eps));
        const int first_below = below ? (__ffs(below) - 1) : 32;
        const bool live = ok && !skip && lane < first_below;
        const W gw = ok ? AR::load(grad_weights + r.begin + j) : AR::cast(0.f);
        const W c = live ? gw * AR::load(weights + r.begin + j) : AR::cast(0.f);
        W s = c;  // inclusive scan of the processed contributions
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const W t = shfl_up(s, d);
            if (lane >= d) s += t;
        }
        const W chunk_total = shfl_idx(s, 31);
        if (live) {
            const W accum = total - (done + (s - c));
            const W denom = (W)fmaxf((float)AR::one_minus(a), 1e-10f);
            AR::store(grad_alphas + r.begin + j, (gw * Tj - accum) / denom);
        }
        done += chunk_total;
        stopped = below != 0;
        T_run = T_end;
    }
}

// ------------------------------------------------------------------------------------------------
// exclusive scan of per-pack counts -> pack_infos (three small kernels, no host round trip)
// ------------------------------------------------------------------------------------------------
constexpr int kScanBlock = 1024;

template <typename TI>
__global__ void __launch_bounds__(kScanBlock) scan_block_sums_kernel(uint64_t n, const TI* __restrict__ counts, int64_t* __restrict__ block_sums) {
    __shared__ int64_t warp_sums[32];
    const uint64_t i = (uint64_t)blockIdx.x * kScanBlock + threadIdx.x;
    int64_t v = i < n ? (int64_t)counts[i] : 0;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += __shfl_xor_sync(kFull, v, m);
    if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        int64_t s = warp_sums[threadIdx.x];
#pragma unroll
        for (int m = 16; m > 0; m >>= 1) s += __shfl_xor_sync(kFull, s, m);
        if (threadIdx.x == 0) block_sums[blockIdx.x] = s;
    }
}

// single block: exclusive scan of block_sums in place, total -> *total
__global__ void __launch_bounds__(kScanBlock) scan_of_block_sums_kernel(uint64_t nb, int64_t* __restrict__ block_sums, int64_t* __restrict__ total) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint64_t base = 0; base < nb; base += kScanBlock) {
        const uint64_t i = base + threadIdx.x;
        const int64_t x = i < nb ? block_sums[i] : 0;
        int64_t s = x;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(kFull, s, d);
            if ((threadIdx.x & 31) >= d) s += t;
        }
        if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x < 32) {
            int64_t ws = warp_sums[threadIdx.x];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int64_t t = __shfl_up_sync(kFull, ws, d);
                if (threadIdx.x >= d) ws += t;
            }
            warp_sums[threadIdx.x] = ws;  // inclusive over warps
        }
        __syncthreads();
        const int wid = threadIdx.x >> 5;
        const int64_t warp_off = wid ? warp_sums[wid - 1] : 0;
        const int64_t carry = carry_s;
        if (i < nb) block_sums[i] = carry + warp_off + s - x;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) carry_s = carry + warp_off + s;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry_s;
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(kScanBlock) scan_apply_kernel(uint64_t n, const TI* __restrict__ counts, const int64_t* __restrict__ block_offsets,
                                                                 TO* __restrict__ pack_infos) {
    __shared__ int64_t warp_sums[32];
    const uint64_t i = (uint64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const int64_t x = i < n ? (int64_t)counts[i] : 0;
    int64_t s = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const int64_t t = __shfl_up_sync(kFull, s, d);
        if ((threadIdx.x & 31) >= d) s += t;
    }
    if ((threadIdx.x & 31) == 31) warp_sums[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        int64_t ws = warp_sums[threadIdx.x];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const int64_t t = __shfl_up_sync(kFull, ws, d);
            if (threadIdx.x >= d) ws += t;
        }
        warp_sums[threadIdx.x] = ws;
    }
    __syncthreads();
    const int wid = threadIdx.x >> 5;
    const int64_t off = block_offsets[blockIdx.x] + (wid ? warp_sums[wid - 1] : 0) + s - x;
    if (i < n) {
        pack_infos[2 * i] = (TO)off;
        pack_infos[2 * i + 1] = (TO)x;
    }
}

template <typename TI, typename TO>
int scan_counts(uint64_t n, const TI* counts, TO* pack_infos, int64_t* total, void* ws, uint64_t* ws_bytes, cudaStream_t stream) {
    const uint64_t nb = div_up<uint64_t>(n ? n : 1, kScanBlock);
    const uint64_t need = nb * sizeof(int64_t);
    if (ws == nullptr) {
        if (ws_bytes) *ws_bytes = need;
        return 0;
    }
    NR3D_CHECK(ws_bytes && *ws_bytes >= need, "scan: workspace too small");
    int64_t* bs = reinterpret_cast<int64_t*>(ws);
    scan_block_sums_kernel<TI><<<(unsigned)nb, kScanBlock, 0, stream>>>(n, counts, bs);
    NR3D_LAUNCH_CHECK("scan_block_sums");
    scan_of_block_sums_kernel<<<1, kScanBlock, 0, stream>>>(nb, bs, total);
    NR3D_LAUNCH_CHECK("scan_of_block_sums");
    if (n) {
        scan_apply_kernel<TI, TO><<<(unsigned)nb, kScanBlock, 0, stream>>>(n, counts, bs, pack_infos);
        NR3D_LAUNCH_CHECK("scan_apply");
    }
    return 0;
}
template int scan_counts<int32_t, int32_t>(uint64_t, const int32_t*, int32_t*, int64_t*, void*, uint64_t*, cudaStream_t);
template int scan_counts<int64_t, int64_t>(uint64_t, const int64_t*, int64_t*, int64_t*, void*, uint64_t*, cudaStream_t);

// ------------------------------------------------------------------------------------------------
// interleave_linstep (pack_ops_cuda.cu:47-83), sample_step (pack_ops_cuda.cu:480-545), boundaries (:2765-2782)
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kPackThreads)
interleave_linstep_kernel(uint64_t P, const int64_t* __restrict__ pack_infos, const T* __restrict__ starts, const T* __restrict__ steps,
                          double start_s, double step_s, T* __restrict__ out, int64_t* __restrict__ nidx) {
    using A = typename Acc<T>::type;
    const uint64_t p = (uint64_t)blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    if (p >= P) return;
    const int lane = lane_id();
    const PackRange r = pack_range(pack_infos, p);
    const A start = starts ? ld<T>(starts + p) : (A)(T)start_s;
    const A step = steps ? ld<T>(steps + p) : (A)(T)step_s;
    for (uint64_t j = lane; j < r.len; j += 32) {
        // reference: out[j] = start + (scalar_t)j * step_size  (j is uint32 there)
        st<T, A>(out + r.begin + j, (A)(start + (A)(T)(uint32_t)j * step));
        if (nidx) nidx[r.begin + j] = (int64_t)p;
    }
}

template <typename T> __device__ __forceinline__ T clamp_t(T v, T lo, T hi) { return v < lo ? lo : (hi < v ? hi : v); }

template <typename T>
__global__ void sample_step_count_kernel(uint64_t P, uint32_t max_steps, T dt_gamma, T min_step, T max_step, const T* __restrict__ nears,
                                         const T* __restrict__ fars, int64_t* __restrict__ n_per_pack) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const T far = fars[p];
    T t = nears[p];
    uint32_t n = 0;
    while (t <= far && n < max_steps) {
        t += clamp_t<T>(t * dt_gamma, min_step, max_step);
        n++;
    }
    n_per_pack[p] = n;
}
template <typename T>
__global__ void sample_step_fill_kernel(uint64_t P, T dt_gamma, T min_step, T max_step, const T* __restrict__ nears,
                                        const int64_t* __restrict__ pack_infos, T* __restrict__ t_samples, T* __restrict__ deltas,
                                        int64_t* __restrict__ nidx) {
    const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    const uint64_t begin = (uint64_t)pack_infos[2 * p];
    const uint64_t n = (uint64_t)pack_infos[2 * p + 1];
    T t = nears[p];
    for (uint64_t s = 0; s < n; ++s) {
        t_samples[begin + s] = t;
        nidx[begin + s] = (int64_t)p;
        const T dt = clamp_t<T>(t * dt_gamma, min_step, max_step);
        deltas[begin + s] = dt;
        t += dt;
    }
}

template <typename T>
__global__ void mark_boundaries_kernel(uint64_t n, const T* __restrict__ ids, int32_t* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    out[i] = (i == 0) ? 1 : (ids[i - 1] == ids[i] ? 0 : 1);
}

// staged thread-per-pack kernels (pack_staged.cu): bit-identical to the reference's sequential loops, coalesced through shared memory
int staged_alpha_fwd(int32_t dtype, uint64_t P, uint64_t total, const void* alphas, const int64_t* pack_infos, float eps, float thre, void* weights,
                     int64_t* num_steps, uint8_t* selector, cudaStream_t st);
int staged_alpha_bwd(int32_t dtype, uint64_t P, uint64_t total, const void* alphas, const void* weights, const void* grad_weights,
                     const int64_t* pack_infos, float eps, float thre, void* grad_alphas, cudaStream_t st);
int staged_pack_sum(int32_t dtype, uint64_t P, uint64_t total, const void* in, const int64_t* pack_infos, void* out, cudaStream_t st);
int staged_pack_scan(int32_t dtype, bool prod, uint64_t P, uint64_t total, const void* in, const int64_t* pack_infos, bool exclusive, bool reverse, void* out,
                     cudaStream_t st);

#ifndef NR3D_PACK_STAGED      // 0: the warp-per-pack kernels of this file serve every call (A/B runs)
#define NR3D_PACK_STAGED 1
#endif

static inline unsigned warp_grid(uint64_t warps) { return (unsigned)div_up<uint64_t>(warps, kWarpsPerBlock); }

#define NR3D_PACK_DISPATCH(dtype, NAME, ...)                                                        \
    switch (dtype) {                                                                                \
    case NR3D_F32: { using T = float; __VA_ARGS__; } break;                                         \
    case NR3D_F64: { using T = double; __VA_ARGS__; } break;                                        \
    case NR3D_F16: { using T = __half; __VA_ARGS__; } break;                                        \
    case NR3D_I32: { using T = int32_t; __VA_ARGS__; } break;                                       \
    case NR3D_I64: { using T = int64_t; __VA_ARGS__; } break;                                       \
    default: return fail(NAME ": unsupported dtype code %d (supported: f32, f64, f16, i32, i64)", (int)dtype); }

#define NR3D_PACK_DISPATCH_FLOAT(dtype, NAME, ...)                                                  \
    switch (dtype) {                                                                                \
    case NR3D_F32: { using T = float; __VA_ARGS__; } break;                                         \
    case NR3D_F64: { using T = double; __VA_ARGS__; } break;                                        \
    case NR3D_F16: { using T = __half; __VA_ARGS__; } break;                                        \
    default: return fail(NAME ": expected a floating dtype (f16/f32/f64), got code %d", (int)dtype); }

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_pack_sum(int32_t dtype, uint64_t P, uint32_t C, uint64_t S, const void* feats, const int64_t* pack_infos, void* out, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(feats && pack_infos && out && C > 0, "packed_sum: null argument");
    if (NR3D_PACK_STAGED && C == 1 && S > 0 && staged_pack_sum(dtype, P, S, feats, pack_infos, out, (cudaStream_t)stream) == 0) {
        NR3D_LAUNCH_CHECK("packed_sum");
        return 0;
    }
    NR3D_PACK_DISPATCH(dtype, "packed_sum",
        (pack_sum_kernel<T><<<warp_grid(P), kPackThreads, 0, (cudaStream_t)stream>>>(P, C, (const T*)feats, pack_infos, (T*)out)));
    NR3D_LAUNCH_CHECK("packed_sum");
    return 0;
}

int nr3d_pack_cumsum(int32_t dtype, uint64_t P, uint32_t C, uint64_t S, const void* feats, const int64_t* pack_infos, int32_t exclusive,
                     int32_t reverse, void* out, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(feats && pack_infos && out && C > 0, "packed_cumsum: null argument");
    if (NR3D_PACK_STAGED && C == 1 && S > 0 && staged_pack_scan(dtype, false, P, S, feats, pack_infos, exclusive != 0, reverse != 0, out, (cudaStream_t)stream) == 0) {
        NR3D_LAUNCH_CHECK("packed_cumsum");
        return 0;
    }
    NR3D_PACK_DISPATCH(dtype, "packed_cumsum",
        (pack_scan_kernel<T, false><<<warp_grid(P * C), kPackThreads, 0, (cudaStream_t)stream>>>(P, C, (const T*)feats, pack_infos, exclusive != 0, reverse != 0, (T*)out)));
    NR3D_LAUNCH_CHECK("packed_cumsum");
    return 0;
}

int nr3d_pack_cumprod(int32_t dtype, uint64_t P, uint32_t C, uint64_t S, const void* feats, const int64_t* pack_infos, int32_t exclusive,
                      int32_t reverse, int32_t bug_compat, void* out, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(feats && pack_infos && out && C > 0, "packed_cumprod: null argument");
    if (exclusive && bug_compat) return 0;  // reference CUDA output: the zero-initialised tensor
    if (NR3D_PACK_STAGED && C == 1 && S > 0 && staged_pack_scan(dtype, true, P, S, feats, pack_infos, exclusive != 0, reverse != 0, out, (cudaStream_t)stream) == 0) {
        NR3D_LAUNCH_CHECK("packed_cumprod");
        return 0;
    }
    NR3D_PACK_DISPATCH(dtype, "packed_cumprod",
        (pack_scan_kernel<T, true><<<warp_grid(P * C), kPackThreads, 0, (cudaStream_t)stream>>>(P, C, (const T*)feats, pack_infos, exclusive != 0, reverse != 0, (T*)out)));
    NR3D_LAUNCH_CHECK("packed_cumprod");
    return 0;
}

int nr3d_pack_diff(int32_t dtype, uint64_t P, uint32_t C, const void* feats, const int64_t* pack_infos, const void* appends,
                   const void* last_fill, void* out, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(feats && pack_infos && out && C > 0, "packed_diff: null argument");
    NR3D_CHECK(!(appends && last_fill), "You should only specify AT MOST one of [appends, prepends, last_fill, first_fill]");
    NR3D_PACK_DISPATCH(dtype, "packed_diff",
        (pack_diff_kernel<T, false><<<warp_grid(P), kPackThreads, 0, (cudaStream_t)stream>>>(P, C, (const T*)feats, pack_infos, (const T*)appends, (const T*)last_fill, (T*)out)));
    NR3D_LAUNCH_CHECK("packed_diff");
    return 0;
}

int nr3d_pack_backward_diff(int32_t dtype, uint64_t P, uint32_t C, const void* feats, const int64_t* pack_infos, const void* prepends,
                            const void* first_fill, void* out, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(feats && pack_infos && out && C > 0, "packed_backward_diff: null argument");
    NR3D_CHECK(!(prepends && first_fill), "You should only specify AT MOST one of [appends, prepends, last_fill, first_fill]");
    NR3D_PACK_DISPATCH(dtype, "packed_backward_diff",
        (pack_diff_kernel<T, true><<<warp_grid(P), kPackThreads, 0, (cudaStream_t)stream>>>(P, C, (const T*)feats, pack_infos, (const T*)prepends, (const T*)first_fill, (T*)out)));
    NR3D_LAUNCH_CHECK("packed_backward_diff");
    return 0;
}

int nr3d_pack_binary(int32_t op, int32_t dtype, uint64_t P, uint32_t C, const void* feats, const void* other, const int64_t* pack_infos,
                     void* out, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(feats && other && pack_infos && out && C > 0, "packed binary op: null argument");
#define NR3D_BIN(OPC) case OPC: NR3D_PACK_DISPATCH(dtype, "packed binary op", \
        (pack_binary_kernel<T, OPC><<<warp_grid(P), kPackThreads, 0, (cudaStream_t)stream>>>(P, C, (const T*)feats, (const T*)other, pack_infos, out))); break;
    switch (op) {
        NR3D_BIN(0) NR3D_BIN(1) NR3D_BIN(2) NR3D_BIN(3) NR3D_BIN(5) NR3D_BIN(6) NR3D_BIN(7) NR3D_BIN(8) NR3D_BIN(9) NR3D_BIN(10)
    default: return fail("packed binary op: unsupported op code %d", (int)op);
    }
#undef NR3D_BIN
    NR3D_LAUNCH_CHECK("packed_binary");
    return 0;
}

int nr3d_pack_alpha_to_vw_fwd(int32_t dtype, uint64_t P, uint64_t S, const void* alphas, const int64_t* pack_infos, float early_stop_eps,
                              float alpha_thre, void* weights, int64_t* num_steps, uint8_t* selector, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(alphas && pack_infos, "packed_alpha_to_vw_forward: null argument");
    if (NR3D_PACK_STAGED && S > 0 && staged_alpha_fwd(dtype, P, S, alphas, pack_infos, early_stop_eps, alpha_thre, weights, num_steps, selector, (cudaStream_t)stream) == 0) {
        NR3D_LAUNCH_CHECK("packed_alpha_to_vw_forward");
        return 0;
    }
    NR3D_PACK_DISPATCH_FLOAT(dtype, "packed_alpha_to_vw_forward",
        (alpha_to_vw_fwd_kernel<T><<<warp_grid(P), kPackThreads, 0, (cudaStream_t)stream>>>(P, (const T*)alphas, pack_infos, early_stop_eps, alpha_thre, (T*)weights, num_steps, selector)));
    NR3D_LAUNCH_CHECK("packed_alpha_to_vw_forward");
    return 0;
}

int nr3d_pack_alpha_to_vw_bwd(int32_t dtype, uint64_t P, uint64_t S, const void* weights, const void* grad_weights, const void* alphas,
                              const int64_t* pack_infos, float early_stop_eps, float alpha_thre, void* grad_alphas, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(weights && grad_weights && alphas && pack_infos && grad_alphas, "packed_alpha_to_vw_backward: null argument");
    if (NR3D_PACK_STAGED && S > 0 && staged_alpha_bwd(dtype, P, S, alphas, weights, grad_weights, pack_infos, early_stop_eps, alpha_thre, grad_alphas, (cudaStream_t)stream) == 0) {
        NR3D_LAUNCH_CHECK("packed_alpha_to_vw_backward");
        return 0;
    }
    NR3D_PACK_DISPATCH_FLOAT(dtype, "packed_alpha_to_vw_backward",
        (alpha_to_vw_bwd_kernel<T><<<warp_grid(P), kPackThreads, 0, (cudaStream_t)stream>>>(P, (const T*)alphas, (const T*)weights, (const T*)grad_weights, pack_infos, early_stop_eps, alpha_thre, (T*)grad_alphas)));
    NR3D_LAUNCH_CHECK("packed_alpha_to_vw_backward");
    return 0;
}

int nr3d_pack_infos_from_counts(uint64_t P, const int64_t* counts, int64_t* pack_infos, int64_t* total, void* ws, uint64_t* ws_bytes,
                                void* stream) {
    return scan_counts<int64_t, int64_t>(P, counts, pack_infos, total, ws, ws_bytes, (cudaStream_t)stream);
}

int nr3d_pack_interleave_linstep(int32_t dtype, uint64_t P, const int64_t* pack_infos, const void* starts, const void* steps,
                                 double start, double step, void* out, int64_t* nidx, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(pack_infos && out, "interleave_linstep: null argument");
    NR3D_PACK_DISPATCH(dtype, "interleave_linstep",
        (interleave_linstep_kernel<T><<<warp_grid(P), kPackThreads, 0, (cudaStream_t)stream>>>(P, pack_infos, (const T*)starts, (const T*)steps, start, step, (T*)out, nidx)));
    NR3D_LAUNCH_CHECK("interleave_linstep");
    return 0;
}

int nr3d_pack_sample_step_count(int32_t dtype, uint64_t P, const void* nears, const void* fars, uint32_t max_steps, double dt_gamma,
                                double min_step, double max_step, int64_t* n_per_pack, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(nears && fars && n_per_pack, "interleave_sample_step_wrt_depth_clamped: null argument");
    const unsigned grid = (unsigned)div_up<uint64_t>(P, 128);
    switch (dtype) {
    case NR3D_F32: sample_step_count_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(P, max_steps, (float)dt_gamma, (float)min_step, (float)max_step, (const float*)nears, (const float*)fars, n_per_pack); break;
    case NR3D_F64: sample_step_count_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(P, max_steps, dt_gamma, min_step, max_step, (const double*)nears, (const double*)fars, n_per_pack); break;
    default: return fail("interleave_sample_step_wrt_depth_clamped: supported dtypes are f32/f64 (got code %d)", (int)dtype);
    }
    NR3D_LAUNCH_CHECK("sample_step_count");
    return 0;
}

int nr3d_pack_sample_step_fill(int32_t dtype, uint64_t P, const void* nears, const int64_t* pack_infos, double dt_gamma, double min_step,
                               double max_step, void* t_samples, void* deltas, int64_t* nidx, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(nears && pack_infos && t_samples && deltas && nidx, "interleave_sample_step_wrt_depth_clamped: null argument");
    const unsigned grid = (unsigned)div_up<uint64_t>(P, 128);
    switch (dtype) {
    case NR3D_F32: sample_step_fill_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(P, (float)dt_gamma, (float)min_step, (float)max_step, (const float*)nears, pack_infos, (float*)t_samples, (float*)deltas, nidx); break;
    case NR3D_F64: sample_step_fill_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(P, dt_gamma, min_step, max_step, (const double*)nears, pack_infos, (double*)t_samples, (double*)deltas, nidx); break;
    default: return fail("interleave_sample_step_wrt_depth_clamped: supported dtypes are f32/f64 (got code %d)", (int)dtype);
    }
    NR3D_LAUNCH_CHECK("sample_step_fill");
    return 0;
}

int nr3d_pack_mark_boundaries(int32_t dtype, uint64_t S, const void* pack_ids, int32_t* boundaries, void* stream) {
    if (S == 0) return 0;
    NR3D_CHECK(pack_ids && boundaries, "mark_pack_boundaries_cuda: null argument");
    const unsigned grid = (unsigned)div_up<uint64_t>(S, 256);
    cudaStream_t st_ = (cudaStream_t)stream;
    switch (dtype) {
    case NR3D_U8: mark_boundaries_kernel<uint8_t><<<grid, 256, 0, st_>>>(S, (const uint8_t*)pack_ids, boundaries); break;
    case NR3D_I8: mark_boundaries_kernel<int8_t><<<grid, 256, 0, st_>>>(S, (const int8_t*)pack_ids, boundaries); break;
    case NR3D_I16: mark_boundaries_kernel<int16_t><<<grid, 256, 0, st_>>>(S, (const int16_t*)pack_ids, boundaries); break;
    case NR3D_I32: mark_boundaries_kernel<int32_t><<<grid, 256, 0, st_>>>(S, (const int32_t*)pack_ids, boundaries); break;
    case NR3D_I64: mark_boundaries_kernel<int64_t><<<grid, 256, 0, st_>>>(S, (const int64_t*)pack_ids, boundaries); break;
    default: return fail("mark_pack_boundaries_cuda: expected an integral dtype, got code %d", (int)dtype);
    }
    NR3D_LAUNCH_CHECK("mark_pack_boundaries");
    return 0;
}

}  // extern "C"
