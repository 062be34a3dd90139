// pack_staged.cu -- the order-sensitive pack primitives (alpha composite forward / backward, one-channel pack sum) as "staged
// thread-per-pack" kernels for sm_100a.
//
// The reference runs one THREAD per pack with a sequential loop (csrc/pack_ops/pack_ops_cuda.cu:798-824, 1735-1848): the arithmetic
// order is what makes its early-stop decisions, compaction counts and rounding what they are, but neighbouring threads touch addresses
// one whole pack apart, so no access coalesces.  The first B200 version (pack_ops.cu) gave each pack a warp and carried the sequential
// product through a 32-step shuffle chain: coalesced, bit-exact, but 64 issue slots per 32 samples and one short pack per warp
// (ncu, 30 Mi samples: 11 % of HBM bandwidth forward, 21 % backward -- profiles/r1_m2_ncu_summary.txt).
// Here a CTA of 256 threads owns 256 CONSECUTIVE packs.  When they tile one contiguous span of samples (always, for pack_infos built
// from a cumsum) the span is walked in windows:
//   1. ONE THREAD issues a TMA bulk copy (cp.async.bulk.shared::cluster.global, SASS UBLKCP) of the window's samples into shared memory
//      and the CTA waits on the mbarrier -- the window is contiguous in HBM, so this is the one place on the hot path where a bulk copy
//      fits the data (the few elements outside the 16-byte aligned interior come in through ordinary loads);
//   2. every thread walks ITS OWN pack's part of the window sequentially out of shared memory -- the reference's loop, statement for
//      statement, so results are bit-identical to the reference thread's (hundreds of independent chains per CTA instead of one per warp);
//      the state (transmittance, counters) stays in registers from window to window;
//   3. the CTA writes the window's outputs back with coalesced stores (every element of the span is written, zeros included).
// Packs that do not tile a span (gaps, overlaps, arbitrary order) are handled one pack per span through the same code.
//
// Sizing (measured on B200, 262144 packs x 115 samples, profiles/r2_pack_march_bench*.txt and the ncu capture next to them): the chain is
// latency bound (one dependent step every ~40 clocks per thread), so what counts is (a) how many lanes of a warp have work in a window and
// (b) how many warps per SM are inside a chain at once.  256-thread CTAs with 4096-sample windows had 36 of 256 threads busy per window and
// ran 1.15 waves (154 us forward); ONE-WARP CTAs whose window covers the whole span of their 32 packs keep every lane busy, need no
// block barrier, and 8192 small CTAs balance themselves over the SMs.  The chain is unrolled by four with the loads hoisted above the
// stores of the same shared array (the compiler must otherwise order every load behind the previous store).
#include "common.cuh"

namespace nr3d {

constexpr int kCtThreads = 32;                // packs per CTA: ONE WARP -- see the sizing note below
__device__ __forceinline__ uint64_t umin64(uint64_t a, uint64_t b) { return a < b ? a : b; }
__device__ __forceinline__ uint64_t umax64(uint64_t a, uint64_t b) { return a > b ? a : b; }
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

struct CtaSpan {
    uint64_t begin, len;     // this thread's pack
    uint64_t sb, se;         // span of the CTA's packs when they tile it
    bool tiled;
    int last;                // last thread that owns a pack
};

// s_b / s_e: [kCtThreads + 1] / [kCtThreads] scratch
__device__ __forceinline__ CtaSpan load_cta_span(const int64_t* __restrict__ pack_infos, uint64_t P, uint64_t p0, int tid, uint64_t* s_b, uint64_t* s_e) {
    CtaSpan s;
    const uint64_t p = p0 + tid;
    s.begin = 0; s.len = 0;
    if (p < P) {
        const longlong2 pi = __ldg(reinterpret_cast<const longlong2*>(pack_infos) + p);
        s.begin = (uint64_t)pi.x;
        s.len = pi.y > 0 ? (uint64_t)pi.y : 0;
    }
    s_b[tid] = s.begin;
    s_e[tid] = s.begin + s.len;
    if (tid == 0) s_b[kCtThreads] = 0;
    __syncthreads();
    s.last = (int)umin64(kCtThreads - 1, P - 1 - p0);
    const bool ok = tid >= s.last || s_b[tid + 1] == s.begin + s.len;
    s.tiled = __syncthreads_and(ok) != 0;
    s.sb = s_b[0];
    s.se = s_e[s.last];
    return s;
}

// ---- TMA bulk staging of one window -------------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_parity(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t spin = 0; !ok; ++spin) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (spin > (1u << 26)) __trap();   // never hang the device
    }
}

// Geometry of a window [w0, w0 + n) of an array of `total` elements of T whose base is 16-byte aligned: the bulk copy covers the aligned
// interior [a0, b1), element e of the array lands at buf[e - a0]; elements of the window at or beyond b1 (fewer than 16 / sizeof(T), only
// at the very end of the array) come in through ordinary loads.
template <typename T>
struct Win {
    uint64_t a0, b1;
    uint32_t bytes;
    __device__ __forceinline__ Win(uint64_t w0, uint32_t n, uint64_t total) {
        constexpr uint64_t A = 16 / sizeof(T);
        a0 = w0 / A * A;
        b1 = umin64((w0 + n + A - 1) / A * A, total / A * A);
        bytes = b1 > a0 ? (uint32_t)((b1 - a0) * sizeof(T)) : 0u;
    }
};
template <typename T>
__device__ __forceinline__ void stage_tail(T* buf, const T* __restrict__ src, const Win<T>& w, uint64_t w0, uint32_t n, int tid) {
    for (uint64_t e = umax64(w.b1, w0) + tid; e < w0 + n; e += kCtThreads) buf[e - w.a0] = src[e];
}

template <typename T> struct StArith;
template <> struct StArith<float> {
    using W = float;
    static __device__ __forceinline__ W up(float v) { return v; }
    static __device__ __forceinline__ W cast(float v) { return v; }
    static __device__ __forceinline__ W one_minus(W a) { return 1.f - a; }
    static __device__ __forceinline__ W mul(W a, W b) { return a * b; }
    static __device__ __forceinline__ float down(W v) { return v; }
};
template <> struct StArith<double> {
    using W = double;
    static __device__ __forceinline__ W up(double v) { return v; }
    static __device__ __forceinline__ W cast(float v) { return (double)v; }
    static __device__ __forceinline__ W one_minus(W a) { return (double)(1.f) - a; }
    static __device__ __forceinline__ W mul(W a, W b) { return a * b; }
    static __device__ __forceinline__ double down(W v) { return v; }
};
template <> struct StArith<__half> {   // values held as float, rounded to half after every operation the reference rounds
    using W = float;
    static __device__ __forceinline__ W up(__half v) { return __half2float(v); }
    static __device__ __forceinline__ W cast(float v) { return __half2float(__float2half_rn(v)); }
    static __device__ __forceinline__ W one_minus(W a) { return 1.f - a; }
    static __device__ __forceinline__ W mul(W a, W b) { return __half2float(__float2half_rn(a * b)); }
    static __device__ __forceinline__ __half down(W v) { return __float2half_rn(v); }
};

template <typename T> __device__ __forceinline__ T fma_t(T a, T b, T c);
template <> __device__ __forceinline__ float fma_t<float>(float a, float b, float c) { return fmaf(a, b, c); }
template <> __device__ __forceinline__ double fma_t<double>(double a, double b, double c) { return fma(a, b, c); }

// ------------------------------------------------------------------------------------------------
// packed_alpha_to_vw forward (pack_ops_cuda.cu:1735-1790)
// ------------------------------------------------------------------------------------------------
template <typename T, int WIN>
__global__ void __launch_bounds__(kCtThreads)
alpha_to_vw_fwd_cta_kernel(uint64_t P, uint64_t total, const T* __restrict__ alphas, const int64_t* __restrict__ pack_infos, float eps_, float thre_,
                           T* __restrict__ weights, int64_t* __restrict__ num_steps, uint8_t* __restrict__ selector) {
    using AR = StArith<T>;
    using W = typename AR::W;
    constexpr int A = 16 / sizeof(T);
    __shared__ __align__(128) T buf[WIN + A];          // alphas in, weights out (in place)
    __shared__ uint8_t sel_s[WIN + A];
    __shared__ uint64_t s_b[kCtThreads + 1], s_e[kCtThreads];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const uint64_t p0 = (uint64_t)blockIdx.x * kCtThreads;
    const uint32_t bar = smem_addr(&s_bar);
    if (tid == 0) mbar_init(bar);
    const CtaSpan sp = load_cta_span(pack_infos, P, p0, tid, s_b, s_e);   // (has the __syncthreads that publishes the barrier)
    const W eps = AR::cast(eps_), thre = AR::cast(thre_);
    W Tr = AR::cast(1.f);
    int cnt = 0;
    bool stopped = false;
    uint32_t parity = 0;
    const int n_spans = sp.tiled ? 1 : sp.last + 1;
    for (int s = 0; s < n_spans; ++s) {
        const uint64_t sb = sp.tiled ? sp.sb : s_b[s], se = sp.tiled ? sp.se : s_e[s];
        const bool mine = sp.tiled ? tid <= sp.last : tid == s;
        for (uint64_t w0 = sb; w0 < se; w0 += WIN) {
            const uint32_t n = (uint32_t)umin64((uint64_t)WIN, se - w0);
            const Win<T> w(w0, n, total);
            if (tid == 0 && w.bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // earlier generic-proxy accesses of buf before the async write
                mbar_expect_tx(bar, w.bytes);
                bulk_g2s(smem_addr(buf), alphas + w.a0, w.bytes, bar);
            }
            stage_tail(buf, alphas, w, w0, n, tid);
            if (w.bytes) { mbar_wait_parity(bar, parity); parity ^= 1u; }
            __syncthreads();
            if (mine) {
                // this thread's part of the window (empty -- lo >= hi -- when its pack lies before or after it)
                const uint64_t lo = umax64(sp.begin, w0), hi = umin64(sp.begin + sp.len, w0 + n);
                const uint32_t t1 = lo < hi ? (uint32_t)(hi - w.a0) : 0u;
                uint32_t t = lo < hi ? (uint32_t)(lo - w.a0) : 0u;
                auto step = [&](W a, W& wv, uint8_t& sv) {   // == the reference thread's loop body, branch-free
                    stopped = stopped || (Tr < eps);
                    const bool live = !stopped && !(a <= thre);
                    wv = live ? AR::mul(a, Tr) : AR::cast(0.f);
                    Tr = live ? AR::mul(Tr, AR::one_minus(a)) : Tr;
                    sv = live ? 1 : 0;
                    cnt += live ? 1 : 0;
                };
                for (; t + 4 <= t1; t += 4) {
                    W a[4], wv[4];
                    uint8_t sv[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) a[u] = AR::up(buf[t + u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) step(a[u], wv[u], sv[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) { buf[t + u] = AR::down(wv[u]); sel_s[t + u] = sv[u]; }
                }
                for (; t < t1; ++t) {
                    W wv;
                    uint8_t sv;
                    step(AR::up(buf[t]), wv, sv);
                    buf[t] = AR::down(wv);
                    sel_s[t] = sv;
                }
            }
            __syncthreads();
            const uint32_t off = (uint32_t)(w0 - w.a0);
            if (weights)
                for (uint32_t t = tid; t < n; t += kCtThreads) weights[w0 + t] = buf[off + t];
            if (selector)
                for (uint32_t t = tid; t < n; t += kCtThreads) selector[w0 + t] = sel_s[off + t];
            __syncthreads();
        }
    }
    if (num_steps && p0 + tid < P) num_steps[p0 + tid] = cnt;
}

// ------------------------------------------------------------------------------------------------
// packed_alpha_to_vw backward (pack_ops_cuda.cu:1792-1848), float / double.  The reference compiles with nvcc's default contraction:
//   accum += gw * w           -> fma(gw, w, accum)
//   (gw * T - accum) / ...    -> fma(gw, T, -accum) / ...
//   accum -= gw * w           -> fma(-gw, w, accum)
// written out explicitly here so that the rounding does not depend on this compiler's choices.
// ------------------------------------------------------------------------------------------------
template <typename T, int WIN>
__global__ void __launch_bounds__(kCtThreads)
alpha_to_vw_bwd_cta_kernel(uint64_t P, uint64_t total, const T* __restrict__ alphas, const T* __restrict__ weights, const T* __restrict__ grad_weights,
                           const int64_t* __restrict__ pack_infos, float eps_, float thre_, T* __restrict__ grad_alphas) {
    constexpr int A = 16 / sizeof(T);
    __shared__ __align__(128) T a_s[WIN + A];          // alphas in, dL/dalpha out (in place)
    __shared__ __align__(128) T w_s[WIN + A];
    __shared__ __align__(128) T g_s[WIN + A];
    __shared__ uint64_t s_b[kCtThreads + 1], s_e[kCtThreads];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const uint64_t p0 = (uint64_t)blockIdx.x * kCtThreads;
    const uint32_t bar = smem_addr(&s_bar);
    if (tid == 0) mbar_init(bar);
    const CtaSpan sp = load_cta_span(pack_infos, P, p0, tid, s_b, s_e);
    const T eps = (T)eps_, thre = (T)thre_;
    uint32_t parity = 0;
    const int n_spans = sp.tiled ? 1 : sp.last + 1;
    for (int s = 0; s < n_spans; ++s) {
        const uint64_t sb = sp.tiled ? sp.sb : s_b[s], se = sp.tiled ? sp.se : s_e[s];
        const bool mine = sp.tiled ? tid <= sp.last : tid == s;
        // pass 1: accum = sum_j grad_w[j] * w[j] over the whole pack, in pack order
        T accum = (T)0;
        for (uint64_t w0 = sb; w0 < se; w0 += WIN) {
            const uint32_t n = (uint32_t)umin64((uint64_t)WIN, se - w0);
            const Win<T> w(w0, n, total);
            if (tid == 0 && w.bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, 2 * w.bytes);
                bulk_g2s(smem_addr(w_s), weights + w.a0, w.bytes, bar);
                bulk_g2s(smem_addr(g_s), grad_weights + w.a0, w.bytes, bar);
            }
            stage_tail(w_s, weights, w, w0, n, tid);
            stage_tail(g_s, grad_weights, w, w0, n, tid);
            if (w.bytes) { mbar_wait_parity(bar, parity); parity ^= 1u; }
            __syncthreads();
            if (mine) {
                // this thread's part of the window (empty -- lo >= hi -- when its pack lies before or after it)
                const uint64_t lo = umax64(sp.begin, w0), hi = umin64(sp.begin + sp.len, w0 + n);
                const uint32_t t1 = lo < hi ? (uint32_t)(hi - w.a0) : 0u;
                for (uint32_t t = lo < hi ? (uint32_t)(lo - w.a0) : 0u; t < t1; ++t) accum = fma_t<T>(g_s[t], w_s[t], accum);
            }
            __syncthreads();
        }
        // pass 2: the sequential backward
        T Tr = (T)1.f;
        bool stopped = false;
        for (uint64_t w0 = sb; w0 < se; w0 += WIN) {
            const uint32_t n = (uint32_t)umin64((uint64_t)WIN, se - w0);
            const Win<T> w(w0, n, total);
            if (tid == 0 && w.bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, 3 * w.bytes);
                bulk_g2s(smem_addr(a_s), alphas + w.a0, w.bytes, bar);
                bulk_g2s(smem_addr(w_s), weights + w.a0, w.bytes, bar);
                bulk_g2s(smem_addr(g_s), grad_weights + w.a0, w.bytes, bar);
            }
            stage_tail(a_s, alphas, w, w0, n, tid);
            stage_tail(w_s, weights, w, w0, n, tid);
            stage_tail(g_s, grad_weights, w, w0, n, tid);
            if (w.bytes) { mbar_wait_parity(bar, parity); parity ^= 1u; }
            __syncthreads();
            if (mine) {
                // this thread's part of the window (empty -- lo >= hi -- when its pack lies before or after it)
                const uint64_t lo = umax64(sp.begin, w0), hi = umin64(sp.begin + sp.len, w0 + n);
                const uint32_t t1 = lo < hi ? (uint32_t)(hi - w.a0) : 0u;
                uint32_t t = lo < hi ? (uint32_t)(lo - w.a0) : 0u;
                auto step = [&](T a, T gw, T wv) -> T {
                    stopped = stopped || (Tr < eps);
                    const bool live = !stopped && !(a < thre);   // `<` here, `<=` in the forward pass (reference quirk, kept)
                    const T om = (T)(1.f) - a;
                    const T out = fma_t<T>(gw, Tr, -accum) / (T)fmaxf((float)om, 1e-10f);
                    accum = live ? fma_t<T>(-gw, wv, accum) : accum;
                    Tr = live ? Tr * om : Tr;
                    return live ? out : (T)0;
                };
                for (; t + 4 <= t1; t += 4) {
                    T a[4], gw[4], wv[4], o[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) { a[u] = a_s[t + u]; gw[u] = g_s[t + u]; wv[u] = w_s[t + u]; }
#pragma unroll
                    for (int u = 0; u < 4; ++u) o[u] = step(a[u], gw[u], wv[u]);
#pragma unroll
                    for (int u = 0; u < 4; ++u) a_s[t + u] = o[u];
                }
                for (; t < t1; ++t) a_s[t] = step(a_s[t], g_s[t], w_s[t]);
            }
            __syncthreads();
            const uint32_t off = (uint32_t)(w0 - w.a0);
            for (uint32_t t = tid; t < n; t += kCtThreads) grad_alphas[w0 + t] = a_s[off + t];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// packed_sum, one channel (pack_ops_cuda.cu:798-824): sequential sum per pack, like the reference thread
// ------------------------------------------------------------------------------------------------
template <typename T, int WIN>
__global__ void __launch_bounds__(kCtThreads)
pack_sum_cta_kernel(uint64_t P, uint64_t total, const T* __restrict__ in, const int64_t* __restrict__ pack_infos, T* __restrict__ out) {
    constexpr int A = 16 / sizeof(T);
    __shared__ __align__(128) T buf[WIN + A];
    __shared__ uint64_t s_b[kCtThreads + 1], s_e[kCtThreads];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const uint64_t p0 = (uint64_t)blockIdx.x * kCtThreads;
    const uint32_t bar = smem_addr(&s_bar);
    if (tid == 0) mbar_init(bar);
    const CtaSpan sp = load_cta_span(pack_infos, P, p0, tid, s_b, s_e);
    T acc = (T)0;
    uint32_t parity = 0;
    const int n_spans = sp.tiled ? 1 : sp.last + 1;
    for (int s = 0; s < n_spans; ++s) {
        const uint64_t sb = sp.tiled ? sp.sb : s_b[s], se = sp.tiled ? sp.se : s_e[s];
        const bool mine = sp.tiled ? tid <= sp.last : tid == s;
        for (uint64_t w0 = sb; w0 < se; w0 += WIN) {
            const uint32_t n = (uint32_t)umin64((uint64_t)WIN, se - w0);
            const Win<T> w(w0, n, total);
            if (tid == 0 && w.bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, w.bytes);
                bulk_g2s(smem_addr(buf), in + w.a0, w.bytes, bar);
            }
            stage_tail(buf, in, w, w0, n, tid);
            if (w.bytes) { mbar_wait_parity(bar, parity); parity ^= 1u; }
            __syncthreads();
            if (mine) {
                // this thread's part of the window (empty -- lo >= hi -- when its pack lies before or after it)
                const uint64_t lo = umax64(sp.begin, w0), hi = umin64(sp.begin + sp.len, w0 + n);
                const uint32_t t1 = lo < hi ? (uint32_t)(hi - w.a0) : 0u;
                for (uint32_t t = lo < hi ? (uint32_t)(lo - w.a0) : 0u; t < t1; ++t) acc += buf[t];
            }
            __syncthreads();
        }
    }
    if (p0 + tid < P) out[p0 + tid] = acc;
}

// ------------------------------------------------------------------------------------------------
// packed_cumsum / packed_cumprod, one channel (pack_ops_cuda.cu:864-1095): out[i] = in[i - offset] (op) out[i - 1] walked sequentially like
// the reference thread, forwards or backwards -- bit-identical to the reference for sums and inclusive products; the exclusive product
// follows the documented semantics (leading 1; pack_ops.cu handles the reference's all-zeros quirk before this is reached).
// ------------------------------------------------------------------------------------------------
template <typename T, bool PROD, int WIN>
__global__ void __launch_bounds__(kCtThreads)
pack_scan_cta_kernel(uint64_t P, uint64_t total, const T* __restrict__ in, const int64_t* __restrict__ pack_infos, bool exclusive, bool reverse,
                     T* __restrict__ out) {
    constexpr int A = 16 / sizeof(T);
    __shared__ __align__(128) T buf[WIN + A];
    __shared__ uint64_t s_b[kCtThreads + 1], s_e[kCtThreads];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const uint64_t p0 = (uint64_t)blockIdx.x * kCtThreads;
    const uint32_t bar = smem_addr(&s_bar);
    if (tid == 0) mbar_init(bar);
    const CtaSpan sp = load_cta_span(pack_infos, P, p0, tid, s_b, s_e);
    const T ident = PROD ? (T)1 : (T)0;
    T acc = ident, prev = ident;
    bool started = false;
    uint32_t parity = 0;
    auto step = [&](T v) -> T {
        T o;
        if (!started) { started = true; o = exclusive ? ident : v; }
        else { const T a = exclusive ? prev : v; o = PROD ? (T)(a * acc) : (T)(a + acc); }
        acc = o;
        prev = v;
        return o;
    };
    const int n_spans = sp.tiled ? 1 : sp.last + 1;
    for (int s = 0; s < n_spans; ++s) {
        const uint64_t sb = sp.tiled ? sp.sb : s_b[s], se = sp.tiled ? sp.se : s_e[s];
        const bool mine = sp.tiled ? tid <= sp.last : tid == s;
        const uint64_t n_win = (se - sb + WIN - 1) / WIN;
        for (uint64_t k = 0; k < n_win; ++k) {
            const uint64_t w0 = sb + (reverse ? n_win - 1 - k : k) * WIN;
            const uint32_t n = (uint32_t)umin64((uint64_t)WIN, se - w0);
            const Win<T> w(w0, n, total);
            if (tid == 0 && w.bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, w.bytes);
                bulk_g2s(smem_addr(buf), in + w.a0, w.bytes, bar);
            }
            stage_tail(buf, in, w, w0, n, tid);
            if (w.bytes) { mbar_wait_parity(bar, parity); parity ^= 1u; }
            __syncthreads();
            if (mine) {
                const uint64_t lo = umax64(sp.begin, w0), hi = umin64(sp.begin + sp.len, w0 + n);
                if (lo < hi) {
                    const uint32_t t0 = (uint32_t)(lo - w.a0), t1 = (uint32_t)(hi - w.a0);
                    if (!reverse) {
                        uint32_t t = t0;
                        for (; t + 4 <= t1; t += 4) {
                            T v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = buf[t + u];
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = step(v[u]);
#pragma unroll
                            for (int u = 0; u < 4; ++u) buf[t + u] = v[u];
                        }
                        for (; t < t1; ++t) buf[t] = step(buf[t]);
                    } else {
                        uint32_t t = t1;
                        for (; t >= t0 + 4; t -= 4) {
                            T v[4];
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = buf[t - 1 - u];
#pragma unroll
                            for (int u = 0; u < 4; ++u) v[u] = step(v[u]);
#pragma unroll
                            for (int u = 0; u < 4; ++u) buf[t - 1 - u] = v[u];
                        }
                        for (; t > t0; --t) buf[t - 1] = step(buf[t - 1]);
                    }
                }
            }
            __syncthreads();
            const uint32_t off = (uint32_t)(w0 - w.a0);
            for (uint32_t t = tid; t < n; t += kCtThreads) out[w0 + t] = buf[off + t];
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Two per-pack reductions of the render step in one pass (no reference export; replaces the composition
//   acc = packed_sum(w, pack_infos);  depth = packed_sum(w * t, pack_infos)         (nerf_ray_query.py:182-188: weights -> accumulated
// opacity and expected depth) and its autograd chain.  Forward: sequential sums like the reference's packed_sum thread, the product w * t
// rounded before it is added (what the separate torch `mul` kernel produces).  Backward: dL/dw[j] = fl(fl(g_depth[p] * t[j]) + g_acc[p]).
// ------------------------------------------------------------------------------------------------
template <int WIN>
__global__ void __launch_bounds__(kCtThreads)
pack_wsum_fwd_cta_kernel(uint64_t P, uint64_t total, const float* __restrict__ w_in, const float* __restrict__ t_in, const int64_t* __restrict__ pack_infos,
                         float* __restrict__ acc_out, float* __restrict__ dep_out) {
    constexpr int A = 4;
    __shared__ __align__(128) float w_s[WIN + A];
    __shared__ __align__(128) float t_s[WIN + A];
    __shared__ uint64_t s_b[kCtThreads + 1], s_e[kCtThreads];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const uint64_t p0 = (uint64_t)blockIdx.x * kCtThreads;
    const uint32_t bar = smem_addr(&s_bar);
    if (tid == 0) mbar_init(bar);
    const CtaSpan sp = load_cta_span(pack_infos, P, p0, tid, s_b, s_e);
    float acc = 0.f, dep = 0.f;
    uint32_t parity = 0;
    const int n_spans = sp.tiled ? 1 : sp.last + 1;
    for (int s = 0; s < n_spans; ++s) {
        const uint64_t sb = sp.tiled ? sp.sb : s_b[s], se = sp.tiled ? sp.se : s_e[s];
        const bool mine = sp.tiled ? tid <= sp.last : tid == s;
        for (uint64_t w0 = sb; w0 < se; w0 += WIN) {
            const uint32_t n = (uint32_t)umin64((uint64_t)WIN, se - w0);
            const Win<float> w(w0, n, total);
            if (tid == 0 && w.bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, 2 * w.bytes);
                bulk_g2s(smem_addr(w_s), w_in + w.a0, w.bytes, bar);
                bulk_g2s(smem_addr(t_s), t_in + w.a0, w.bytes, bar);
            }
            stage_tail(w_s, w_in, w, w0, n, tid);
            stage_tail(t_s, t_in, w, w0, n, tid);
            if (w.bytes) { mbar_wait_parity(bar, parity); parity ^= 1u; }
            __syncthreads();
            if (mine) {
                const uint64_t lo = umax64(sp.begin, w0), hi = umin64(sp.begin + sp.len, w0 + n);
                const uint32_t t1 = lo < hi ? (uint32_t)(hi - w.a0) : 0u;
                for (uint32_t t = lo < hi ? (uint32_t)(lo - w.a0) : 0u; t < t1; ++t) {
                    const float wv = w_s[t];
                    acc += wv;
                    dep = __fadd_rn(dep, __fmul_rn(wv, t_s[t]));
                }
            }
            __syncthreads();
        }
    }
    if (p0 + tid < P) { acc_out[p0 + tid] = acc; dep_out[p0 + tid] = dep; }
}

template <int WIN>
__global__ void __launch_bounds__(kCtThreads)
pack_wsum_bwd_cta_kernel(uint64_t P, uint64_t total, const float* __restrict__ t_in, const int64_t* __restrict__ pack_infos, const float* __restrict__ g_acc,
                         const float* __restrict__ g_dep, float* __restrict__ grad_w) {
    constexpr int A = 4;
    __shared__ __align__(128) float t_s[WIN + A];      // depths in, dL/dw out (in place)
    __shared__ uint64_t s_b[kCtThreads + 1], s_e[kCtThreads];
    __shared__ __align__(8) uint64_t s_bar;
    const int tid = threadIdx.x;
    const uint64_t p0 = (uint64_t)blockIdx.x * kCtThreads;
    const uint32_t bar = smem_addr(&s_bar);
    if (tid == 0) mbar_init(bar);
    const CtaSpan sp = load_cta_span(pack_infos, P, p0, tid, s_b, s_e);
    const float ga = (p0 + tid < P && g_acc) ? __ldg(g_acc + p0 + tid) : 0.f, gd = (p0 + tid < P && g_dep) ? __ldg(g_dep + p0 + tid) : 0.f;
    uint32_t parity = 0;
    const int n_spans = sp.tiled ? 1 : sp.last + 1;
    for (int s = 0; s < n_spans; ++s) {
        const uint64_t sb = sp.tiled ? sp.sb : s_b[s], se = sp.tiled ? sp.se : s_e[s];
        const bool mine = sp.tiled ? tid <= sp.last : tid == s;
        for (uint64_t w0 = sb; w0 < se; w0 += WIN) {
            const uint32_t n = (uint32_t)umin64((uint64_t)WIN, se - w0);
            const Win<float> w(w0, n, total);
            if (tid == 0 && w.bytes) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                mbar_expect_tx(bar, w.bytes);
                bulk_g2s(smem_addr(t_s), t_in + w.a0, w.bytes, bar);
            }
            stage_tail(t_s, t_in, w, w0, n, tid);
            if (w.bytes) { mbar_wait_parity(bar, parity); parity ^= 1u; }
            __syncthreads();
            if (mine) {
                const uint64_t lo = umax64(sp.begin, w0), hi = umin64(sp.begin + sp.len, w0 + n);
                const uint32_t t1 = lo < hi ? (uint32_t)(hi - w.a0) : 0u;
                for (uint32_t t = lo < hi ? (uint32_t)(lo - w.a0) : 0u; t < t1; ++t) t_s[t] = __fadd_rn(__fmul_rn(gd, t_s[t]), ga);
            }
            __syncthreads();
            const uint32_t off = (uint32_t)(w0 - w.a0);
            for (uint32_t t = tid; t < n; t += kCtThreads) grad_w[w0 + t] = t_s[off + t];
            __syncthreads();
        }
    }
}

static inline unsigned ct_grid(uint64_t P) { return (unsigned)div_up<uint64_t>(P, (uint64_t)kCtThreads); }
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Launchers used by the C-ABI entry points in pack_ops.cu.  `total` = number of elements of the sample arrays (the bulk copies never read
// past it).  Return 1 when the call is not served here (dtype without a staged kernel, or a base pointer that is not 16-byte aligned):
// pack_ops.cu's warp-per-pack kernels take it.
int staged_alpha_fwd(int32_t dtype, uint64_t P, uint64_t total, const void* alphas, const int64_t* pack_infos, float eps, float thre, void* weights,
                     int64_t* num_steps, uint8_t* selector, cudaStream_t st) {
    if (!aligned16(alphas)) return 1;
    switch (dtype) {
    case NR3D_F32: alpha_to_vw_fwd_cta_kernel<float, 4096><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const float*)alphas, pack_infos, eps, thre, (float*)weights, num_steps, selector); return 0;
    case NR3D_F64: alpha_to_vw_fwd_cta_kernel<double, 2048><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const double*)alphas, pack_infos, eps, thre, (double*)weights, num_steps, selector); return 0;
    case NR3D_F16: alpha_to_vw_fwd_cta_kernel<__half, 4096><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const __half*)alphas, pack_infos, eps, thre, (__half*)weights, num_steps, selector); return 0;
    default: return 1;
    }
}

int staged_alpha_bwd(int32_t dtype, uint64_t P, uint64_t total, const void* alphas, const void* weights, const void* grad_weights, const int64_t* pack_infos,
                     float eps, float thre, void* grad_alphas, cudaStream_t st) {
    if (!aligned16(alphas) || !aligned16(weights) || !aligned16(grad_weights)) return 1;
    switch (dtype) {
    case NR3D_F32: alpha_to_vw_bwd_cta_kernel<float, 2048><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const float*)alphas, (const float*)weights, (const float*)grad_weights, pack_infos, eps, thre, (float*)grad_alphas); return 0;
    case NR3D_F64: alpha_to_vw_bwd_cta_kernel<double, 1024><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const double*)alphas, (const double*)weights, (const double*)grad_weights, pack_infos, eps, thre, (double*)grad_alphas); return 0;
    default: return 1;
    }
}

int staged_pack_scan(int32_t dtype, bool prod, uint64_t P, uint64_t total, const void* in, const int64_t* pack_infos, bool exclusive, bool reverse, void* out,
                     cudaStream_t st) {
    if (!aligned16(in)) return 1;
    if (dtype == NR3D_F32) {
        if (prod) pack_scan_cta_kernel<float, true, 4096><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const float*)in, pack_infos, exclusive, reverse, (float*)out);
        else pack_scan_cta_kernel<float, false, 4096><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const float*)in, pack_infos, exclusive, reverse, (float*)out);
        return 0;
    }
    if (dtype == NR3D_F64) {
        if (prod) pack_scan_cta_kernel<double, true, 2048><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const double*)in, pack_infos, exclusive, reverse, (double*)out);
        else pack_scan_cta_kernel<double, false, 2048><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const double*)in, pack_infos, exclusive, reverse, (double*)out);
        return 0;
    }
    return 1;
}

int staged_pack_sum(int32_t dtype, uint64_t P, uint64_t total, const void* in, const int64_t* pack_infos, void* out, cudaStream_t st) {
    if (!aligned16(in)) return 1;
    switch (dtype) {
    case NR3D_F32: pack_sum_cta_kernel<float, 4096><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const float*)in, pack_infos, (float*)out); return 0;
    case NR3D_F64: pack_sum_cta_kernel<double, 2048><<<ct_grid(P), kCtThreads, 0, st>>>(P, total, (const double*)in, pack_infos, (double*)out); return 0;
    default: return 1;
    }
}

}  // namespace nr3d

using namespace nr3d;

extern "C" {

int nr3d_pack_weighted_sums_fwd(uint64_t P, uint64_t S, const float* weights, const float* depths, const int64_t* pack_infos, float* acc, float* depth,
                                void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(weights && depths && pack_infos && acc && depth, "packed_weighted_sums: null argument");
    NR3D_CHECK(aligned16(weights) && aligned16(depths), "packed_weighted_sums: weights / depths must be 16-byte aligned");
    pack_wsum_fwd_cta_kernel<4096><<<ct_grid(P), kCtThreads, 0, (cudaStream_t)stream>>>(P, S, weights, depths, pack_infos, acc, depth);
    NR3D_LAUNCH_CHECK("packed_weighted_sums");
    return 0;
}

int nr3d_pack_weighted_sums_bwd(uint64_t P, uint64_t S, const float* depths, const int64_t* pack_infos, const float* grad_acc, const float* grad_depth,
                                float* grad_weights, void* stream) {
    if (P == 0) return 0;
    NR3D_CHECK(depths && pack_infos && grad_weights && (grad_acc || grad_depth), "packed_weighted_sums backward: null argument");
    NR3D_CHECK(aligned16(depths), "packed_weighted_sums backward: depths must be 16-byte aligned");
    pack_wsum_bwd_cta_kernel<4096><<<ct_grid(P), kCtThreads, 0, (cudaStream_t)stream>>>(P, S, depths, pack_infos, grad_acc, grad_depth, grad_weights);
    NR3D_LAUNCH_CHECK("packed_weighted_sums backward");
    return 0;
}

}  // extern "C"
