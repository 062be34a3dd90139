// common.cuh -- shared helpers of libnr3d_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/nr3d_b200.h"

namespace nr3d {

// ---------------------------------------------------------------------------------------------
// error reporting: thread-local message + negative status (see include/nr3d_b200.h)
// ---------------------------------------------------------------------------------------------
char* tls_error_buffer();
int fail(const char* fmt, ...);
extern std::atomic<uint64_t> g_launch_count;
inline void count_launch(uint64_t n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

#define NR3D_CHECK(cond, ...) do { if (!(cond)) return ::nr3d::fail(__VA_ARGS__); } while (0)
#define NR3D_LAUNCH_CHECK(name) do { cudaError_t e__ = cudaPeekAtLastError(); \
    if (e__ != cudaSuccess) { (void)cudaGetLastError(); return ::nr3d::fail("%s: launch failed: %s", name, cudaGetErrorString(e__)); } \
    ::nr3d::count_launch(); } while (0)

constexpr int kSMs = 148;  // B200

template <typename T> __host__ __device__ inline T div_up(T a, T b) { return (a + b - 1) / b; }

// ---------------------------------------------------------------------------------------------
// vector reductions to global memory (sm_90+): one L2 atomic op for 2 / 4 packed fp32 values
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void red_add_v2_f32(float* addr, float a, float b) {
    asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ void red_add_h2(__half* addr, __half2 v) {
    uint32_t u = *reinterpret_cast<uint32_t*>(&v);
    asm volatile("red.global.add.noftz.f16x2 [%0], %1;" ::"l"(addr), "r"(u) : "memory");
}

// streaming (evict-first) stores for write-once outputs so they do not displace the parameter tables in L2
__device__ __forceinline__ void st_cs(float* p, float v) { __stcs(p, v); }
__device__ __forceinline__ void st_cs(__half* p, __half v) {
    asm volatile("st.global.cs.b16 [%0], %1;" ::"l"(p), "h"(*reinterpret_cast<unsigned short*>(&v)) : "memory");
}

// predicated forms (one instruction, no branch / reconvergence bookkeeping around a store that only some lanes of a warp perform)
__device__ __forceinline__ void st_cs_if(bool on, float* p, float v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.f32 [%0], %1;\n\t}" ::"l"(p), "f"(v), "r"((int)on) : "memory");
}
__device__ __forceinline__ void st_cs_if(bool on, __half* p, __half v) {
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.b32 q, %2, 0;\n\t@q st.global.cs.b16 [%0], %1;\n\t}" ::"l"(p), "h"(*reinterpret_cast<unsigned short*>(&v)), "r"((int)on) : "memory");
}

template <typename T> struct Cvt;
template <> struct Cvt<float> {
    static __device__ __forceinline__ float to_f(float v) { return v; }
    static __device__ __forceinline__ float from_f(float v) { return v; }
    static __device__ __forceinline__ float add(float a, float b) { return a + b; }
    static __device__ __forceinline__ float zero() { return 0.f; }
};
template <> struct Cvt<__half> {
    static __device__ __forceinline__ float to_f(__half v) { return __half2float(v); }
    static __device__ __forceinline__ __half from_f(float v) { return __float2half_rn(v); }
    static __device__ __forceinline__ __half add(__half a, __half b) { return __hadd(a, b); }
    static __device__ __forceinline__ __half zero() { return __float2half_rn(0.f); }
};

}  // namespace nr3d
