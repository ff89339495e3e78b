// Shared helpers for the gg_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gg_b200.h"

namespace gg {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

inline cudaStream_t as_stream(gg_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }
inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Carves aligned sub-buffers out of a caller-provided workspace.
struct Carver {
    char* base;
    size_t off = 0;
    explicit Carver(void* p) : base(static_cast<char*>(p)) {}
    template <typename T>
    T* take(size_t n) {
        off = align_up(off, 256);
        T* p = reinterpret_cast<T*>(base + off);
        off += n * sizeof(T);
        return p;
    }
    size_t used() const { return align_up(off, 256); }
};

}  // namespace gg

#define GG_REQUIRE(cond, ...)           \
    do {                                \
        if (!(cond)) {                  \
            gg::set_error(__VA_ARGS__); \
            return GG_ERR_INVALID;      \
        }                               \
    } while (0)

#define GG_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t _e = (expr);                                                           \
        if (_e != cudaSuccess) {                                                           \
            gg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, \
                          __LINE__);                                                       \
            return GG_ERR_CUDA;                                                            \
        }                                                                                  \
    } while (0)

// Check the launch that was just issued and count it.
#define GG_LAUNCHED()                      \
    do {                                   \
        gg::count_launch();                \
        GG_CUDA(cudaPeekAtLastError());    \
    } while (0)

// 128-bit read-only gather that does not pollute L1 (feature rows are touched once per edge by
// this SM; reuse across SMs is served by the 126 MB L2).
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
// ---- L2 residency control ------------------------------------------------------------------------
// The aggregation gathers feature rows at random (reused across rows, worth keeping in the 126 MB L2) while
// it streams neighbour ids, weights and output rows through once.  Without hints the streams evict the
// rows; with them the gathers are tagged evict_last and the streams evict_first.
__device__ __forceinline__ uint64_t l2_policy(int kind) {  // 0 normal, 1 evict_last, 2 evict_first
    uint64_t p;
    if (kind == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    else if (kind == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    else asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ldg_nc_f4_hint(const float4* p, uint64_t pol) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ int32_t ldg_nc_s32_hint(const int32_t* p, uint64_t pol) {
    int32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.s32 %0, [%1], %2;" : "=r"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ float ldg_nc_f32_hint(const float* p, uint64_t pol) {
    float r;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(r) : "l"(p), "l"(pol));
    return r;
}
__device__ __forceinline__ void stg_f4_hint(float4* p, const float4& v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
                 "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void fma4(float4& acc, float w, const float4& v) {
    acc.x = fmaf(w, v.x, acc.x);
    acc.y = fmaf(w, v.y, acc.y);
    acc.z = fmaf(w, v.z, acc.z);
    acc.w = fmaf(w, v.w, acc.w);
}
__device__ __forceinline__ void add4(float4& acc, const float4& v) {
    acc.x += v.x;
    acc.y += v.y;
    acc.z += v.z;
    acc.w += v.w;
}
