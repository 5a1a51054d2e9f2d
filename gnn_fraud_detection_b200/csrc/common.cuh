// Shared device/host helpers for libgnnfd_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/gnnfd_b200.h"

namespace gnnfd {

// ---------------------------------------------------------------------------------------------
// error plumbing: every extern "C" entry returns an int code; the message is thread-local.
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);

#define GNNFD_REQUIRE(cond, code, ...)                 \
    do {                                               \
        if (!(cond)) {                                 \
            ::gnnfd::set_error(__VA_ARGS__);           \
            return (code);                             \
        }                                              \
    } while (0)

#define GNNFD_CUDA(expr)                                                                      \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            ::gnnfd::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),        \
                               __FILE__, __LINE__);                                           \
            return GNNFD_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

#define GNNFD_LAUNCH_CHECK()                                                                  \
    do {                                                                                      \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess) {                                                              \
            ::gnnfd::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e),    \
                               __FILE__, __LINE__);                                           \
            return GNNFD_ERR_CUDA;                                                            \
        }                                                                                     \
    } while (0)

inline int sm_count()
{
    static int cached = 0;
    if (!cached) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&cached, cudaDevAttrMultiProcessorCount, dev);
        if (cached <= 0) cached = 148;
    }
    return cached;
}

template <typename T>
inline T* carve(char*& p, size_t count)
{
    uintptr_t a = (reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255);
    T* r = reinterpret_cast<T*>(a);
    p = reinterpret_cast<char*>(a + count * sizeof(T));
    return r;
}
inline size_t carve_bytes(size_t count, size_t elem) { return ((count * elem + 255) & ~size_t(255)) + 256; }

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL, v, o));
    return v;
}
__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

// 128-bit read-only loads.  Feature rows are gathered at random: keep them out of L1 (no reuse inside
// an SM), let L2 decide.  Small per-node vectors (logits) go through the normal read-only path.
__device__ __forceinline__ float4 ldg_stream(const float4* p)
{
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
// software prefetch of one 128-byte line into L2 (no register, no scoreboard)
__device__ __forceinline__ void prefetch_l2(const void* p)
{
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
__device__ __forceinline__ uint4 ldg_stream_u4(const uint4* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void stg_stream(float4* p, float4 v)
{
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w)
                 : "memory");
}

__device__ __forceinline__ float leaky(float z, float slope) { return z > 0.f ? z : z * slope; }

}  // namespace gnnfd
