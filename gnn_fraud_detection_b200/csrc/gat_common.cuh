// Lane/slot geometry shared by the forward and backward message-passing kernels.
//
// A projected feature row has D = H*C elements.  One warp owns one row; a warp-wide 128-bit load covers
// 32*VW consecutive elements (VW = 4 fp32 / 8 bf16 per lane), so a row is NS = D/(32*VW) "slots" per
// lane, each slot a fully coalesced 512-byte warp transaction.  All VW elements of a slot lie in one
// head; G = C/VW consecutive lanes share a head, HP = 32/G heads are covered by one slot instruction,
// and the head of (slot q, lane l) is q*HP + l/G.
#pragma once
#include "common.cuh"

namespace gnnfd {

template <int H_, int C_, typename XT_>
struct Geo {
    using XT = XT_;
    static constexpr int H = H_, C = C_, D = H_ * C_;
    static constexpr int VW = sizeof(XT_) == 4 ? 4 : 8;
    static constexpr int NS = D / (32 * VW);
    static constexpr int G = C / VW;
    static constexpr int HP = 32 / G;
    static_assert(D % (32 * VW) == 0, "row must be a whole number of warp-wide 128-bit loads");
    static_assert(C % VW == 0 && G >= 1 && G <= 32 && (G & (G - 1)) == 0, "C/VW must be a power of two <= 32");
    static_assert(NS * HP == H, "slots x heads-per-slot must tile the heads");
    static_assert(H % 4 == 0 && H <= 8, "logit vectors are loaded as float4s");
};

// arr[q*HP + sub] with compile-time q and run-time sub, without dynamic register indexing
template <int HP, int N>
__device__ __forceinline__ float pick(const float (&arr)[N], int q, int sub)
{
    float r = arr[q * HP];
#pragma unroll
    for (int k = 1; k < HP; ++k) r = (sub == k) ? arr[q * HP + k] : r;
    return r;
}

template <int H>
__device__ __forceinline__ void load_vecH(const float* __restrict__ p, float (&v)[H])
{
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p) + k);
        v[4 * k + 0] = t.x; v[4 * k + 1] = t.y; v[4 * k + 2] = t.z; v[4 * k + 3] = t.w;
    }
}
template <int H>
__device__ __forceinline__ void store_vecH(float* __restrict__ p, const float (&v)[H])
{
#pragma unroll
    for (int k = 0; k < H / 4; ++k)
        reinterpret_cast<float4*>(p)[k] = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
}

// one slot (VW elements) of a feature row, streamed past L1
__device__ __forceinline__ void load_slot(const float* __restrict__ row, int q, int lane, float (&v)[4])
{
    const float4 t = ldg_stream(reinterpret_cast<const float4*>(row) + lane + 32 * q);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void load_slot(const __nv_bfloat16* __restrict__ row, int q, int lane, float (&v)[8])
{
    const uint4 t = ldg_stream_u4(reinterpret_cast<const uint4*>(row) + lane + 32 * q);
    v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
    v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
    v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
    v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
}

constexpr int ROW_WARPS = 8;                 // warps (rows) per CTA
constexpr int ROW_THREADS = ROW_WARPS * 32;

}  // namespace gnnfd
