// Input-space formulation of the GAT layer (first layer: no gradient w.r.t. x, concat = False, H = 8, C = 64).
//
// PyG's GATConv projects first (xw = x W^T, [N, H*C]) and aggregates 2 KB projected rows per edge.  When the input
// width K is smaller than H*C (166 vs 512 for the reference's first layer, src/models/gat.py:39) the same layer is
//     Z[i,h,:]  = sum_{j->i} alpha[e,h] * x[j,:]                       (per-edge gather: K*4 bytes instead of 2 KB)
//     out[i,c]  = (1/H) sum_{h,k} Z[i,h,k] * W[h*C+c,k] + bias[c]      (one [n, H*K] x [H*K, C] GEMM on the tensor cores)
//     a_src[n,h] = x[n,:] . (W_h^T att_src[h]),  a_dst likewise       (logits straight from x)
// and the backward needs no projected features either:
//     Gd[i,h,:] = (1/H) W_h^T dO[i]                                    ([n, C] x [C, H*K] GEMM)
//     d_alpha[e,h] = Gd[i,h,:] . x[j,:]      -> softmax / LeakyReLU backward -> dz, da_dst, da_src = sum_e dz
//     dW[h*C+c,k] = (1/H) sum_i dO[i,c] Z[i,h,k] + att_src[h,c] (da_src^T x)[h,k] + att_dst[h,c] (da_dst^T x)[h,k]
// Across GPUs (destination-range partition, x replicated) this removes the redundant projection of every referenced
// source row and the exchange of per-edge gradients: only [N,H] logits / logit gradients cross NVLink.
//
// Z is written ONCE, by the aggregation kernel, directly in the layout the tensor cores read: per 128-row tile and per
// 64-feature k-block two 16 KB planes (fp16 "hi" and fp16 "lo" of the power-of-two scaled value, hi + lo carries 22
// mantissa bits), each plane the canonical 128B-swizzled shared-memory image, so the GEMMs fetch a k-block with ONE
// cp.async.bulk and need no operand-staging warps.  The same image is a K-major operand (rows = nodes, out = Z W_r) and
// an MN-major operand (rows = features, dW = Z^T dO): the swizzle atom (8 rows x 128 B) is identical for both.
#pragma once
#include <cuda_fp16.h>

#include "gat_stream.cuh"

namespace gnnfd {
namespace in {

constexpr int H = 8, C = 64;
constexpr int NSLOT = 3;                 // a lane holds features 64*r + 2*lane, +1 (r < NSLOT)  =>  K <= 192
constexpr int MAX_K = 64 * NSLOT;
constexpr int TILE = 128;                // rows (nodes) per image tile
constexpr int PLANE = TILE * 128;        // 16 KB: one fp16 plane of one k-block (64 features) of one tile
constexpr int KBLOCK = 2 * PLANE;        // hi plane then lo plane

struct GI { static constexpr int H = in::H, C = in::C; };   // what the shared phase-A code needs

struct Dims {
    int K, KP, F, NKB;
    __host__ __device__ explicit Dims(int k) : K(k), KP((k + 7) & ~7), F(H * ((k + 7) & ~7)), NKB((k + 7) >> 3) {}
};

// byte offset of (row r of the tile, element e of the k-block) inside one 128B-swizzled plane
__host__ __device__ __forceinline__ uint32_t plane_off(int r, int e)
{
    return uint32_t(r >> 3) * 1024u + uint32_t(r & 7) * 128u + ((uint32_t(e >> 3) ^ uint32_t(r & 7)) << 4) + uint32_t(e & 7) * 2u;
}

// power-of-two scale s with m*s in [2^11, 2^12): fp16 keeps 11 bits, leaves x16 headroom below 65504 (attention dropout
// rescales by 1/(1-p)) and 26 bits of range above the smallest normal
__host__ __device__ __forceinline__ float pow2_scale(float m)
{
#ifdef __CUDA_ARCH__
    const uint32_t b = __float_as_uint(m);
#else
    uint32_t b;
    memcpy(&b, &m, 4);
#endif
    const int e = int((b >> 23) & 0xffu) - 127;
    if (!(m > 0.f) || e < -100 || e > 100) return 1.f;
    const uint32_t sb = uint32_t(127 + 11 - e) << 23;
#ifdef __CUDA_ARCH__
    return __uint_as_float(sb);
#else
    float s;
    memcpy(&s, &sb, 4);
    return s;
#endif
}

// v (already scaled) -> fp16 hi + fp16 lo, hi + lo = v to 22 bits (the fp32 residual is exact)
__device__ __forceinline__ void split_h2(float v0, float v1, __half2& hi, __half2& lo)
{
    hi = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(hi);
    lo = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
}

// Device-resident per-call constants and operand images ("prep" buffer), built by gnnfd_in_prepare:
//   float scal[8]      : [0] sx (scale of x / Z), [1] sw (scale of W), [2] 1/(sx*sw*H), [3] 1/sx
//   W image for out = Z W_r   : NKB k-blocks x (hi [64 rows x 128 B] | lo), rows = c, K-major, scaled by sw
//   W image for Gd = dO W_r^T : tf32 hi/lo images of gemm_tc_ws (rows = feature f, reduction over c), scaled by 1/H
//   u [2H][KP]         : W_h^T att_src[h] (rows 0..H-1) and W_h^T att_dst[h] (rows H..2H-1)
constexpr size_t PREP_SCAL_BYTES = 256;
__host__ __device__ inline size_t prep_wout_bytes(const Dims& d) { return size_t(d.NKB) * 16384; }
__host__ __device__ inline size_t prep_wgd_bytes(const Dims& d)
{
    return size_t((d.F + 255) / 256) * 2 /*k-blocks of 32 c*/ * 2 /*hi, lo*/ * (256 * 32) * sizeof(float);
}
__host__ __device__ inline size_t prep_u_bytes(const Dims& d) { return size_t(2 * H) * d.KP * sizeof(float); }
__host__ __device__ inline size_t prep_off_wout(const Dims&) { return 1024; }
__host__ __device__ inline size_t prep_off_wgd(const Dims& d) { return 1024 + ((prep_wout_bytes(d) + 1023) & ~size_t(1023)); }
__host__ __device__ inline size_t prep_off_u(const Dims& d) { return prep_off_wgd(d) + ((prep_wgd_bytes(d) + 1023) & ~size_t(1023)); }
__host__ __device__ inline size_t prep_bytes(const Dims& d) { return prep_off_u(d) + ((prep_u_bytes(d) + 1023) & ~size_t(1023)) + 1024; }

inline size_t zimg_bytes(int64_t n_rows, const Dims& d) { return size_t((n_rows + TILE - 1) / TILE) * d.NKB * KBLOCK; }

}  // namespace in
}  // namespace gnnfd

namespace gnnfd {
namespace in {

// ---- per-warp ring of staged x rows ------------------------------------------------------------------------------
// Row j starts at x + j*ldx, which is only 4- or 8-byte aligned (K = 166: 664-byte stride).  cp.async.bulk needs
// 16-byte aligned source and size, so lane 0 fetches the enclosing 16-byte aligned span into the slot and the consumer
// adds (address & 15).  Sizes are run-time (they depend on K); per-warp layout:
//   [R slots][p_s: 2 x 32 x H floats][j_s: 2 x 32 ints][R+1 mbarriers][extra]
constexpr int IN_WARPS = 4;
constexpr int IN_THREADS = IN_WARPS * 32;
constexpr int IN_R = 10;                                   // ring slots per warp
constexpr int IN_P_BYTES = 2 * 32 * H * 4, IN_J_BYTES = 2 * 32 * 4, IN_BAR_BYTES = ((IN_R + 1) * 8 + 15) / 16 * 16;

__host__ __device__ inline int in_slot_bytes(int K) { return (K * 4 + 12 + 15) / 16 * 16; }
__host__ __device__ inline int in_warp_bytes(int K, int extra)
{
    return (IN_R * in_slot_bytes(K) + IN_P_BYTES + IN_J_BYTES + IN_BAR_BYTES + extra + 127) / 128 * 128;
}

struct InRing {
    uint8_t* ring;
    uint8_t* extra;
    float* p_s;      // [2][32*H]
    int* j_s;        // [2][32]
    uint64_t* full;  // [IN_R] + one spare barrier (index IN_R) for kernel-specific use
    int slot, issued = 0, consumed = 0;
    uint64_t xaddr;
    uint32_t ldxb, rowb;

    __device__ __forceinline__ void init(uint8_t* base, const float* x, int64_t ldx, int K, int lane)
    {
        slot = in_slot_bytes(K);
        ring = base;
        p_s = reinterpret_cast<float*>(base + IN_R * slot);
        j_s = reinterpret_cast<int*>(base + IN_R * slot + IN_P_BYTES);
        full = reinterpret_cast<uint64_t*>(base + IN_R * slot + IN_P_BYTES + IN_J_BYTES);
        extra = base + IN_R * slot + IN_P_BYTES + IN_J_BYTES + IN_BAR_BYTES;
        xaddr = reinterpret_cast<uint64_t>(x);
        ldxb = uint32_t(ldx) * 4u;
        rowb = uint32_t(K) * 4u;
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i <= IN_R; ++i) st_mbar_init(&full[i], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }
    __device__ __forceinline__ bool has_room() const { return issued - consumed < IN_R; }
    __device__ __forceinline__ void issue(int j, int lane)
    {
        if (lane == 0) {
            const int s = issued % IN_R;
            const uint64_t a = xaddr + uint64_t(uint32_t(j)) * ldxb;
            const uint64_t a0 = a & ~uint64_t(15);
            const uint32_t bytes = uint32_t(((a + rowb + 15) & ~uint64_t(15)) - a0);
            st_mbar_expect_tx(&full[s], bytes);
            st_bulk_g2s(ring + s * slot, reinterpret_cast<const void*>(a0), bytes, &full[s]);
        }
        ++issued;
    }
    // wait for the oldest in-flight row (source id j); returns the address of its first element
    __device__ __forceinline__ const float* front(int j)
    {
        const int s = consumed % IN_R;
        st_mbar_wait(&full[s], (consumed / IN_R) & 1);
        const uint32_t off = uint32_t(xaddr + uint64_t(uint32_t(j)) * ldxb) & 15u;
        return reinterpret_cast<const float*>(ring + s * slot + off);
    }
    __device__ __forceinline__ void pop()
    {
        __syncwarp();
        ++consumed;
    }
};

// the features of one staged row owned by this lane: pairs (64r + 2*lane, +1), zero beyond K
template <bool VEC2>
__device__ __forceinline__ void load_xrow(const float* __restrict__ row, int lane, int K, float2 (&v)[NSLOT])
{
#pragma unroll
    for (int r = 0; r < NSLOT; ++r) {
        const int f = 64 * r + 2 * lane;
        if (VEC2) {   // K even, rows 8-byte aligned: a pair is inside the row or outside
            v[r] = (f < K) ? *reinterpret_cast<const float2*>(row + f) : make_float2(0.f, 0.f);
        } else {
            v[r].x = (f < K) ? row[f] : 0.f;
            v[r].y = (f + 1 < K) ? row[f + 1] : 0.f;
        }
    }
}

}  // namespace in
}  // namespace gnnfd
