// Input-space forward of the first GAT layer (see in_common.cuh for the algebra):
//   gnnfd_in_logits   a_src / a_dst straight from x (u = W_h^T att precomputed), + max |x| for the fp16-pair scale
//   gnnfd_in_prepare  per-call constants: scales, the W images of the two dense stages
//   gnnfd_in_fwd      LeakyReLU + online segment softmax + alpha-weighted gather-sum of INPUT rows x[j] (K*4 bytes per
//                     edge instead of the 2 KB projected row), one accumulator set per head, written as the fp16-pair
//                     tensor-core image Z that gnnfd_in_out (in_gemm.cu) multiplies with W.
// Replaces, for the reference's first layer (src/models/gat.py:39,80; tgn.py:43,94), the same PyG stages as
// gnnfd_project_fwd + gnnfd_gat_fwd: lin_src / (x*att).sum(-1) / edge_update / message / aggregate.
// Same warp-stream structure as gat_fwd.cu: one warp per edge-balanced work item, phase A (lane = edge) one chunk
// ahead, rows delivered by the bulk-copy engine into a per-warp ring; packs of short rows; hub rows split into chunks
// merged in chunk order (deterministic).
#include "in_common.cuh"
#include "gat_phase_fwd.cuh"

#include <atomic>
#include <climits>
#include <cstdlib>

namespace gnnfd {
extern std::atomic<long long> g_launches;
int check_graph(const gnnfd_graph_t* g, bool need_csc, const char* who);
int in_build_gd_image(const float* W, int K, void* prep, cudaStream_t st);   // project_tc.cu

namespace in {

// ---- u = W_h^T att  ([2H][KP]) --------------------------------------------------------------------------------------
__global__ void in_u_kernel(const float* __restrict__ W, const float* __restrict__ att_src, const float* __restrict__ att_dst,
                            int K, int KP, float* __restrict__ u)
{
    const int o = blockIdx.x;                 // 0..2H-1
    const int h = o % H;
    const float* att = (o < H ? att_src : att_dst) + h * C;
    for (int k = threadIdx.x; k < KP; k += blockDim.x) {
        float s = 0.f;
        if (k < K)
            for (int c = 0; c < C; ++c) s = fmaf(W[int64_t(h * C + c) * K + k], att[c], s);
        u[o * KP + k] = s;
    }
}

// 16 per-lane values -> the full warp sums, value o ends up in the lanes with (lane >> 1) == o (both lanes of the pair)
__device__ __forceinline__ float reduce16(float (&d)[16], int lane)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4, b1 = lane & 2;
    float a[8], b[4], c[2];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = (b4 ? d[8 + i] : d[i]) + __shfl_xor_sync(FULL, b4 ? d[i] : d[8 + i], 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = (b3 ? a[4 + i] : a[i]) + __shfl_xor_sync(FULL, b3 ? a[i] : a[4 + i], 8);
#pragma unroll
    for (int i = 0; i < 2; ++i) c[i] = (b2 ? b[2 + i] : b[i]) + __shfl_xor_sync(FULL, b2 ? b[i] : b[2 + i], 4);
    float r = (b1 ? c[1] : c[0]) + __shfl_xor_sync(FULL, b1 ? c[0] : c[1], 2);
    r += __shfl_xor_sync(FULL, r, 1);
    return r;
}

// a_src[n,h] = x[n,:] . u[h,:], a_dst[n,h] = x[n,:] . u[H+h,:]; warp per row, lanes over features
template <bool VEC2>
__global__ void __launch_bounds__(256, 2)
in_logits_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int K, const float* __restrict__ u, int KP,
                 float* __restrict__ a_src, float* __restrict__ a_dst, unsigned* __restrict__ xmax_bits)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float2 uu[2 * H][NSLOT];
#pragma unroll
    for (int o = 0; o < 2 * H; ++o)
#pragma unroll
        for (int r = 0; r < NSLOT; ++r) {
            const int f = 64 * r + 2 * lane;
            uu[o][r].x = (f < K) ? u[o * KP + f] : 0.f;
            uu[o][r].y = (f + 1 < K) ? u[o * KP + f + 1] : 0.f;
        }
    float mx = 0.f;
    for (int64_t n = int64_t(blockIdx.x) * 8 + warp; n < N; n += int64_t(gridDim.x) * 8) {
        float2 v[NSLOT];
        load_xrow<VEC2>(x + n * ldx, lane, K, v);
        float d[2 * H];
#pragma unroll
        for (int o = 0; o < 2 * H; ++o) {
            float s = 0.f;
#pragma unroll
            for (int r = 0; r < NSLOT; ++r) s = fmaf(uu[o][r].x, v[r].x, fmaf(uu[o][r].y, v[r].y, s));
            d[o] = s;
        }
#pragma unroll
        for (int r = 0; r < NSLOT; ++r) mx = fmaxf(mx, fmaxf(fabsf(v[r].x), fabsf(v[r].y)));
        const float r = reduce16(d, lane);
        if ((lane & 1) == 0) {
            const int o = lane >> 1;
            if (o < H) a_src[n * H + o] = r;
            else a_dst[n * H + (o - H)] = r;
        }
    }
    mx = warp_max(mx);
    if (lane == 0 && mx > 0.f) atomicMax(xmax_bits, __float_as_uint(mx));   // non-negative floats order like their bits
}

// scal[0] = sx, [1] = sw, [2] = 1/(sx*sw*H), [3] = 1/sx;  one block
__global__ void in_scales_kernel(const float* __restrict__ W, int64_t n_w, const float* __restrict__ xmax, float* __restrict__ scal)
{
    __shared__ float red[32];
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < n_w; i += blockDim.x) m = fmaxf(m, fabsf(W[i]));
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x < 32) {
        m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.f;
        m = warp_max(m);
        if (threadIdx.x == 0) {
            const float sx = pow2_scale(*xmax), sw = pow2_scale(m);
            scal[0] = sx;
            scal[1] = sw;
            scal[2] = 1.f / (sx * sw * float(H));
            scal[3] = 1.f / sx;
        }
    }
}

// W image of out = Z W_r: k-block kb = features [64kb, 64kb+64), rows = c; element (c, f=(h,k)) = W[h*C+c, k] * sw
__global__ void in_wout_image_kernel(const float* __restrict__ W, int K, int KP, int NKB, const float* __restrict__ scal,
                                     uint8_t* __restrict__ img)
{
    const int kb = blockIdx.x;
    const float sw = scal[1];
    for (int idx = threadIdx.x; idx < C * 64; idx += blockDim.x) {
        const int c = idx >> 6, e = idx & 63;
        const int f = kb * 64 + e, h = f / KP, k = f - h * KP;
        float v = 0.f;
        if (h < H && k < K) v = W[int64_t(h * C + c) * K + k] * sw;
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const uint32_t off = plane_off(c, e);
        *reinterpret_cast<__half*>(img + size_t(kb) * 16384 + off) = hi;
        *reinterpret_cast<__half*>(img + size_t(kb) * 16384 + 8192 + off) = lo;
    }
}

// ---- sinks: what happens when a row (or a hub chunk) is complete --------------------------------------------------------
struct ZSink {      // normalise, save the row statistics, write the row of the fp16-pair image
    uint8_t* zimg; float* rowmax; float* rowsum; const float* scal; int KP, NKB;
    __device__ __forceinline__ void write_row(int row, float2 (&acc)[H][NSLOT], const float (&inv)[H], int lane) const
    {
        const float sx = scal[0];
        const int rr = row & (TILE - 1);
        uint8_t* base = zimg + size_t(row >> 7) * size_t(NKB) * KBLOCK + uint32_t(rr >> 3) * 1024u + uint32_t(rr & 7) * 128u;
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float f = inv[h] * sx;
#pragma unroll
            for (int r = 0; r < NSLOT; ++r) {
                const int k = 64 * r + 2 * lane;
                if (k < KP) {
                    const int ft = h * KP + k, kb = ft >> 6, e = ft & 63;
                    __half2 hi, lo;
                    split_h2(acc[h][r].x * f, acc[h][r].y * f, hi, lo);
                    uint8_t* p = base + size_t(kb) * KBLOCK + ((uint32_t(e >> 3) ^ uint32_t(rr & 7)) << 4) + uint32_t(e & 7) * 2u;
                    *reinterpret_cast<__half2*>(p) = hi;
                    *reinterpret_cast<__half2*>(p + PLANE) = lo;
                }
            }
        }
    }
    __device__ __forceinline__ void finish(int row, const float (&m)[H], const float (&s)[H], float2 (&acc)[H][NSLOT], int lane) const
    {
        float st[H], inv[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            st[h] = s[h] + 1e-16f;        // PyG softmax: out / (sum + 1e-16)
            inv[h] = 1.f / st[h];
        }
        if (lane == 0) {
            store_vecH<H>(rowmax + int64_t(row) * H, m);
            store_vecH<H>(rowsum + int64_t(row) * H, st);
        }
        write_row(row, acc, inv, lane);
    }
    __device__ __forceinline__ void finish_norm(int row, float2 (&acc)[H][NSLOT], int lane) const
    {
        float inv[H];
#pragma unroll
        for (int h = 0; h < H; ++h) inv[h] = 1.f;
        write_row(row, acc, inv, lane);
    }
    __device__ __forceinline__ void empty(int row, int lane) const
    {
        float m[H], s[H];
        float2 acc[H][NSLOT];
#pragma unroll
        for (int h = 0; h < H; ++h) {
            m[h] = 0.f; s[h] = 0.f;
#pragma unroll
            for (int r = 0; r < NSLOT; ++r) acc[h][r] = make_float2(0.f, 0.f);
        }
        finish(row, m, s, acc, lane);
    }
};
constexpr int PART_ACC = H * NSLOT * 64;     // floats of one partial accumulator set
struct ZPartialSink {   // hub chunk c: unnormalised partial (m, s, acc)
    float* part_ms; float* part_acc; int c;
    __device__ __forceinline__ void finish(int, const float (&m)[H], const float (&s)[H], float2 (&acc)[H][NSLOT], int lane) const
    {
        if (lane == 0) {
            store_vecH<H>(part_ms + int64_t(c) * 2 * H, m);
            store_vecH<H>(part_ms + int64_t(c) * 2 * H + H, s);
        }
        float2* p = reinterpret_cast<float2*>(part_acc + int64_t(c) * PART_ACC);
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int r = 0; r < NSLOT; ++r) p[(h * NSLOT + r) * 32 + lane] = acc[h][r];
    }
    __device__ __forceinline__ void finish_norm(int, float2 (&)[H][NSLOT], int) const {}
    __device__ __forceinline__ void empty(int, int) const {}
};

__device__ __forceinline__ void acc_zero(float2 (&acc)[H][NSLOT])
{
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
        for (int r = 0; r < NSLOT; ++r) acc[h][r] = make_float2(0.f, 0.f);
}

// The stream loop (PACK: packs of whole short rows share one phase A, as in gat_fwd_items_pack).
template <bool VEC2, bool DROPOUT, bool PACK, class Sink>
__device__ __forceinline__ void in_fwd_stream(ChunkCursor& cur, InRing& ring, const Sink& sink, float* rowmax, float* rowsum,
                                              const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                              const int32_t* __restrict__ perm, int K, const float* __restrict__ a_src,
                                              const float* __restrict__ a_dst, float slope, const uint8_t* __restrict__ keep,
                                              float keep_scale, int lane)
{
    auto on_empty = [&](int r) { sink.empty(r, lane); };
    int* r_all = reinterpret_cast<int*>(ring.extra);            // [2][32] row id | last-edge flag per staged edge (packs)
    ChunkStat<H> c0, c1;
    int kind0 = 0, kind1 = 0;
    int b0 = 0, beg = 0, k = 0, la = 0, lb = 0;
    auto next = [&](ChunkStat<H>& c) -> int {
        if (PACK) return cur.next_any(rowptr, lane, c.row, beg, c.n, c.first, c.last, k, la, lb, on_empty);
        return cur.next(rowptr, c.row, beg, c.n, c.first, c.last, on_empty) ? 1 : 0;
    };
    auto phase_a = [&](ChunkStat<H>& c, int kind, int buf) {
        if (PACK && kind == 2)
            fwd_phase_a_pack<GI, DROPOUT>(c.row, beg, c.n, k, la, lb, col, perm, a_src, a_dst, slope, keep, keep_scale, rowmax,
                                          rowsum, ring.p_s + buf * 32 * H, ring.j_s + buf * 32, r_all + buf * 32, lane);
        else
            fwd_phase_a<GI, DROPOUT>(c, beg, col, perm, a_src, a_dst, slope, keep, keep_scale, ring.p_s + buf * 32 * H,
                                     ring.j_s + buf * 32, lane);
    };
    kind0 = next(c0);
    if (!kind0) return;
    phase_a(c0, kind0, b0);
    int issued0 = 0, issued1 = 0;
    float m[H], s[H];
    float2 acc[H][NSLOT];
    while (true) {
        const int* j0 = ring.j_s + b0 * 32;
        const int* j1 = ring.j_s + (b0 ^ 1) * 32;
        while (ring.has_room() && issued0 < c0.n) ring.issue(j0[issued0++], lane);
        kind1 = next(c1);
        issued1 = 0;
        if (kind1) phase_a(c1, kind1, b0 ^ 1);
        if (c0.first) {
#pragma unroll
            for (int h = 0; h < H; ++h) { m[h] = -INFINITY; s[h] = 0.f; }
            acc_zero(acc);
        }
        float fch[H];
        if (PACK && kind0 == 2) {
#pragma unroll
            for (int h = 0; h < H; ++h) fch[h] = 1.f;
        } else {
            float fold[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float mn = fmaxf(m[h], c0.cm[h]);
                fold[h] = expf(m[h] - mn);              // 0 on the first chunk (m = -inf)
                fch[h] = expf(c0.cm[h] - mn);
                s[h] = s[h] * fold[h] + c0.cs[h] * fch[h];
                m[h] = mn;
            }
            if (!c0.first) {
#pragma unroll
                for (int h = 0; h < H; ++h)
#pragma unroll
                    for (int r = 0; r < NSLOT; ++r) { acc[h][r].x *= fold[h]; acc[h][r].y *= fold[h]; }
            }
        }
        const float* p0 = ring.p_s + b0 * 32 * H;
        const int* r0 = r_all + b0 * 32;
        for (int t = 0; t < c0.n; ++t) {
            const float* row = ring.front(j0[t]);
            float2 v[NSLOT];
            load_xrow<VEC2>(row, lane, K, v);
            float w[H];
            {
                const float4 w0 = *reinterpret_cast<const float4*>(p0 + t * H);
                const float4 w1 = *reinterpret_cast<const float4*>(p0 + t * H + 4);
                w[0] = w0.x * fch[0]; w[1] = w0.y * fch[1]; w[2] = w0.z * fch[2]; w[3] = w0.w * fch[3];
                w[4] = w1.x * fch[4]; w[5] = w1.y * fch[5]; w[6] = w1.z * fch[6]; w[7] = w1.w * fch[7];
            }
#pragma unroll
            for (int h = 0; h < H; ++h)
#pragma unroll
                for (int r = 0; r < NSLOT; ++r) {
                    acc[h][r].x = fmaf(w[h], v[r].x, acc[h][r].x);
                    acc[h][r].y = fmaf(w[h], v[r].y, acc[h][r].y);
                }
            ring.pop();
            if (issued0 < c0.n) ring.issue(j0[issued0++], lane);
            else if (kind1 && issued1 < c1.n) ring.issue(j1[issued1++], lane);
            if (PACK && kind0 == 2) {
                const int rs = r0[t];
                if (rs < 0) {                                   // last edge of a packed row (weights already normalised)
                    sink.finish_norm(rs & 0x7fffffff, acc, lane);
                    acc_zero(acc);
                }
            }
        }
        if (kind0 == 1 && c0.last) sink.finish(c0.row, m, s, acc, lane);
        if (!kind1) break;
        c0 = c1;
        kind0 = kind1;
        issued0 = issued1;
        b0 ^= 1;
    }
}

constexpr int FWD_EXTRA = 256;   // r_all

template <bool VEC2, bool DROPOUT, bool PACK>
__global__ void __launch_bounds__(IN_THREADS, 3)
gat_in_fwd_items(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                 const float* __restrict__ x, int64_t ldx, int K, const float* __restrict__ a_src,
                 const float* __restrict__ a_dst, gnnfd_item_plan_t items, int hub_threshold, float slope,
                 const uint8_t* __restrict__ keep, float keep_scale, ZSink sink)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = blockIdx.x * IN_WARPS + warp;
    if (item >= items.n_items) return;
    InRing ring;
    ring.init(smem + warp * in_warp_bytes(K, FWD_EXTRA), x, ldx, K, lane);
    ChunkCursor cur;
    cur.start_rows(items.item_start[item], items.item_start[item + 1], hub_threshold);
    in_fwd_stream<VEC2, DROPOUT, PACK>(cur, ring, sink, sink.rowmax, sink.rowsum, rowptr, col, perm, K, a_src, a_dst, slope,
                                       keep, keep_scale, lane);
}

// one warp per (hub row, chunk): partial (m, s, unnormalised acc)
template <bool VEC2, bool DROPOUT>
__global__ void __launch_bounds__(IN_THREADS, 3)
gat_in_fwd_hub_chunks(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                      const float* __restrict__ x, int64_t ldx, int K, const float* __restrict__ a_src,
                      const float* __restrict__ a_dst, gnnfd_hub_plan_t plan, float slope,
                      const uint8_t* __restrict__ keep, float keep_scale, float* __restrict__ part_ms,
                      float* __restrict__ part_acc)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * IN_WARPS + warp;
    if (c >= plan.n_chunk) return;
    const int slot = plan.chunk_hub[c];
    const int i = plan.hub_row[slot];
    const int beg = rowptr[i] + (c - plan.hub_chunk_ptr[slot]) * plan.chunk;
    const int end = min(rowptr[i + 1], beg + plan.chunk);
    InRing ring;
    ring.init(smem + warp * in_warp_bytes(K, FWD_EXTRA), x, ldx, K, lane);
    ChunkCursor cur;
    cur.start_segment(i, beg, end);
    ZPartialSink sink{part_ms, part_acc, c};
    in_fwd_stream<VEC2, DROPOUT, false>(cur, ring, sink, nullptr, nullptr, rowptr, col, perm, K, a_src, a_dst, slope, keep,
                                        keep_scale, lane);
}

// one CTA per hub row: warp w folds chunks w, w+8, ... (online-softmax combine), the eight warp states are folded in
// warp order -- a fixed order, so the result is deterministic
__global__ void __launch_bounds__(ROW_THREADS)
gat_in_fwd_hub_merge(gnnfd_hub_plan_t plan, const float* __restrict__ part_ms, const float* __restrict__ part_acc, ZSink sink)
{
    extern __shared__ __align__(16) float msm[];               // [ROW_WARPS][2H] then [ROW_WARPS][PART_ACC]
    float* st_ms = msm;
    float2* st_acc = reinterpret_cast<float2*>(msm + ROW_WARPS * 2 * H);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x;
    const int i = plan.hub_row[slot];
    const int c0 = plan.hub_chunk_ptr[slot], c1 = plan.hub_chunk_ptr[slot + 1];
    float M[H], s[H];
    float2 acc[H][NSLOT];
#pragma unroll
    for (int h = 0; h < H; ++h) { M[h] = -INFINITY; s[h] = 0.f; }
    acc_zero(acc);
    auto fold = [&](const float (&mc)[H], const float (&sc)[H], const float2* __restrict__ pacc) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float mn = fmaxf(M[h], mc[h]);
            const float fo = (M[h] == -INFINITY) ? 0.f : expf(M[h] - mn);
            const float fn = (mc[h] == -INFINITY) ? 0.f : expf(mc[h] - mn);
            s[h] = s[h] * fo + sc[h] * fn;
            M[h] = mn;
#pragma unroll
            for (int r = 0; r < NSLOT; ++r) {
                const float2 v = pacc[(h * NSLOT + r) * 32 + lane];
                acc[h][r].x = fmaf(v.x, fn, acc[h][r].x * fo);
                acc[h][r].y = fmaf(v.y, fn, acc[h][r].y * fo);
            }
        }
    };
    for (int c = c0 + warp; c < c1; c += ROW_WARPS) {
        float mc[H], sc[H];
        load_vecH<H>(part_ms + int64_t(c) * 2 * H, mc);
        load_vecH<H>(part_ms + int64_t(c) * 2 * H + H, sc);
        fold(mc, sc, reinterpret_cast<const float2*>(part_acc + int64_t(c) * PART_ACC));
    }
    if (lane == 0) {
        store_vecH<H>(st_ms + warp * 2 * H, M);
        store_vecH<H>(st_ms + warp * 2 * H + H, s);
    }
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
        for (int r = 0; r < NSLOT; ++r) st_acc[warp * (PART_ACC / 2) + (h * NSLOT + r) * 32 + lane] = acc[h][r];
    __syncthreads();
    if (warp != 0) return;
    for (int w = 1; w < ROW_WARPS; ++w) {
        float mc[H], sc[H];
#pragma unroll
        for (int h = 0; h < H; ++h) { mc[h] = st_ms[w * 2 * H + h]; sc[h] = st_ms[w * 2 * H + H + h]; }
        fold(mc, sc, st_acc + w * (PART_ACC / 2));
    }
    sink.finish(i, M, s, acc, lane);
}

template <class Kn>
static int in_set_smem(Kn kernel, int bytes)
{
    GNNFD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return GNNFD_OK;
}

static bool vec2_ok(const float* x, int64_t ldx, int K)
{
    return (K % 2 == 0) && (ldx % 2 == 0) && (reinterpret_cast<uintptr_t>(x) & 7) == 0;
}

}  // namespace in
}  // namespace gnnfd

using namespace gnnfd;
using namespace gnnfd::in;

extern "C" {

int gnnfd_in_supported(int64_t K, int H_, int C_, int concat)
{
    return (K >= 1 && K <= MAX_K && H_ == H && C_ == C && !concat) ? 1 : 0;
}

int gnnfd_in_sizes(int64_t n_dst, int64_t K, size_t* prep_bytes_out, size_t* zimg_bytes_out, int64_t* gd_ld_out)
{
    GNNFD_REQUIRE(K >= 1 && K <= MAX_K && n_dst >= 0, GNNFD_ERR_ARG, "in_sizes: K must be in [1,%d]", MAX_K);
    const Dims d((int)K);
    if (prep_bytes_out) *prep_bytes_out = prep_bytes(d);
    if (zimg_bytes_out) *zimg_bytes_out = zimg_bytes(n_dst, d) + 1024;
    if (gd_ld_out) *gd_ld_out = d.F;
    return GNNFD_OK;
}

/* a_src / a_dst [N,H] for rows [0,N) of x, and xmax[0] = max(xmax[0], max |x|) (caller zeroes xmax before the first call;
 * across GPUs the per-rank maxima are max-reduced before gnnfd_in_prepare).  ws: 2H*KP floats. */
int gnnfd_in_logits(const float* x, int64_t ldx, int64_t N, int64_t K, const float* W, const float* att_src,
                    const float* att_dst, float* a_src, float* a_dst, float* xmax, void* prep, gnnfd_stream_t stream)
{
    GNNFD_REQUIRE(K >= 1 && K <= MAX_K && ldx >= K && N >= 0, GNNFD_ERR_ARG, "in_logits: bad shape (K <= %d)", MAX_K);
    GNNFD_REQUIRE(W && att_src && att_dst && xmax && prep, GNNFD_ERR_ARG, "in_logits: NULL argument");
    GNNFD_REQUIRE(N == 0 || (x && a_src && a_dst), GNNFD_ERR_ARG, "in_logits: NULL tensor");
    cudaStream_t st = (cudaStream_t)stream;
    const Dims d((int)K);
    float* u = reinterpret_cast<float*>(reinterpret_cast<char*>(prep) + prep_off_u(d));
    in_u_kernel<<<2 * H, 192, 0, st>>>(W, att_src, att_dst, d.K, d.KP, u);
    g_launches += 1;
    if (N > 0) {
        int64_t blocks = (N + 7) / 8;
        const int64_t cap = int64_t(sm_count()) * 2 * 4;
        if (blocks > cap) blocks = cap;
        if (vec2_ok(x, ldx, d.K))
            in_logits_kernel<true><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, N, d.K, u, d.KP, a_src, a_dst,
                                                                    reinterpret_cast<unsigned*>(xmax));
        else
            in_logits_kernel<false><<<(unsigned)blocks, 256, 0, st>>>(x, ldx, N, d.K, u, d.KP, a_src, a_dst,
                                                                     reinterpret_cast<unsigned*>(xmax));
        g_launches += 1;
    }
    GNNFD_LAUNCH_CHECK();
    return GNNFD_OK;
}

/* scales + W images of the two dense stages into prep (gnnfd_in_sizes gives its size). */
int gnnfd_in_prepare(const float* W, int64_t K, const float* xmax, void* prep, gnnfd_stream_t stream)
{
    GNNFD_REQUIRE(K >= 1 && K <= MAX_K && W && xmax && prep, GNNFD_ERR_ARG, "in_prepare: bad argument");
    GNNFD_REQUIRE((reinterpret_cast<uintptr_t>(prep) & 1023) == 0, GNNFD_ERR_ARG, "in_prepare: prep must be 1024-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Dims d((int)K);
    char* p = reinterpret_cast<char*>(prep);
    float* scal = reinterpret_cast<float*>(p);
    in_scales_kernel<<<1, 1024, 0, st>>>(W, int64_t(H) * C * K, xmax, scal);
    in_wout_image_kernel<<<d.NKB, 256, 0, st>>>(W, d.K, d.KP, d.NKB, scal, reinterpret_cast<uint8_t*>(p + prep_off_wout(d)));
    g_launches += 2;
    GNNFD_LAUNCH_CHECK();
    return in_build_gd_image(W, d.K, prep, st);
}

int gnnfd_in_fwd_workspace_bytes(const gnnfd_graph_t* g, size_t* bytes)
{
    GNNFD_REQUIRE(g && bytes, GNNFD_ERR_ARG, "in_fwd_workspace_bytes: NULL argument");
    const size_t nc = (size_t)g->hub_dst.n_chunk;
    *bytes = carve_bytes(nc * 2 * H, 4) + carve_bytes(nc * size_t(PART_ACC), 4) + 256;
    return GNNFD_OK;
}

/* Aggregation in input space: zimg (fp16-pair image of Z, rows = destinations), rowmax / rowsum [n_dst,H]. */
int gnnfd_in_fwd(const gnnfd_graph_t* g, const float* x, int64_t ldx, int64_t K, const float* a_src, const float* a_dst,
                 float negative_slope, const uint8_t* keep_mask, float p_drop, const void* prep, void* zimg,
                 float* rowmax, float* rowsum, void* ws, size_t ws_bytes, gnnfd_stream_t stream)
{
    int rc = check_graph(g, false, "in_fwd");
    if (rc) return rc;
    GNNFD_REQUIRE(K >= 1 && K <= MAX_K && ldx >= K, GNNFD_ERR_ARG, "in_fwd: bad shape (K <= %d)", MAX_K);
    GNNFD_REQUIRE(g->n_dst == 0 || (x && a_src && a_dst && prep && zimg && rowmax && rowsum), GNNFD_ERR_ARG, "in_fwd: NULL tensor");
    GNNFD_REQUIRE(p_drop >= 0.f && p_drop <= 0.9f, GNNFD_ERR_ARG, "in_fwd: dropout p must be in [0,0.9] on the input-space path");
    GNNFD_REQUIRE((reinterpret_cast<uintptr_t>(zimg) & 1023) == 0 && (reinterpret_cast<uintptr_t>(x) & 3) == 0, GNNFD_ERR_ARG,
                  "in_fwd: zimg must be 1024-byte aligned");
    GNNFD_REQUIRE(g->n_src * ldx * 4 < (int64_t(1) << 46), GNNFD_ERR_ARG, "in_fwd: x too large");
    if (g->n_dst == 0) return GNNFD_OK;
    GNNFD_REQUIRE(g->items_dst.n_items > 0 && g->items_dst.item_start, GNNFD_ERR_ARG,
                  "in_fwd: the graph has no work-item plan over rowptr (gnnfd_item_plan)");
    cudaStream_t st = (cudaStream_t)stream;
    const Dims d((int)K);
    const bool drop = keep_mask != nullptr && p_drop > 0.f;
    const float ks = drop ? 1.f / (1.f - p_drop) : 1.f;
    const int thr = g->hub_dst.n_hub > 0 ? g->hub_dst.threshold : INT_MAX;
    const bool v2 = vec2_ok(x, ldx, d.K);
    const int smem = IN_WARPS * in_warp_bytes(d.K, FWD_EXTRA);
    const unsigned grid = (unsigned)((g->items_dst.n_items + IN_WARPS - 1) / IN_WARPS);
    const float* scal = reinterpret_cast<const float*>(prep);
    ZSink sink{reinterpret_cast<uint8_t*>(zimg), rowmax, rowsum, scal, d.KP, d.NKB};
    // rows of the last tile beyond n_dst are read by the node-reduction GEMM (dW): they must be zero, not stale
    if (g->n_dst % TILE) {
        const size_t last = size_t(g->n_dst / TILE) * d.NKB * KBLOCK;
        GNNFD_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(zimg) + last, 0, size_t(d.NKB) * KBLOCK, st));
    }
    static const bool pack = [] {
        const char* e = getenv("GNNFD_FWD_PACK");
        return e ? atoi(e) != 0 : true;
    }();
#define GNNFD_IN_FWD(VV, DD, PP)                                                                                       \
    rc = in_set_smem(gat_in_fwd_items<VV, DD, PP>, smem);                                                              \
    if (rc) return rc;                                                                                                 \
    gat_in_fwd_items<VV, DD, PP><<<grid, IN_THREADS, smem, st>>>(g->rowptr, g->col, g->perm, x, ldx, d.K, a_src, a_dst, \
                                                                 g->items_dst, thr, negative_slope, keep_mask, ks, sink)
    if (pack) {
        if (v2) { if (drop) { GNNFD_IN_FWD(true, true, true); } else { GNNFD_IN_FWD(true, false, true); } }
        else    { if (drop) { GNNFD_IN_FWD(false, true, true); } else { GNNFD_IN_FWD(false, false, true); } }
    } else {
        if (v2) { if (drop) { GNNFD_IN_FWD(true, true, false); } else { GNNFD_IN_FWD(true, false, false); } }
        else    { if (drop) { GNNFD_IN_FWD(false, true, false); } else { GNNFD_IN_FWD(false, false, false); } }
    }
#undef GNNFD_IN_FWD
    g_launches += 1;
    if (g->hub_dst.n_hub > 0) {
        const gnnfd_hub_plan_t& pl = g->hub_dst;
        const size_t need = carve_bytes(size_t(pl.n_chunk) * 2 * H, 4) + carve_bytes(size_t(pl.n_chunk) * PART_ACC, 4);
        GNNFD_REQUIRE(ws && ws_bytes >= need, GNNFD_ERR_WORKSPACE, "in_fwd: workspace %zu < %zu", ws_bytes, need);
        char* p = reinterpret_cast<char*>(ws);
        float* part_ms = carve<float>(p, size_t(pl.n_chunk) * 2 * H);
        float* part_acc = carve<float>(p, size_t(pl.n_chunk) * PART_ACC);
        const unsigned gc = (unsigned)((pl.n_chunk + IN_WARPS - 1) / IN_WARPS);
#define GNNFD_IN_HUB(VV, DD)                                                                                           \
    rc = in_set_smem(gat_in_fwd_hub_chunks<VV, DD>, smem);                                                             \
    if (rc) return rc;                                                                                                 \
    gat_in_fwd_hub_chunks<VV, DD><<<gc, IN_THREADS, smem, st>>>(g->rowptr, g->col, g->perm, x, ldx, d.K, a_src, a_dst, pl, \
                                                                negative_slope, keep_mask, ks, part_ms, part_acc)
        if (v2) { if (drop) { GNNFD_IN_HUB(true, true); } else { GNNFD_IN_HUB(true, false); } }
        else    { if (drop) { GNNFD_IN_HUB(false, true); } else { GNNFD_IN_HUB(false, false); } }
#undef GNNFD_IN_HUB
        const int msm = (ROW_WARPS * 2 * H + ROW_WARPS * PART_ACC) * 4;
        rc = in_set_smem(gat_in_fwd_hub_merge, msm);
        if (rc) return rc;
        gat_in_fwd_hub_merge<<<(unsigned)pl.n_hub, ROW_THREADS, msm, st>>>(pl, part_ms, part_acc, sink);
        g_launches += 2;
    }
    GNNFD_LAUNCH_CHECK();
    return GNNFD_OK;
}

}  // extern "C"
