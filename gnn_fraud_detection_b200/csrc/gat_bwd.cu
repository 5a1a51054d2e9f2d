// (4) backward of the fused attention / aggregation (autograd mirror of GATConv.forward, triggered by
// loss.backward() at src/train.py:142; closed forms in SURVEY.md 8(a3)).
//
// dst-major pass (CSR, warp per destination row): recompute alpha from the saved row statistics,
//   d_alpha[e,h] = <dO_h[i], xw[j]> (the only 2 KB/edge gather of the backward), softmax backward with
//   t = sum_k alpha_k d_alpha_k taken over the row, LeakyReLU backward; writes alpha_used and dz in CSR
//   order and da_dst.  Rows of <= 32 edges stay in registers; longer rows make a second cheap sweep
//   (no feature gather); hub rows are split into chunks with deterministic partial merges.
// src-major pass (CSC, warp per source row): dxw[j] = sum_e alpha_used[e] * dO_h[dst(e)] -- with the
//   head mean every head shares dOut[i]/H, so the per-edge gather is only C*4 = 256 B -- plus the logit
//   terms da_src[j]*att_src + da_dst[j]*att_dst fused into the row epilogue; da_src[j] = sum_e dz[e].
#include "gat_common.cuh"

#include <atomic>
#include <climits>

namespace gnnfd {
extern std::atomic<long long> g_launches;
int check_graph(const gnnfd_graph_t* g, bool need_csc, const char* who);

constexpr int BWD_U = 4;

// per-row constants of the dst-major pass
template <class GE>
struct DstRow {
    float adst[GE::H], m[GE::H], inv[GE::H];
    float g[GE::NS][GE::VW];  // dO_h slice owned by this lane (already divided by H for the head mean)
};

template <class GE, bool CONCAT>
__device__ __forceinline__ void load_dst_row(DstRow<GE>& r, int64_t i, const float* __restrict__ a_dst,
                                             const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                                             const float* __restrict__ d_out, int lane)
{
    constexpr int H = GE::H, NS = GE::NS, VW = GE::VW, C = GE::C, D = GE::D;
    float st[H];
    load_vecH<H>(a_dst + i * H, r.adst);
    load_vecH<H>(rowmax + i * H, r.m);
    load_vecH<H>(rowsum + i * H, st);
#pragma unroll
    for (int h = 0; h < H; ++h) r.inv[h] = 1.f / st[h];
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        const int e0 = VW * (lane + 32 * q);
        const float* p = CONCAT ? d_out + i * D + e0 : d_out + i * C + (e0 % C);
#pragma unroll
        for (int k = 0; k < VW; k += 4) {
            const float4 t = __ldg(reinterpret_cast<const float4*>(p + k));
            const float sc = CONCAT ? 1.f : 1.f / H;
            r.g[q][k] = t.x * sc; r.g[q][k + 1] = t.y * sc; r.g[q][k + 2] = t.z * sc; r.g[q][k + 3] = t.w * sc;
        }
    }
}

// One chunk (<= 32 edges) of the first sweep.  On return lane e (< n) holds alpha, u = alpha*d_alpha,
// the LeakyReLU slope and the dropout scale of its edge.
template <class GE, bool DROPOUT>
__device__ __forceinline__ void dst_chunk(const DstRow<GE>& r, int base, int n, const int32_t* __restrict__ col,
                                          const int32_t* __restrict__ perm,
                                          const typename GE::XT* __restrict__ xw, const float* __restrict__ a_src,
                                          float slope, const uint8_t* __restrict__ keep, float keep_scale,
                                          float* dal_s, int* j_s, int lane, float (&alpha)[GE::H], float (&u)[GE::H],
                                          float (&sl)[GE::H], float (&ks)[GE::H])
{
    constexpr int H = GE::H, NS = GE::NS, VW = GE::VW, HP = GE::HP, D = GE::D, G = GE::G;
    const int sub = lane / G;
    int j = 0;
    if (lane < n) {
        j = col[base + lane];
        float as[H];
        load_vecH<H>(a_src + int64_t(j) * H, as);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float z = as[h] + r.adst[h];
            sl[h] = z > 0.f ? 1.f : slope;
            alpha[h] = expf(z * sl[h] - r.m[h]) * r.inv[h];
            ks[h] = 1.f;
        }
        if (DROPOUT) {
            const uint8_t* kb = keep + int64_t(perm[base + lane]) * H;
#pragma unroll
            for (int h = 0; h < H; ++h) ks[h] = kb[h] ? keep_scale : 0.f;
        }
    } else {
#pragma unroll
        for (int h = 0; h < H; ++h) { alpha[h] = 0.f; sl[h] = 0.f; ks[h] = 0.f; }
    }
    j_s[lane] = j;
    __syncwarp();
    for (int t = 0; t < n; t += BWD_U) {
        float v[BWD_U][NS][VW];
#pragma unroll
        for (int uu = 0; uu < BWD_U; ++uu) {
            const bool ok = t + uu < n;
            const int tt = ok ? t + uu : t;
            const typename GE::XT* row = xw + int64_t(j_s[tt]) * D;
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                if (ok) load_slot(row, q, lane, v[uu][q]);
                else {
#pragma unroll
                    for (int k = 0; k < VW; ++k) v[uu][q][k] = 0.f;
                }
            }
        }
#pragma unroll
        for (int uu = 0; uu < BWD_U; ++uu) {
#pragma unroll
            for (int q = 0; q < NS; ++q) {
                float d = 0.f;
#pragma unroll
                for (int k = 0; k < VW; ++k) d = fmaf(r.g[q][k], v[uu][q][k], d);
#pragma unroll
                for (int o = G / 2; o > 0; o >>= 1) d += __shfl_xor_sync(FULL, d, o);
                if ((lane & (G - 1)) == 0 && t + uu < n) dal_s[(t + uu) * H + q * HP + sub] = d;
            }
        }
    }
    __syncwarp();
    float dal[H];
#pragma unroll
    for (int k = 0; k < H / 4; ++k) {
        const float4 tq = *reinterpret_cast<const float4*>(dal_s + lane * H + 4 * k);
        dal[4 * k] = tq.x; dal[4 * k + 1] = tq.y; dal[4 * k + 2] = tq.z; dal[4 * k + 3] = tq.w;
    }
#pragma unroll
    for (int h = 0; h < H; ++h) u[h] = (lane < n) ? alpha[h] * dal[h] * ks[h] : 0.f;
    __syncwarp();
}

// first sweep over [beg,end): writes alpha_used and u (into the dz buffer), returns t = sum u (all lanes)
template <class GE, bool DROPOUT>
__device__ __forceinline__ void dst_sweep1(const DstRow<GE>& r, int beg, int end, const int32_t* __restrict__ col,
                                           const int32_t* __restrict__ perm,
                                           const typename GE::XT* __restrict__ xw, const float* __restrict__ a_src,
                                           float slope, const uint8_t* __restrict__ keep, float keep_scale,
                                           float* dal_s, int* j_s, int lane, float* __restrict__ alpha_used,
                                           float* __restrict__ dz, float (&t)[GE::H])
{
    constexpr int H = GE::H;
    float tl[H];
#pragma unroll
    for (int h = 0; h < H; ++h) tl[h] = 0.f;
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        float alpha[H], u[H], sl[H], ks[H];
        dst_chunk<GE, DROPOUT>(r, base, n, col, perm, xw, a_src, slope, keep, keep_scale, dal_s, j_s, lane, alpha, u, sl, ks);
        if (lane < n) {
            float au[H];
#pragma unroll
            for (int h = 0; h < H; ++h) { au[h] = alpha[h] * ks[h]; tl[h] += u[h]; }
            store_vecH<H>(alpha_used + int64_t(base + lane) * H, au);
            store_vecH<H>(dz + int64_t(base + lane) * H, u);
        }
    }
#pragma unroll
    for (int h = 0; h < H; ++h) t[h] = warp_sum(tl[h]);
}

// second sweep: dz = slope * (u - alpha * t); returns the lane-local partial of da_dst
template <class GE>
__device__ __forceinline__ void dst_sweep2(const DstRow<GE>& r, int beg, int end, const int32_t* __restrict__ col,
                                           const float* __restrict__ a_src, float slope, const float (&t)[GE::H],
                                           int lane, float* __restrict__ dz, float (&dad)[GE::H])
{
    constexpr int H = GE::H;
    for (int e = beg + lane; e < end; e += 32) {
        float as[H], u[H], o[H];
        load_vecH<H>(a_src + int64_t(col[e]) * H, as);
        load_vecH<H>(dz + int64_t(e) * H, u);
#pragma unroll
        for (int h = 0; h < H; ++h) {
            const float z = as[h] + r.adst[h];
            const float sl = z > 0.f ? 1.f : slope;
            const float al = expf(z * sl - r.m[h]) * r.inv[h];
            o[h] = sl * (u[h] - al * t[h]);
            dad[h] += o[h];
        }
        store_vecH<H>(dz + int64_t(e) * H, o);
    }
}

template <class GE, bool CONCAT, bool DROPOUT>
__global__ void __launch_bounds__(ROW_THREADS)
gat_bwd_dst_rows(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                 const typename GE::XT* __restrict__ xw, const float* __restrict__ a_src,
                 const float* __restrict__ a_dst, const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                 const float* __restrict__ d_out, int64_t n_dst, int hub_threshold, float slope,
                 const uint8_t* __restrict__ keep, float keep_scale, float* __restrict__ alpha_used,
                 float* __restrict__ dz, float* __restrict__ da_dst)
{
    constexpr int H = GE::H;
    __shared__ __align__(16) float dal_sh[ROW_WARPS][32 * H];
    __shared__ int j_sh[ROW_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t i = int64_t(blockIdx.x) * ROW_WARPS + warp;
    if (i >= n_dst) return;
    const int beg = rowptr[i], end = rowptr[i + 1];
    if (end - beg > hub_threshold) return;
    DstRow<GE> r;
    load_dst_row<GE, CONCAT>(r, i, a_dst, rowmax, rowsum, d_out, lane);
    float dad[H];
#pragma unroll
    for (int h = 0; h < H; ++h) dad[h] = 0.f;
    if (end - beg <= 32) {
        const int n = end - beg;
        float alpha[H], u[H], sl[H], ks[H];
        if (n > 0) {
            dst_chunk<GE, DROPOUT>(r, beg, n, col, perm, xw, a_src, slope, keep, keep_scale, dal_sh[warp], j_sh[warp],
                                   lane, alpha, u, sl, ks);
            float au[H], o[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float t = warp_sum(u[h]);
                o[h] = sl[h] * (u[h] - alpha[h] * t);
                au[h] = alpha[h] * ks[h];
                dad[h] = o[h];
            }
            if (lane < n) {
                store_vecH<H>(alpha_used + int64_t(beg + lane) * H, au);
                store_vecH<H>(dz + int64_t(beg + lane) * H, o);
            }
        }
    } else {
        float t[H];
        dst_sweep1<GE, DROPOUT>(r, beg, end, col, perm, xw, a_src, slope, keep, keep_scale, dal_sh[warp], j_sh[warp],
                                lane, alpha_used, dz, t);
        __syncwarp();
        dst_sweep2<GE>(r, beg, end, col, a_src, slope, t, lane, dz, dad);
    }
#pragma unroll
    for (int h = 0; h < H; ++h) dad[h] = warp_sum(dad[h]);
    if (lane == 0) store_vecH<H>(da_dst + i * H, dad);
}

// hub rows, step 1: one warp per chunk, first sweep, partial t
template <class GE, bool CONCAT, bool DROPOUT>
__global__ void __launch_bounds__(ROW_THREADS)
gat_bwd_dst_hub1(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                 const typename GE::XT* __restrict__ xw, const float* __restrict__ a_src,
                 const float* __restrict__ a_dst, const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                 const float* __restrict__ d_out, gnnfd_hub_plan_t plan, float slope,
                 const uint8_t* __restrict__ keep, float keep_scale, float* __restrict__ alpha_used,
                 float* __restrict__ dz, float* __restrict__ part_t)
{
    constexpr int H = GE::H;
    __shared__ __align__(16) float dal_sh[ROW_WARPS][32 * H];
    __shared__ int j_sh[ROW_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * ROW_WARPS + warp;
    if (c >= plan.n_chunk) return;
    const int slot = plan.chunk_hub[c];
    const int64_t i = plan.hub_row[slot];
    const int beg = rowptr[i] + (c - plan.hub_chunk_ptr[slot]) * plan.chunk;
    const int end = min(rowptr[i + 1], beg + plan.chunk);
    DstRow<GE> r;
    load_dst_row<GE, CONCAT>(r, i, a_dst, rowmax, rowsum, d_out, lane);
    float t[H];
    dst_sweep1<GE, DROPOUT>(r, beg, end, col, perm, xw, a_src, slope, keep, keep_scale, dal_sh[warp], j_sh[warp], lane,
                            alpha_used, dz, t);
    if (lane == 0) store_vecH<H>(part_t + int64_t(c) * H, t);
}
// hub rows, step 2: total t of the row (chunk order), second sweep, partial da_dst
template <class GE>
__global__ void __launch_bounds__(ROW_THREADS)
gat_bwd_dst_hub2(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const float* __restrict__ a_src,
                 const float* __restrict__ a_dst, const float* __restrict__ rowmax, const float* __restrict__ rowsum,
                 gnnfd_hub_plan_t plan, float slope, const float* __restrict__ part_t, float* __restrict__ dz,
                 float* __restrict__ part_dad)
{
    constexpr int H = GE::H;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * ROW_WARPS + warp;
    if (c >= plan.n_chunk) return;
    const int slot = plan.chunk_hub[c];
    const int64_t i = plan.hub_row[slot];
    const int c0 = plan.hub_chunk_ptr[slot], c1 = plan.hub_chunk_ptr[slot + 1];
    const int beg = rowptr[i] + (c - c0) * plan.chunk;
    const int end = min(rowptr[i + 1], beg + plan.chunk);
    DstRow<GE> r;
    float st[H];
    load_vecH<H>(a_dst + i * H, r.adst);
    load_vecH<H>(rowmax + i * H, r.m);
    load_vecH<H>(rowsum + i * H, st);
    float t[H], dad[H];
#pragma unroll
    for (int h = 0; h < H; ++h) { r.inv[h] = 1.f / st[h]; t[h] = 0.f; dad[h] = 0.f; }
    for (int cc = c0; cc < c1; ++cc) {   // same order in every chunk of the row => identical t
        float pt[H];
        load_vecH<H>(part_t + int64_t(cc) * H, pt);
#pragma unroll
        for (int h = 0; h < H; ++h) t[h] += pt[h];
    }
    dst_sweep2<GE>(r, beg, end, col, a_src, slope, t, lane, dz, dad);
#pragma unroll
    for (int h = 0; h < H; ++h) dad[h] = warp_sum(dad[h]);
    if (lane == 0) store_vecH<H>(part_dad + int64_t(c) * H, dad);
}
// hub rows, step 3: da_dst[i] = sum of the chunk partials in chunk order
template <int H>
__global__ void gat_bwd_dst_hub3(gnnfd_hub_plan_t plan, const float* __restrict__ part_dad, float* __restrict__ da_dst)
{
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= plan.n_hub * H) return;
    const int slot = idx / H, h = idx % H;
    float s = 0.f;
    for (int c = plan.hub_chunk_ptr[slot]; c < plan.hub_chunk_ptr[slot + 1]; ++c) s += part_dad[int64_t(c) * H + h];
    da_dst[int64_t(plan.hub_row[slot]) * H + h] = s;
}

// ---------------------------------------------------------------------------------------------------------
// src-major pass
// ---------------------------------------------------------------------------------------------------------
template <class GE, bool CONCAT>
__device__ __forceinline__ void src_range(int beg, int end, const int32_t* __restrict__ csc_row,
                                          const int32_t* __restrict__ csc_eid, const float* __restrict__ alpha_used,
                                          const float* __restrict__ dz, const float* __restrict__ d_out,
                                          float (&acc)[GE::NS][4], float (&das)[GE::H], float* p_s, int* i_s, int lane)
{
    constexpr int H = GE::H, NS = GE::NS, HP = GE::HP, D = GE::D, C = GE::C;
    const int sub = lane / GE::G;
    const int cm = (4 * lane) % C;   // C divides 128, so every slot of this lane has the same channel offset
    for (int base = beg; base < end; base += 32) {
        const int n = min(32, end - base);
        int i = 0;
        if (lane < n) {
            const int64_t eid = csc_eid[base + lane];
            i = csc_row[base + lane];
            float al[H], dzv[H];
            load_vecH<H>(alpha_used + eid * H, al);
            load_vecH<H>(dz + eid * H, dzv);
#pragma unroll
            for (int h = 0; h < H; ++h) das[h] += dzv[h];
            store_vecH<H>(p_s + lane * H, al);
        }
        i_s[lane] = i;
        __syncwarp();
        for (int t = 0; t < n; t += BWD_U) {
            float g[BWD_U][CONCAT ? NS : 1][4], wq[BWD_U][NS];
#pragma unroll
            for (int uu = 0; uu < BWD_U; ++uu) {
                const bool ok = t + uu < n;
                const int tt = ok ? t + uu : t;
                const int64_t ii = i_s[tt];
#pragma unroll
                for (int q = 0; q < (CONCAT ? NS : 1); ++q) {
                    const float* p = CONCAT ? d_out + ii * D + 4 * (lane + 32 * q) : d_out + ii * C + cm;
                    float4 tv = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (ok) tv = CONCAT ? ldg_stream(reinterpret_cast<const float4*>(p)) : __ldg(reinterpret_cast<const float4*>(p));
                    g[uu][q][0] = tv.x; g[uu][q][1] = tv.y; g[uu][q][2] = tv.z; g[uu][q][3] = tv.w;
                }
#pragma unroll
                for (int q = 0; q < NS; ++q) wq[uu][q] = ok ? p_s[tt * H + q * HP + sub] : 0.f;
            }
#pragma unroll
            for (int uu = 0; uu < BWD_U; ++uu)
#pragma unroll
                for (int q = 0; q < NS; ++q)
#pragma unroll
                    for (int k = 0; k < 4; ++k) acc[q][k] = fmaf(wq[uu][q], g[uu][CONCAT ? q : 0][k], acc[q][k]);
        }
        __syncwarp();
    }
}

template <class GE, bool CONCAT>
__device__ __forceinline__ void src_epilogue(int64_t j, float (&acc)[GE::NS][4], const float (&das)[GE::H],
                                             const float* __restrict__ att_src, const float* __restrict__ att_dst,
                                             const float* __restrict__ da_dst_full, float* __restrict__ dxw,
                                             float* __restrict__ da_src, int lane)
{
    constexpr int H = GE::H, NS = GE::NS, HP = GE::HP, D = GE::D;
    const int sub = lane / GE::G;
    float dad[H];
    if (da_dst_full) load_vecH<H>(da_dst_full + j * H, dad);
    else {
#pragma unroll
        for (int h = 0; h < H; ++h) dad[h] = 0.f;
    }
    if (lane == 0) store_vecH<H>(da_src + j * H, das);
#pragma unroll
    for (int q = 0; q < NS; ++q) {
        const int e0 = 4 * (lane + 32 * q);
        const float fs = pick<HP>(das, q, sub), fd = pick<HP>(dad, q, sub);
        const float4 as4 = __ldg(reinterpret_cast<const float4*>(att_src + e0));
        const float4 ad4 = __ldg(reinterpret_cast<const float4*>(att_dst + e0));
        const float sc = CONCAT ? 1.f : 1.f / H;
        float4 o;
        o.x = fmaf(acc[q][0], sc, fmaf(fs, as4.x, fd * ad4.x));
        o.y = fmaf(acc[q][1], sc, fmaf(fs, as4.y, fd * ad4.y));
        o.z = fmaf(acc[q][2], sc, fmaf(fs, as4.z, fd * ad4.z));
        o.w = fmaf(acc[q][3], sc, fmaf(fs, as4.w, fd * ad4.w));
        stg_stream(reinterpret_cast<float4*>(dxw + j * D + e0), o);
    }
}

template <class GE, bool CONCAT>
__global__ void __launch_bounds__(ROW_THREADS)
gat_bwd_src_rows(const int32_t* __restrict__ colptr, const int32_t* __restrict__ csc_row,
                 const int32_t* __restrict__ csc_eid, const float* __restrict__ alpha_used,
                 const float* __restrict__ dz, const float* __restrict__ d_out, const float* __restrict__ att_src,
                 const float* __restrict__ att_dst, const float* __restrict__ da_dst_full, int64_t n_src,
                 int hub_threshold, float* __restrict__ dxw, float* __restrict__ da_src)
{
    constexpr int H = GE::H, NS = GE::NS;
    __shared__ __align__(16) float p_sh[ROW_WARPS][32 * H];
    __shared__ int i_sh[ROW_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t j = int64_t(blockIdx.x) * ROW_WARPS + warp;
    if (j >= n_src) return;
    const int beg = colptr[j], end = colptr[j + 1];
    if (end - beg > hub_threshold) return;
    float acc[NS][4], das[H];
#pragma unroll
    for (int q = 0; q < NS; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[q][k] = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) das[h] = 0.f;
    src_range<GE, CONCAT>(beg, end, csc_row, csc_eid, alpha_used, dz, d_out, acc, das, p_sh[warp], i_sh[warp], lane);
#pragma unroll
    for (int h = 0; h < H; ++h) das[h] = warp_sum(das[h]);
    src_epilogue<GE, CONCAT>(j, acc, das, att_src, att_dst, da_dst_full, dxw, da_src, lane);
}

template <class GE, bool CONCAT>
__global__ void __launch_bounds__(ROW_THREADS)
gat_bwd_src_hub_chunks(const int32_t* __restrict__ colptr, const int32_t* __restrict__ csc_row,
                       const int32_t* __restrict__ csc_eid, const float* __restrict__ alpha_used,
                       const float* __restrict__ dz, const float* __restrict__ d_out, gnnfd_hub_plan_t plan,
                       float* __restrict__ part_acc, float* __restrict__ part_das)
{
    constexpr int H = GE::H, NS = GE::NS, D = GE::D;
    __shared__ __align__(16) float p_sh[ROW_WARPS][32 * H];
    __shared__ int i_sh[ROW_WARPS][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * ROW_WARPS + warp;
    if (c >= plan.n_chunk) return;
    const int slot = plan.chunk_hub[c];
    const int64_t j = plan.hub_row[slot];
    const int beg = colptr[j] + (c - plan.hub_chunk_ptr[slot]) * plan.chunk;
    const int end = min(colptr[j + 1], beg + plan.chunk);
    float acc[NS][4], das[H];
#pragma unroll
    for (int q = 0; q < NS; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[q][k] = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) das[h] = 0.f;
    src_range<GE, CONCAT>(beg, end, csc_row, csc_eid, alpha_used, dz, d_out, acc, das, p_sh[warp], i_sh[warp], lane);
#pragma unroll
    for (int h = 0; h < H; ++h) das[h] = warp_sum(das[h]);
    if (lane == 0) store_vecH<H>(part_das + int64_t(c) * H, das);
#pragma unroll
    for (int q = 0; q < NS; ++q)
        *reinterpret_cast<float4*>(part_acc + int64_t(c) * D + 4 * (lane + 32 * q)) =
            make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
}

template <class GE, bool CONCAT>
__global__ void __launch_bounds__(ROW_THREADS)
gat_bwd_src_hub_merge(gnnfd_hub_plan_t plan, const float* __restrict__ part_acc, const float* __restrict__ part_das,
                      const float* __restrict__ att_src, const float* __restrict__ att_dst,
                      const float* __restrict__ da_dst_full, float* __restrict__ dxw, float* __restrict__ da_src)
{
    constexpr int H = GE::H, NS = GE::NS, D = GE::D;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int slot = blockIdx.x * ROW_WARPS + warp;
    if (slot >= plan.n_hub) return;
    const int64_t j = plan.hub_row[slot];
    float acc[NS][4], das[H];
#pragma unroll
    for (int q = 0; q < NS; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[q][k] = 0.f;
#pragma unroll
    for (int h = 0; h < H; ++h) das[h] = 0.f;
    for (int c = plan.hub_chunk_ptr[slot]; c < plan.hub_chunk_ptr[slot + 1]; ++c) {
        float pd[H];
        load_vecH<H>(part_das + int64_t(c) * H, pd);
#pragma unroll
        for (int h = 0; h < H; ++h) das[h] += pd[h];
#pragma unroll
        for (int q = 0; q < NS; ++q) {
            const float4 v = *reinterpret_cast<const float4*>(part_acc + int64_t(c) * D + 4 * (lane + 32 * q));
            acc[q][0] += v.x; acc[q][1] += v.y; acc[q][2] += v.z; acc[q][3] += v.w;
        }
    }
    src_epilogue<GE, CONCAT>(j, acc, das, att_src, att_dst, da_dst_full, dxw, da_src, lane);
}

// ---------------------------------------------------------------------------------------------------------
template <class GE>
static int launch_bwd_dst(const gnnfd_graph_t* g, const void* xw_, const float* a_src, const float* a_dst,
                          const float* rowmax, const float* rowsum, const float* d_out, float slope, int concat,
                          const uint8_t* keep, float p_drop, float* alpha_used, float* dz, float* da_dst, void* ws,
                          size_t ws_bytes, cudaStream_t st)
{
    using XT = typename GE::XT;
    const XT* xw = reinterpret_cast<const XT*>(xw_);
    const int64_t n = g->n_dst;
    if (n == 0) return GNNFD_OK;
    const bool drop = keep != nullptr && p_drop > 0.f;
    const float ks = drop ? 1.f / (1.f - p_drop) : 1.f;
    const int thr = g->hub_dst.n_hub > 0 ? g->hub_dst.threshold : INT_MAX;
    const unsigned grid = (unsigned)((n + ROW_WARPS - 1) / ROW_WARPS);
#define GNNFD_BWD_ROWS(CC, DD)                                                                                        \
    gat_bwd_dst_rows<GE, CC, DD><<<grid, ROW_THREADS, 0, st>>>(g->rowptr, g->col, g->perm, xw, a_src, a_dst, rowmax,    \
                                                               rowsum, d_out, n, thr, slope, keep, ks, alpha_used, dz, \
                                                               da_dst)
    if (concat) { if (drop) GNNFD_BWD_ROWS(true, true); else GNNFD_BWD_ROWS(true, false); }
    else        { if (drop) GNNFD_BWD_ROWS(false, true); else GNNFD_BWD_ROWS(false, false); }
#undef GNNFD_BWD_ROWS
    g_launches += 1;
    if (g->hub_dst.n_hub > 0) {
        const gnnfd_hub_plan_t& pl = g->hub_dst;
        const size_t need = 2 * carve_bytes(size_t(pl.n_chunk) * GE::H, 4);
        GNNFD_REQUIRE(ws && ws_bytes >= need, GNNFD_ERR_WORKSPACE, "gat_bwd_dst: workspace %zu < %zu", ws_bytes, need);
        char* p = reinterpret_cast<char*>(ws);
        float* part_t = carve<float>(p, size_t(pl.n_chunk) * GE::H);
        float* part_dad = carve<float>(p, size_t(pl.n_chunk) * GE::H);
        const unsigned gc = (unsigned)((pl.n_chunk + ROW_WARPS - 1) / ROW_WARPS);
#define GNNFD_BWD_HUB1(CC, DD)                                                                                        \
    gat_bwd_dst_hub1<GE, CC, DD><<<gc, ROW_THREADS, 0, st>>>(g->rowptr, g->col, g->perm, xw, a_src, a_dst, rowmax,      \
                                                             rowsum, d_out, pl, slope, keep, ks, alpha_used, dz, part_t)
        if (concat) { if (drop) GNNFD_BWD_HUB1(true, true); else GNNFD_BWD_HUB1(true, false); }
        else        { if (drop) GNNFD_BWD_HUB1(false, true); else GNNFD_BWD_HUB1(false, false); }
#undef GNNFD_BWD_HUB1
        gat_bwd_dst_hub2<GE><<<gc, ROW_THREADS, 0, st>>>(g->rowptr, g->col, a_src, a_dst, rowmax, rowsum, pl, slope,
                                                         part_t, dz, part_dad);
        gat_bwd_dst_hub3<GE::H><<<(unsigned)((pl.n_hub * GE::H + 255) / 256), 256, 0, st>>>(pl, part_dad, da_dst);
        g_launches += 3;
    }
    GNNFD_LAUNCH_CHECK();
    return GNNFD_OK;
}

template <class GE>
static int launch_bwd_src(const gnnfd_graph_t* g, const float* alpha_used, const float* dz, const float* d_out,
                          const float* att_src, const float* att_dst, const float* da_dst_full, int concat, float* dxw,
                          float* da_src, void* ws, size_t ws_bytes, cudaStream_t st)
{
    const int64_t n = g->n_src;
    if (n == 0) return GNNFD_OK;
    const int thr = g->hub_src.n_hub > 0 ? g->hub_src.threshold : INT_MAX;
    const unsigned grid = (unsigned)((n + ROW_WARPS - 1) / ROW_WARPS);
    if (concat)
        gat_bwd_src_rows<GE, true><<<grid, ROW_THREADS, 0, st>>>(g->colptr, g->csc_row, g->csc_eid, alpha_used, dz, d_out,
                                                                 att_src, att_dst, da_dst_full, n, thr, dxw, da_src);
    else
        gat_bwd_src_rows<GE, false><<<grid, ROW_THREADS, 0, st>>>(g->colptr, g->csc_row, g->csc_eid, alpha_used, dz, d_out,
                                                                  att_src, att_dst, da_dst_full, n, thr, dxw, da_src);
    g_launches += 1;
    if (g->hub_src.n_hub > 0) {
        const gnnfd_hub_plan_t& pl = g->hub_src;
        const size_t need = carve_bytes(size_t(pl.n_chunk) * GE::D, 4) + carve_bytes(size_t(pl.n_chunk) * GE::H, 4);
        GNNFD_REQUIRE(ws && ws_bytes >= need, GNNFD_ERR_WORKSPACE, "gat_bwd_src: workspace %zu < %zu", ws_bytes, need);
        char* p = reinterpret_cast<char*>(ws);
        float* part_acc = carve<float>(p, size_t(pl.n_chunk) * GE::D);
        float* part_das = carve<float>(p, size_t(pl.n_chunk) * GE::H);
        const unsigned gc = (unsigned)((pl.n_chunk + ROW_WARPS - 1) / ROW_WARPS);
        const unsigned gh = (unsigned)((pl.n_hub + ROW_WARPS - 1) / ROW_WARPS);
        if (concat) {
            gat_bwd_src_hub_chunks<GE, true><<<gc, ROW_THREADS, 0, st>>>(g->colptr, g->csc_row, g->csc_eid, alpha_used, dz,
                                                                         d_out, pl, part_acc, part_das);
            gat_bwd_src_hub_merge<GE, true><<<gh, ROW_THREADS, 0, st>>>(pl, part_acc, part_das, att_src, att_dst,
                                                                        da_dst_full, dxw, da_src);
        } else {
            gat_bwd_src_hub_chunks<GE, false><<<gc, ROW_THREADS, 0, st>>>(g->colptr, g->csc_row, g->csc_eid, alpha_used, dz,
                                                                          d_out, pl, part_acc, part_das);
            gat_bwd_src_hub_merge<GE, false><<<gh, ROW_THREADS, 0, st>>>(pl, part_acc, part_das, att_src, att_dst,
                                                                         da_dst_full, dxw, da_src);
        }
        g_launches += 2;
    }
    GNNFD_LAUNCH_CHECK();
    return GNNFD_OK;
}

}  // namespace gnnfd

using namespace gnnfd;

extern "C" {

int gnnfd_gat_bwd_workspace_bytes(const gnnfd_graph_t* g, int H, int C, size_t* bytes)
{
    GNNFD_REQUIRE(g && bytes, GNNFD_ERR_ARG, "gat_bwd_workspace_bytes: NULL argument");
    const size_t nd = (size_t)g->hub_dst.n_chunk, ns = (size_t)g->hub_src.n_chunk;
    const size_t a = 2 * carve_bytes(nd * H, 4);
    const size_t b = carve_bytes(ns * size_t(H) * C, 4) + carve_bytes(ns * H, 4);
    *bytes = (a > b ? a : b) + 256;
    return GNNFD_OK;
}

int gnnfd_gat_bwd_dst(const gnnfd_graph_t* g, const void* xw, int xw_dtype, const float* a_src, const float* a_dst,
                      const float* rowmax, const float* rowsum, const float* d_out, int H, int C,
                      float negative_slope, int concat, const uint8_t* keep_mask, float p_drop, float* alpha_used,
                      float* dz, float* da_dst, void* ws, size_t ws_bytes, gnnfd_stream_t stream)
{
    int rc = check_graph(g, false, "gat_bwd_dst");
    if (rc) return rc;
    GNNFD_REQUIRE(g->n_dst == 0 || (xw && a_src && a_dst && rowmax && rowsum && d_out && da_dst), GNNFD_ERR_ARG,
                  "gat_bwd_dst: NULL tensor");
    GNNFD_REQUIRE(g->n_edges == 0 || (alpha_used && dz), GNNFD_ERR_ARG, "gat_bwd_dst: alpha_used/dz is NULL");
    GNNFD_REQUIRE(p_drop >= 0.f && p_drop < 1.f, GNNFD_ERR_ARG, "gat_bwd_dst: dropout p must be in [0,1)");
    cudaStream_t st = (cudaStream_t)stream;
    if (H == 8 && C == 64 && xw_dtype == GNNFD_F32)
        return launch_bwd_dst<Geo<8, 64, float>>(g, xw, a_src, a_dst, rowmax, rowsum, d_out, negative_slope, concat,
                                                 keep_mask, p_drop, alpha_used, dz, da_dst, ws, ws_bytes, st);
    if (H == 8 && C == 64 && xw_dtype == GNNFD_BF16)
        return launch_bwd_dst<Geo<8, 64, __nv_bfloat16>>(g, xw, a_src, a_dst, rowmax, rowsum, d_out, negative_slope,
                                                         concat, keep_mask, p_drop, alpha_used, dz, da_dst, ws, ws_bytes, st);
    if (H == 4 && C == 32 && xw_dtype == GNNFD_F32)
        return launch_bwd_dst<Geo<4, 32, float>>(g, xw, a_src, a_dst, rowmax, rowsum, d_out, negative_slope, concat,
                                                 keep_mask, p_drop, alpha_used, dz, da_dst, ws, ws_bytes, st);
    GNNFD_REQUIRE(false, GNNFD_ERR_UNSUPPORTED, "gat_bwd_dst: (heads=%d, out_channels=%d, dtype=%d) is not built", H, C,
                  xw_dtype);
    return GNNFD_ERR_UNSUPPORTED;
}

int gnnfd_gat_bwd_src(const gnnfd_graph_t* g, const float* alpha_used, const float* dz, const float* d_out,
                      const float* att_src, const float* att_dst, const float* da_dst_full, int H, int C, int concat,
                      float* dxw, float* da_src, void* ws, size_t ws_bytes, gnnfd_stream_t stream)
{
    int rc = check_graph(g, true, "gat_bwd_src");
    if (rc) return rc;
    GNNFD_REQUIRE(g->n_src == 0 || (att_src && att_dst && dxw && da_src), GNNFD_ERR_ARG, "gat_bwd_src: NULL tensor");
    GNNFD_REQUIRE(g->n_edges == 0 || (alpha_used && dz && d_out), GNNFD_ERR_ARG, "gat_bwd_src: NULL edge tensor");
    cudaStream_t st = (cudaStream_t)stream;
    if (H == 8 && C == 64)
        return launch_bwd_src<Geo<8, 64, float>>(g, alpha_used, dz, d_out, att_src, att_dst, da_dst_full, concat, dxw,
                                                 da_src, ws, ws_bytes, st);
    if (H == 4 && C == 32)
        return launch_bwd_src<Geo<4, 32, float>>(g, alpha_used, dz, d_out, att_src, att_dst, da_dst_full, concat, dxw,
                                                 da_src, ws, ws_bytes, st);
    GNNFD_REQUIRE(false, GNNFD_ERR_UNSUPPORTED, "gat_bwd_src: (heads=%d, out_channels=%d) is not built", H, C);
    return GNNFD_ERR_UNSUPPORTED;
}

}  // extern "C"
