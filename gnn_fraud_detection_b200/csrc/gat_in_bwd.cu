// Input-space backward of the first GAT layer, edge pass (see in_common.cuh for the algebra): the autograd mirror of
// edge_update / softmax / aggregate (loss.backward() at src/train.py:142) without projected features:
//     d_alpha[e,h] = Gd[i,h,:] . x[j,:]          Gd = dO W_r^T / H from gnnfd_in_bwd_gd (one 4*F-byte row per destination)
//     u = alpha * d_alpha (dropout-scaled), t = sum_row u, dz = slope' * (u - alpha t), da_dst[i] = sum_row dz
// dz is written in SOURCE-MAJOR order (slot csr2csc[e]) so that da_src = per-source sums (gnnfd_in_bwd_dasrc) reads it
// contiguously.  No alpha_used / dxw: the weight gradient comes from the saved Z image (in_gemm.cu).
// Same warp-stream structure as gat_bwd_dst.cu (work items, phase A one chunk ahead, packs of short rows, hub rows in
// chunks with a deterministic merge); x rows arrive through the per-warp bulk-copy ring, and the destination's Gd row
// is prefetched by the copy engine into a per-warp shared-memory buffer while the previous row is processed.
#include "in_common.cuh"
#include "gat_phase_bwd.cuh"

#include <atomic>
#include <climits>
#include <cstdlib>

namespace gnnfd {
extern std::atomic<long long> g_launches;
int check_graph(const gnnfd_graph_t* g, bool need_csc, const char* who);

namespace in {

// per-warp scratch beyond the ring: dal_s [32][H] floats, bits_s [2][32] ints, the Gd row buffer [F] floats
__host__ __device__ inline int bwd_extra_bytes(int F) { return 32 * H * 4 + 2 * 32 * 4 + F * 4; }

// Gd row of the destination being processed: fetched with one bulk copy into shared memory (no registers in flight)
struct GdBuf {
    float* buf;
    uint64_t* bar;
    const float* gd;
    int F, KP;
    int pending = -1, loads = 0;
    __device__ __forceinline__ void want(int row, int lane)
    {
        if (pending == row) return;
        if (lane == 0) {
            st_mbar_expect_tx(bar, uint32_t(F) * 4u);
            st_bulk_g2s(buf, gd + int64_t(row) * F, uint32_t(F) * 4u, bar);
        }
        pending = row;
    }
    __device__ __forceinline__ void take(int row, int lane, float2 (&g)[H][NSLOT])
    {
        if (pending != row) {
            if (pending >= 0) {                 // a wrong guess is in flight: let it land before the buffer is reused
                st_mbar_wait(bar, uint32_t(loads & 1));
                ++loads;
                __syncwarp();
                pending = -1;
            }
            want(row, lane);
        }
        st_mbar_wait(bar, uint32_t(loads & 1));
        ++loads;
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
            for (int r = 0; r < NSLOT; ++r) {
                const int k = 64 * r + 2 * lane;
                g[h][r] = (k < KP) ? *reinterpret_cast<const float2*>(buf + h * KP + k) : make_float2(0.f, 0.f);
            }
        __syncwarp();                           // every lane has read the buffer before the next copy is issued
        pending = -1;
    }
};

// 8 per-lane partial dot products -> full warp sums; head (lane>>2)&7 in bit order b4 b3 b2 ends in the lanes of its quad
__device__ __forceinline__ float reduce8(const float (&d)[H], int lane)
{
    const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
    float a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (b4 ? d[4 + i] : d[i]) + __shfl_xor_sync(FULL, b4 ? d[i] : d[4 + i], 16);
#pragma unroll
    for (int i = 0; i < 2; ++i) b[i] = (b3 ? a[2 + i] : a[i]) + __shfl_xor_sync(FULL, b3 ? a[i] : a[2 + i], 8);
    float c = (b2 ? b[1] : b[0]) + __shfl_xor_sync(FULL, b2 ? b[0] : b[1], 4);
    c += __shfl_xor_sync(FULL, c, 2);
    c += __shfl_xor_sync(FULL, c, 1);
    return c;
}

// HUB = false: whole rows (optionally packs of short rows), results go to da_dst.  HUB = true: one (row, range) segment,
// the partial t of the segment goes to part_t[chunk_id]; the second sweep is a separate kernel.
template <bool VEC2, bool DROPOUT, bool PACK, bool HUB>
__device__ __forceinline__ void in_bwd_stream(ChunkCursor& cur, InRing& ring, GdBuf& gdb, const int32_t* __restrict__ rowptr,
                                              const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                                              const int32_t* __restrict__ csr2csc, int K, const float* __restrict__ a_src,
                                              const float* __restrict__ a_dst, const float* __restrict__ rowmax,
                                              const float* __restrict__ rowsum, float slope, const uint8_t* __restrict__ keep,
                                              float keep_scale, float* __restrict__ dz, float* __restrict__ da_dst,
                                              float* __restrict__ part_t, int chunk_id, int lane)
{
    float* dal_s = reinterpret_cast<float*>(ring.extra);
    int* bits_s = reinterpret_cast<int*>(ring.extra + 32 * H * 4);
    auto on_empty = [&](int r) {
        if (!HUB && lane < H) da_dst[int64_t(r) * H + lane] = 0.f;
    };
    BwdChunk c0, c1;
    int kind0 = 0, kind1 = 0, k0 = 0, k1 = 0, la = 0, lb = 0;
    int b0 = 0;
    auto next = [&](BwdChunk& c, int& kk) -> int {
        if (PACK) return cur.next_any(rowptr, lane, c.row, c.beg, c.n, c.first, c.last, kk, la, lb, on_empty);
        return cur.next(rowptr, c.row, c.beg, c.n, c.first, c.last, on_empty) ? 1 : 0;
    };
    auto phase_a = [&](const BwdChunk& c, int kind, int kk, int buf) {
        if (PACK && kind == 2)   // CONCAT = true: the dO-row prefetch of the projected-feature kernel does not apply here
            bwd_phase_a_pack<GI, true, DROPOUT>(c.row, c.beg, c.n, kk, la, lb, col, perm, a_src, a_dst, rowmax, rowsum, nullptr,
                                                slope, keep, ring.p_s + buf * 32 * H, ring.j_s + buf * 32, bits_s + buf * 32, lane);
        else
            bwd_phase_a<GI, DROPOUT>(c, col, perm, a_src, a_dst, rowmax, rowsum, slope, keep, ring.p_s + buf * 32 * H,
                                     ring.j_s + buf * 32, bits_s + buf * 32, lane);
    };
    kind0 = next(c0, k0);
    if (!kind0) return;
    gdb.want(c0.row, lane);
    phase_a(c0, kind0, k0, b0);
    float2 g[H][NSLOT];
    gdb.take(c0.row, lane, g);
    if (PACK && kind0 == 2 && k0 > 1) gdb.want(c0.row + 1, lane);
    int issued0 = 0, issued1 = 0;
    float trow[H];
    int row_beg = c0.beg;
#pragma unroll
    for (int h = 0; h < H; ++h) trow[h] = 0.f;
    while (true) {
        const int* j0 = ring.j_s + b0 * 32;
        const int* j1 = ring.j_s + (b0 ^ 1) * 32;
        while (ring.has_room() && issued0 < c0.n) ring.issue(j0[issued0++], lane);
        kind1 = next(c1, k1);
        issued1 = 0;
        if (kind1) {
            phase_a(c1, kind1, k1, b0 ^ 1);
            // the next destination row in stream order, unless a row of the current pack is still to come
            if (c1.first && gdb.pending < 0 && !(PACK && kind0 == 2 && k0 > 1)) gdb.want(c1.row, lane);
        }
        if (c0.first) {
            row_beg = c0.beg;
#pragma unroll
            for (int h = 0; h < H; ++h) trow[h] = 0.f;
        }
        int rows_done = 0;
        const int* bits0 = bits_s + b0 * 32;
        // phase B: dot products of the staged x rows with this row's Gd
        for (int t = 0; t < c0.n; ++t) {
            const float* row = ring.front(j0[t]);
            float2 v[NSLOT];
            load_xrow<VEC2>(row, lane, K, v);
            float d[H];
#pragma unroll
            for (int h = 0; h < H; ++h) {
                float s = 0.f;
#pragma unroll
                for (int r = 0; r < NSLOT; ++r) s = fmaf(g[h][r].x, v[r].x, fmaf(g[h][r].y, v[r].y, s));
                d[h] = s;
            }
            const float tot = reduce8(d, lane);
            if ((lane & 3) == 0) dal_s[t * H + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = tot;
            ring.pop();
            if (issued0 < c0.n) ring.issue(j0[issued0++], lane);
            else if (kind1 && issued1 < c1.n) ring.issue(j1[issued1++], lane);
            if (PACK && kind0 == 2 && t + 1 < c0.n && ((bits0[t] >> 26) & 1)) {
                // row boundary inside the pack: switch to the next row's Gd, prefetch the one after it
                ++rows_done;
                gdb.take(c0.row + rows_done, lane, g);
                if (rows_done + 1 < k0) gdb.want(c0.row + rows_done + 1, lane);
                else if (kind1 && c1.first) gdb.want(c1.row, lane);
            }
        }
        __syncwarp();
        // phase C: lane = edge
        {
            float alpha[H], dal[H], u[H];
            const float* p0 = ring.p_s + b0 * 32 * H;
            const int bits = bits0[lane];
#pragma unroll
            for (int kk = 0; kk < H / 4; ++kk) {
                const float4 a4 = *reinterpret_cast<const float4*>(p0 + lane * H + 4 * kk);
                const float4 d4 = *reinterpret_cast<const float4*>(dal_s + lane * H + 4 * kk);
                alpha[4 * kk] = a4.x; alpha[4 * kk + 1] = a4.y; alpha[4 * kk + 2] = a4.z; alpha[4 * kk + 3] = a4.w;
                dal[4 * kk] = d4.x; dal[4 * kk + 1] = d4.y; dal[4 * kk + 2] = d4.z; dal[4 * kk + 3] = d4.w;
            }
            const bool live = lane < c0.n;
            const int64_t pos = live ? csr2csc[c0.beg + lane] : 0;    // source-major slot of this edge
#pragma unroll
            for (int h = 0; h < H; ++h) {
                const float ks = (bits >> (8 + h)) & 1 ? keep_scale : 0.f;
                u[h] = live ? alpha[h] * dal[h] * ks : 0.f;
            }
            if (PACK && kind0 == 2) {
                const int sa = (bits >> 16) & 31, sb = live ? (bits >> 21) & 31 : lane;
                const bool lastf = live && ((bits >> 26) & 1);
                float o[H], dad[H];
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float tt = seg_total(u[h], sa, sb, lane);
                    const float sl = (bits >> h) & 1 ? 1.f : slope;
                    o[h] = live ? sl * (u[h] - alpha[h] * tt) : 0.f;
                    dad[h] = seg_total(o[h], sa, sb, lane);
                }
                if (live) store_vecH<H>(dz + pos * H, o);
                const unsigned lastm = __ballot_sync(FULL, lastf);
                if (lastf) {
                    const int64_t r = c0.row + __popc(lastm & ((1u << lane) - 1u));
                    store_vecH<H>(da_dst + r * H, dad);
                }
            } else if (!HUB && c0.first && c0.last) {
                float o[H], dad[H];
#pragma unroll
                for (int h = 0; h < H; ++h) {
                    const float tt = warp_sum(u[h]);
                    const float sl = (bits >> h) & 1 ? 1.f : slope;
                    o[h] = live ? sl * (u[h] - alpha[h] * tt) : 0.f;
                    dad[h] = warp_sum(o[h]);
                }
                if (live) store_vecH<H>(dz + pos * H, o);
                if (lane == 0) store_vecH<H>(da_dst + int64_t(c0.row) * H, dad);
            } else {
                if (live) store_vecH<H>(dz + pos * H, u);       // parked until t is known
#pragma unroll
                for (int h = 0; h < H; ++h) trow[h] += warp_sum(u[h]);
                if (c0.last) {
                    if (HUB) {
                        if (lane == 0) store_vecH<H>(part_t + int64_t(chunk_id) * H, trow);
                    } else {
                        __syncwarp();
                        RowStat<H> r;
                        load_row_stat<H>(r, c0.row, a_dst, rowmax, rowsum);
                        float dad[H];
#pragma unroll
                        for (int h = 0; h < H; ++h) dad[h] = 0.f;
                        dst_sweep2<GI>(r, row_beg, c0.beg + c0.n, col, csr2csc, a_src, slope, trow, lane, dz, H, dad);
#pragma unroll
                        for (int h = 0; h < H; ++h) dad[h] = warp_sum(dad[h]);
                        if (lane == 0) store_vecH<H>(da_dst + int64_t(c0.row) * H, dad);
                    }
                }
            }
        }
        __syncwarp();
        if (!kind1) break;
        if (c1.first) {
            gdb.take(c1.row, lane, g);
            if (PACK && kind1 == 2 && k1 > 1) gdb.want(c1.row + 1, lane);
        }
        c0 = c1;
        kind0 = kind1;
        k0 = k1;
        issued0 = issued1;
        b0 ^= 1;
    }
}

__device__ __forceinline__ void gdbuf_init(GdBuf& gdb, InRing& ring, const float* gd, int F, int KP)
{
    gdb.buf = reinterpret_cast<float*>(ring.extra + 32 * H * 4 + 2 * 32 * 4);
    gdb.bar = &ring.full[IN_R];
    gdb.gd = gd;
    gdb.F = F;
    gdb.KP = KP;
}

template <bool VEC2, bool DROPOUT, bool PACK>
__global__ void __launch_bounds__(IN_THREADS, 3)
gat_in_bwd_items(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                 const int32_t* __restrict__ csr2csc, const float* __restrict__ x, int64_t ldx, int K,
                 const float* __restrict__ a_src, const float* __restrict__ a_dst, const float* __restrict__ rowmax,
                 const float* __restrict__ rowsum, const float* __restrict__ gd, gnnfd_item_plan_t items, int item_lo,
                 int item_hi, int hub_threshold, float slope, const uint8_t* __restrict__ keep, float keep_scale,
                 float* __restrict__ dz, float* __restrict__ da_dst)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int item = item_lo + blockIdx.x * IN_WARPS + warp;
    if (item >= item_hi) return;
    const Dims d(K);
    InRing ring;
    ring.init(smem + warp * in_warp_bytes(K, bwd_extra_bytes(d.F)), x, ldx, K, lane);
    GdBuf gdb;
    gdbuf_init(gdb, ring, gd, d.F, d.KP);
    ChunkCursor cur;
    cur.start_rows(items.item_start[item], items.item_start[item + 1], hub_threshold);
    in_bwd_stream<VEC2, DROPOUT, PACK, false>(cur, ring, gdb, rowptr, col, perm, csr2csc, K, a_src, a_dst, rowmax, rowsum, slope,
                                              keep, keep_scale, dz, da_dst, nullptr, 0, lane);
}

// hub rows, step 1: one warp per chunk -- first sweep, partial t (rows outside [row_lo, row_hi) belong to another block)
template <bool VEC2, bool DROPOUT>
__global__ void __launch_bounds__(IN_THREADS, 3)
gat_in_bwd_hub1(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col, const int32_t* __restrict__ perm,
                const int32_t* __restrict__ csr2csc, const float* __restrict__ x, int64_t ldx, int K,
                const float* __restrict__ a_src, const float* __restrict__ a_dst, const float* __restrict__ rowmax,
                const float* __restrict__ rowsum, const float* __restrict__ gd, gnnfd_hub_plan_t plan, int row_lo, int row_hi,
                float slope, const uint8_t* __restrict__ keep, float keep_scale, float* __restrict__ dz,
                float* __restrict__ part_t)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int c = blockIdx.x * IN_WARPS + warp;
    if (c >= plan.n_chunk) return;
    const int slot = plan.chunk_hub[c];
    const int i = plan.hub_row[slot];
    if (i < row_lo || i >= row_hi) return;
    const int beg = rowptr[i] + (c - plan.hub_chunk_ptr[slot]) * plan.chunk;
    const int end = min(rowptr[i + 1], beg + plan.chunk);
    const Dims d(K);
    InRing ring;
    ring.init(smem + warp * in_warp_bytes(K, bwd_extra_bytes(d.F)), x, ldx, K, lane);
    GdBuf gdb;
    gdbuf_init(gdb, ring, gd, d.F, d.KP);
    ChunkCursor cur;
    cur.start_segment(i, beg, end);
    in_bwd_stream<VEC2, DROPOUT, false, true>(cur, ring, gdb, rowptr, col, perm, csr2csc, K, a_src, a_dst, rowmax, rowsum, slope,
                                              keep, keep_scale, dz, nullptr, part_t, c, lane);
}

template <class Kn>
static int in_set_smem_bwd(Kn kernel, int bytes)
{
    GNNFD_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    return GNNFD_OK;
}

}  // namespace in
}  // namespace gnnfd

using namespace gnnfd;
using namespace gnnfd::in;

extern "C" {

int gnnfd_in_bwd_edges_workspace_bytes(const gnnfd_graph_t* g, size_t* bytes)
{
    GNNFD_REQUIRE(g && bytes, GNNFD_ERR_ARG, "in_bwd_edges_workspace_bytes: NULL argument");
    *bytes = 2 * carve_bytes(size_t(g->hub_dst.n_chunk) * H, 4) + carve_bytes(size_t(g->hub_dst.n_hub) * H, 4) + 256;
    return GNNFD_OK;
}

/* dz [E',H] (source-major order) and da_dst [n_dst,H] for the destination rows of the work items [item_lo, item_hi)
 * (= rows [row_lo, row_hi), which must be the rows those items cover; pass 0, n_items, 0, n_dst for everything).
 * gd holds the Gd rows of [gd_row0, ...) with leading dimension F.  phase bit 0: process the row block (items + first
 * hub sweep); bit 1: finish the hub rows (second sweep; after the LAST block, with the same ws). */
int gnnfd_in_bwd_edges(const gnnfd_graph_t* g, const float* x, int64_t ldx, int64_t K, const float* a_src,
                       const float* a_dst, const float* rowmax, const float* rowsum, const float* gd, int64_t gd_row0,
                       int64_t item_lo, int64_t item_hi, int64_t row_lo, int64_t row_hi, float negative_slope,
                       const uint8_t* keep_mask, float p_drop, float* dz, float* da_dst, void* ws, size_t ws_bytes,
                       int phase, gnnfd_stream_t stream)
{
    int rc = check_graph(g, true, "in_bwd_edges");
    if (rc) return rc;
    GNNFD_REQUIRE(K >= 1 && K <= MAX_K && ldx >= K, GNNFD_ERR_ARG, "in_bwd_edges: bad shape (K <= %d)", MAX_K);
    if (g->n_dst == 0) return GNNFD_OK;
    GNNFD_REQUIRE(x && a_src && a_dst && rowmax && rowsum && gd && da_dst, GNNFD_ERR_ARG, "in_bwd_edges: NULL tensor");
    GNNFD_REQUIRE(g->n_edges == 0 || (dz && g->csr2csc), GNNFD_ERR_ARG, "in_bwd_edges: dz / csr2csc is NULL");
    GNNFD_REQUIRE(p_drop >= 0.f && p_drop <= 0.9f, GNNFD_ERR_ARG, "in_bwd_edges: dropout p must be in [0,0.9]");
    GNNFD_REQUIRE(g->items_dst.n_items > 0 && g->items_dst.item_start, GNNFD_ERR_ARG, "in_bwd_edges: no work-item plan over rowptr");
    GNNFD_REQUIRE(item_lo >= 0 && item_lo <= item_hi && item_hi <= g->items_dst.n_items && row_lo >= 0 && row_lo <= row_hi &&
                      row_hi <= g->n_dst && gd_row0 >= 0 && gd_row0 <= row_lo,
                  GNNFD_ERR_ARG, "in_bwd_edges: bad item / row range");
    GNNFD_REQUIRE((reinterpret_cast<uintptr_t>(gd) & 15) == 0, GNNFD_ERR_ARG, "in_bwd_edges: gd must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const Dims d((int)K);
    const bool drop = keep_mask != nullptr && p_drop > 0.f;
    const float ks = drop ? 1.f / (1.f - p_drop) : 1.f;
    const int thr = g->hub_dst.n_hub > 0 ? g->hub_dst.threshold : INT_MAX;
    const bool v2 = (d.K % 2 == 0) && (ldx % 2 == 0) && (reinterpret_cast<uintptr_t>(x) & 7) == 0;
    const int smem = IN_WARPS * in_warp_bytes(d.K, bwd_extra_bytes(d.F));
    const float* gd0 = gd - gd_row0 * d.F;          // indexed by global destination row
    const gnnfd_hub_plan_t& pl = g->hub_dst;
    float *part_t = nullptr, *part_dad = nullptr, *t_total = nullptr;
    if (pl.n_hub > 0) {
        size_t need = 0;
        gnnfd_in_bwd_edges_workspace_bytes(g, &need);
        GNNFD_REQUIRE(ws && ws_bytes >= need, GNNFD_ERR_WORKSPACE, "in_bwd_edges: workspace %zu < %zu", ws_bytes, need);
        char* p = reinterpret_cast<char*>(ws);
        part_t = carve<float>(p, size_t(pl.n_chunk) * H);
        part_dad = carve<float>(p, size_t(pl.n_chunk) * H);
        t_total = carve<float>(p, size_t(pl.n_hub) * H);
    }
    static const bool pack = [] {
        const char* e = getenv("GNNFD_BWD_PACK");
        return e ? atoi(e) != 0 : true;
    }();
    if ((phase & 1) && item_hi > item_lo) {
        const unsigned grid = (unsigned)((item_hi - item_lo + IN_WARPS - 1) / IN_WARPS);
#define GNNFD_IN_BWD(VV, DD, PP)                                                                                       \
    rc = in_set_smem_bwd(gat_in_bwd_items<VV, DD, PP>, smem);                                                          \
    if (rc) return rc;                                                                                                 \
    gat_in_bwd_items<VV, DD, PP><<<grid, IN_THREADS, smem, st>>>(g->rowptr, g->col, g->perm, g->csr2csc, x, ldx, d.K, a_src, \
                                                                 a_dst, rowmax, rowsum, gd0, g->items_dst, (int)item_lo,   \
                                                                 (int)item_hi, thr, negative_slope, keep_mask, ks, dz, da_dst)
        if (pack) {
            if (v2) { if (drop) { GNNFD_IN_BWD(true, true, true); } else { GNNFD_IN_BWD(true, false, true); } }
            else    { if (drop) { GNNFD_IN_BWD(false, true, true); } else { GNNFD_IN_BWD(false, false, true); } }
        } else {
            if (v2) { if (drop) { GNNFD_IN_BWD(true, true, false); } else { GNNFD_IN_BWD(true, false, false); } }
            else    { if (drop) { GNNFD_IN_BWD(false, true, false); } else { GNNFD_IN_BWD(false, false, false); } }
        }
#undef GNNFD_IN_BWD
        g_launches += 1;
        if (pl.n_hub > 0) {
            const unsigned gc = (unsigned)((pl.n_chunk + IN_WARPS - 1) / IN_WARPS);
#define GNNFD_IN_HUB1(VV, DD)                                                                                          \
    rc = in_set_smem_bwd(gat_in_bwd_hub1<VV, DD>, smem);                                                               \
    if (rc) return rc;                                                                                                 \
    gat_in_bwd_hub1<VV, DD><<<gc, IN_THREADS, smem, st>>>(g->rowptr, g->col, g->perm, g->csr2csc, x, ldx, d.K, a_src, a_dst,  \
                                                          rowmax, rowsum, gd0, pl, (int)row_lo, (int)row_hi, negative_slope, \
                                                          keep_mask, ks, dz, part_t)
            if (v2) { if (drop) { GNNFD_IN_HUB1(true, true); } else { GNNFD_IN_HUB1(true, false); } }
            else    { if (drop) { GNNFD_IN_HUB1(false, true); } else { GNNFD_IN_HUB1(false, false); } }
#undef GNNFD_IN_HUB1
            g_launches += 1;
        }
    }
    if ((phase & 2) && pl.n_hub > 0) {
        const unsigned gh = (unsigned)((pl.n_hub + ROW_WARPS - 1) / ROW_WARPS);
        const unsigned gc2 = (unsigned)((pl.n_chunk + ROW_WARPS - 1) / ROW_WARPS);
        gat_hub_chunk_sum<H, false><<<gh, ROW_THREADS, 0, st>>>(pl, part_t, t_total);
        gat_bwd_dst_hub2<GI><<<gc2, ROW_THREADS, 0, st>>>(g->rowptr, g->col, g->csr2csc, a_src, a_dst, rowmax, rowsum, pl,
                                                          negative_slope, t_total, dz, H, part_dad);
        gat_hub_chunk_sum<H, true><<<gh, ROW_THREADS, 0, st>>>(pl, part_dad, da_dst);
        g_launches += 3;
    }
    GNNFD_LAUNCH_CHECK();
    return GNNFD_OK;
}

}  // extern "C"
