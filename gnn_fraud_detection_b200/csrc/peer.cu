// Multi-GPU exchange over peer memory (include/gnnfd_b200.h section 7): reductions that read the other ranks' buffers through
// NVLink -- either peer loads summed in rank order, or multimem.ld_reduce, where the NVSwitch adds the G copies and returns
// one value.  The producing side of the exchange is fused into the producing kernels (in_logits_kernel stores its rows on
// every rank); these are the consuming side.
#include <atomic>

#include "common.cuh"

namespace gnnfd {
extern std::atomic<long long> g_launches;

struct PeerBufs {
    int n;
    const float* ptr[GNNFD_MAX_PEERS];
};

template <int OP>
__device__ __forceinline__ float comb(float a, float b) { return OP == 0 ? a + b : fmaxf(a, b); }

// peer loads, fixed rank order (deterministic); float4 body + scalar tail
template <int OP>
__global__ void peer_pull_reduce_kernel(PeerBufs bufs, int64_t offset, int64_t n, float* __restrict__ out)
{
    const bool vec = (offset & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    const int64_t n4 = vec ? n >> 2 : 0;
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += stride) {
        float4 a = __ldcv(reinterpret_cast<const float4*>(bufs.ptr[0] + offset) + i);
        for (int g = 1; g < bufs.n; ++g) {
            const float4 b = __ldcv(reinterpret_cast<const float4*>(bufs.ptr[g] + offset) + i);
            a.x = comb<OP>(a.x, b.x); a.y = comb<OP>(a.y, b.y); a.z = comb<OP>(a.z, b.z); a.w = comb<OP>(a.w, b.w);
        }
        reinterpret_cast<float4*>(out)[i] = a;
    }
    for (int64_t i = n4 * 4 + blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n; i += stride) {
        float a = __ldcv(bufs.ptr[0] + offset + i);
        for (int g = 1; g < bufs.n; ++g) a = comb<OP>(a, __ldcv(bufs.ptr[g] + offset + i));
        out[i] = a;
    }
}

// in-switch reduction: one multimem.ld_reduce returns the sum over all ranks of 16 bytes
__global__ void peer_mc_reduce_kernel(const float* __restrict__ mc, int64_t offset, int64_t n4, float4* __restrict__ out)
{
    const int64_t stride = int64_t(gridDim.x) * blockDim.x;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < n4; i += stride) {
        float4 v;
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                     : "l"(mc + offset + 4 * i)
                     : "memory");
        out[i] = v;
    }
}

}  // namespace gnnfd

using namespace gnnfd;

extern "C" {

int gnnfd_peer_reduce(const gnnfd_peers_t* buf, int64_t offset, int64_t n, int op, int use_multicast, float* out,
                      gnnfd_stream_t stream)
{
    GNNFD_REQUIRE(buf && out && offset >= 0 && n >= 0 && (op == 0 || op == 1), GNNFD_ERR_ARG, "peer_reduce: bad argument");
    GNNFD_REQUIRE(buf->n_peers >= 1 && buf->n_peers <= GNNFD_MAX_PEERS, GNNFD_ERR_ARG, "peer_reduce: 1..%d peers", GNNFD_MAX_PEERS);
    if (n == 0) return GNNFD_OK;
    cudaStream_t st = (cudaStream_t)stream;
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > int64_t(sm_count()) * 8) blocks = int64_t(sm_count()) * 8;
    if (blocks < 1) blocks = 1;
    if (use_multicast) {
        GNNFD_REQUIRE(op == 0 && buf->multicast && (offset & 3) == 0 && (n & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0,
                      GNNFD_ERR_ARG, "peer_reduce: the multicast path needs op = sum, a multicast mapping, offset/n multiples of 4");
        peer_mc_reduce_kernel<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float*>(buf->multicast), offset, n / 4,
                                                               reinterpret_cast<float4*>(out));
    } else {
        PeerBufs b{};
        b.n = buf->n_peers;
        for (int g = 0; g < b.n; ++g) {
            GNNFD_REQUIRE(buf->ptr[g], GNNFD_ERR_ARG, "peer_reduce: peer buffer %d is NULL", g);
            b.ptr[g] = reinterpret_cast<const float*>(buf->ptr[g]);
        }
        if (op == 0) peer_pull_reduce_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(b, offset, n, out);
        else peer_pull_reduce_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(b, offset, n, out);
    }
    g_launches += 1;
    GNNFD_LAUNCH_CHECK();
    return GNNFD_OK;
}

}  // extern "C"
