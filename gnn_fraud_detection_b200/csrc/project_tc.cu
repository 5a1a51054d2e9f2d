// placeholder until the tcgen05 path lands
#include "common.cuh"
namespace gnnfd {
bool tc_supported(int64_t, int64_t, int, int) { return false; }
size_t tc_ws_bytes(int64_t, int64_t, int, int) { return 0; }
int project_fwd_tc(const float*, int64_t, const float*, const float*, const float*, int64_t, int64_t, int, int, int, void*, float*, float*, void*, size_t, cudaStream_t) { return GNNFD_ERR_UNSUPPORTED; }
int project_bwd_dx_tc(const float*, const float*, int64_t, int64_t, int, float*, int64_t, void*, size_t, cudaStream_t) { return GNNFD_ERR_UNSUPPORTED; }
int project_bwd_dw_tc(const float*, const float*, int64_t, int64_t, int64_t, int, float*, void*, size_t, cudaStream_t) { return GNNFD_ERR_UNSUPPORTED; }
}
