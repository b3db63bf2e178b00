// kernels_tc.cu — tcgen05 path (under construction: reports "unsupported" for
// every shape so that GMM_PATH_AUTO uses the SIMT kernels and GMM_PATH_TENSOR
// fails loudly).
#include "kernels_tc.cuh"
#include "host_math.h"

namespace gmm {
struct TcState { int dummy; };
bool tc_supported(int, int) { return false; }
int tc_create(TcState** out, const float*, int, int, int, int, cudaStream_t) { *out = nullptr; return GMM_OK; }
void tc_destroy(TcState*) {}
int tc_set_shift(TcState*, const double*, cudaStream_t) { return GMM_OK; }
int tc_upload_params(TcState*, const clusters_t*, int, cudaStream_t) { return fail(GMM_ERR_STATE, "tensor path unavailable"); }
int tc_launch_estep(TcState*, int, float*, double*, cudaStream_t) { return fail(GMM_ERR_STATE, "tensor path unavailable"); }
int tc_launch_mstep(TcState*, int, const float*, double*, cudaStream_t) { return fail(GMM_ERR_STATE, "tensor path unavailable"); }
}  // namespace gmm
