// kernels_tc.cuh — tcgen05 (5th-generation tensor core) path of the EM hot
// path.  Interface used by gmm_api.cu; implementation in kernels_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include "../../include/gmm.h"

namespace gmm {

struct TcState;

// Shapes the tensor-core kernels cover.
bool tc_supported(int D, int K);

int  tc_create(TcState** out, const float* d_x_aos, int n, int D, int Kmax, int num_sms, cudaStream_t stream);
void tc_destroy(TcState*);
int  tc_set_shift(TcState*, const double* shift, cudaStream_t stream);
int  tc_upload_params(TcState*, const clusters_t* host, int K, cudaStream_t stream);
int  tc_launch_estep(TcState*, int K, float* d_memb, double* d_ll, cudaStream_t stream);
int  tc_launch_mstep(TcState*, int K, const float* d_memb, double* d_stats, cudaStream_t stream);

}  // namespace gmm
