// kernels_tc.cuh — tcgen05 (5th-generation tensor core) path of the EM hot
// path.  Interface used by gmm_api.cu; implementation in kernels_tc.cu.
#pragma once
#include <cuda_runtime.h>
#include "../../include/gmm.h"

namespace gmm {

struct TcState;

// Shapes the tensor-core kernels cover (independently for the two steps).
bool tc_mstep_supported(int D, int K);
bool tc_estep_supported(int D, int K);

// memb_pitch: row pitch (in floats) of the cluster-major responsibilities buffer AND of the SoA event copy.
int  tc_create(TcState** out, const float* d_x_aos, const float* d_x_soa, int n, int D, int Kmax, float* d_memb, size_t memb_pitch,
               int num_sms, cudaStream_t stream);
void tc_destroy(TcState*);
void tc_set_host_threads(TcState*, int n);
// M-step responsibilities as an FP16 hi/lo pair (1), as one round-to-nearest FP16 value (0), or chosen per launch
// from the smallest cluster size the caller passes to tc_launch_mstep (2, default: pair below 2048 events).
void tc_set_gamma_split(TcState*, int mode);
// Centre/scale used inside the tensor kernels: z = (x - shift) * inv_scale, both rounded to
// float; `shift` is updated in place to the float-rounded values actually used.
int  tc_set_shift_scale(TcState*, double* shift, const double* scale, cudaStream_t stream);
int  tc_upload_params(TcState*, const clusters_t* host, int K, cudaStream_t stream);
// The same in three steps, so that the caller can fuse the per-cluster work with its own per-cluster
// finalisation in ONE parallel loop: begin (serial), cluster k in [0, tc_params_padded) (independent, thread
// safe; returns 0 or a defect code to be max-reduced), commit (serial: error report or H2D of the operand).
int  tc_params_begin(TcState*, int K, cudaStream_t stream);
int  tc_params_padded(const TcState*, int K);
int  tc_params_cluster(TcState*, const clusters_t* host, int k, int K);
int  tc_params_commit(TcState*, int K, int bad, cudaStream_t stream);
int  tc_launch_estep(TcState*, int K, double* d_ll, cudaStream_t stream);
// Accumulates sum_n g[k][n] * phi_f(x_n - shift) into d_stats[k*F + f] (double, original units).
// min_nk: smallest N_k of the current parameters (global, all ranks); only read under gamma-split mode 2.
// *pair_out (optional): 1 when the FP16-pair kernel (three products) ran, 0 for the single-FP16 one (two products).
int  tc_launch_mstep(TcState*, int K, double* d_stats, cudaStream_t stream, float min_nk, int* pair_out = nullptr);
// Zeroes the per-CTA scratch of the last tc_launch_mstep for the next one; meant to be enqueued behind the D2H copy
// of the statistics so that it runs while the host finalises.
int  tc_mstep_cleanup(TcState*, cudaStream_t stream);

}  // namespace gmm
