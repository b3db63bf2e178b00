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
// Centre/scale used inside the tensor kernels: z = (x - shift) * inv_scale, both rounded to
// float; `shift` is updated in place to the float-rounded values actually used.  xmin / xmax: per-dimension extremes
// of the WHOLE data set (all ranks): they fix the power-of-two quanta of the M-step's fixed-point operand parts.
int  tc_set_shift_scale(TcState*, double* shift, const double* scale, const double* xmin, const double* xmax, cudaStream_t stream);
// True once the quanta are set and the data range fits the fixed-point budget (|z| <= 64 global standard deviations);
// otherwise the caller uses the FP64 SIMT M-step.
bool tc_mstep_ready(const TcState*);
// False when an event lies beyond 2^14 global standard deviations (its standardised coordinates would overflow FP16).
bool tc_estep_range_ok(const TcState*);
int  tc_upload_params(TcState*, const clusters_t* host, int K, cudaStream_t stream);
// The same in three steps, so that the caller can fuse the per-cluster work with its own per-cluster
// finalisation in ONE parallel loop: begin (serial), cluster k in [0, tc_params_padded) (independent, thread
// safe; returns 0 or a defect code to be max-reduced), commit (serial: error report or H2D of the operand).
int  tc_params_begin(TcState*, int K, cudaStream_t stream);
int  tc_params_padded(const TcState*, int K);
int  tc_params_cluster(TcState*, const clusters_t* host, int k, int K);
// The same with the upper-triangular factor W (Rinv = W^T W, double, row-major [D][D]) supplied by the caller
// (constants_cluster_spd): no second factorisation of the inverse.
int  tc_params_cluster_w(TcState*, const clusters_t* host, int k, int K, const double* W);
int  tc_params_commit(TcState*, int K, int bad, cudaStream_t stream);
int  tc_launch_estep(TcState*, int K, double* d_ll, cudaStream_t stream);
// Accumulates sum_n g[k][n] * phi_f(x_n - shift) into d_stats[k*F + f] (double, original units).
int  tc_launch_mstep(TcState*, int K, double* d_stats, cudaStream_t stream);

}  // namespace gmm
