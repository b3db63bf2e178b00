// host_math.h — GPU-free host numerics of the engine (internal C++ API).
// The extern "C" wrappers of include/gmm.h live in host_math.cpp.
#pragma once
#include <cstddef>
#include <functional>
#include <string>
#include <vector>
#include "../../include/gmm.h"

namespace gmm {

void set_error(const std::string& msg);          // thread-local, read by gmm_last_error()
int  fail(int code, const std::string& msg);     // set_error + return code

// Number of per-cluster sufficient statistics: 1 + D + D(D+1)/2.
inline int num_features(int D) { return 1 + D + D * (D + 1) / 2; }
// Index of the second-moment feature (i >= j) inside a cluster's feature row.
inline int feat2(int D, int i, int j) { return 1 + D + i * (i + 1) / 2 + j; }

// In-place inverse by LU factorisation WITHOUT pivoting (the semantics of
// invert_cpu, invert_matrix.cpp:25-101, and of the device `invert`,
// gaussian_kernel.cu:107-169).  Returns sum log|u_ii| in *logabsdet (natural
// log).  T = float or double.
template <class T> void lu_inverse_nopivot(T* a, int n, T* logabsdet, T* work /* n*n */);

// Host finalisation of one M-step from reduced statistics (see gmm.h).
void finalize_from_stats(const double* stats, const double* shift, int K, int D, clusters_t* c,
                         int num_threads, bool with_constants = true);

// constants_kernel semantics on host arrays: Rinv, constant (ln det), pi.
void constants_from_R(int K, int D, clusters_t* c, int num_threads);
// The two parts of constants_from_R, for callers that run their own loop over the clusters:
// inverse + constant of one cluster, and the mixing weights pi (needs every N[k]).
void constants_cluster(int k, int D, clusters_t* c);
// Same results for a symmetric positive definite R from one (reverse) Cholesky factorisation; also returns the
// upper-triangular W with Rinv = W^T W.  false = not positive definite, nothing written (use constants_cluster).
bool constants_cluster_spd(int k, int D, clusters_t* c, double* W);
// N, mean and covariance of ONE cluster from the packed statistics (the loop body of finalize_from_stats).
void finalize_cluster(const double* stats, const double* shift, int k, int D, clusters_t* c);
void mixing_weights(int K, clusters_t* c);

// Seeding from global column sums (double): sum x, sum x^2 over all N events,
// and the K seed rows (already gathered).  gaussian_kernel.cu:269-328,
// gaussian.cu:108-123.
void seed_from_moments(const double* sum_x, const double* sum_x2, long long N, int D, int K,
                       const float* seed_rows /* [K][D] */, clusters_t* c);
// Row index of seed event c (gaussian.cu:110-120: (int)(c*seed), seed in float).
long long seed_event_index(int c, int K, long long N);

float rissanen(float loglik, int K, int D, long long N);
float em_epsilon(int D, long long N);

// One order-reduction step (gaussian.cu:860-907).  Returns new K.  The K(K-1)/2 trial merges run on the caller's
// worker team when `pfor` is given (pfor(n, fn) calls fn(0..n-1) in parallel), else on an OpenMP team of num_threads.
using ParallelFor = std::function<void(int, const std::function<void(int)>&)>;
int reduce_order(clusters_t* c, int K, int D, int* c1, int* c2, int num_threads, const ParallelFor* pfor = nullptr);

// Packed E-step parameters for the SIMT kernel: per cluster
//   [ mean(D) | c_ii, 2c_ij (j>i) row by row (D(D+1)/2) | constant + ln(pi) ] padded to stride.
int  epack_stride(int D);
void build_epack(int K, int D, const clusters_t* c, float* out);

}  // namespace gmm
