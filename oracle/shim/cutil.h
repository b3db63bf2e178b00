/* Minimal stand-in for the NVIDIA GPU Computing SDK's cutil.h (CUDA <= 4.x),
 * which the reference includes (gaussian.cu:11) but does not vendor.  Written
 * for this repo; only what gaussian.cu uses.  Test infrastructure only. */
#ifndef GMM_SHIM_CUTIL_H
#define GMM_SHIM_CUTIL_H
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <vector>

#define CUDA_SAFE_CALL(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { \
    fprintf(stderr, "CUDA error %s:%d: %s\n", __FILE__, __LINE__, cudaGetErrorString(e_)); exit(1); } } while (0)
#define CUT_CHECK_ERROR(msg) do { cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    fprintf(stderr, "%s: %s\n", msg, cudaGetErrorString(e_)); exit(1); } } while (0)
#define CUT_BANK_CHECKER(array, index) array[index]

struct shim_timer { double total_ms; std::chrono::steady_clock::time_point t0; bool running; };
static std::vector<shim_timer>& shim_timers() { static std::vector<shim_timer> v; return v; }
static inline void cutCreateTimer(unsigned int* id) { shim_timers().push_back({0.0, {}, false}); *id = (unsigned)shim_timers().size() - 1; }
static inline void cutStartTimer(unsigned int id) { shim_timers()[id].t0 = std::chrono::steady_clock::now(); shim_timers()[id].running = true; }
static inline void cutStopTimer(unsigned int id) { auto& t = shim_timers()[id]; if (t.running) { t.total_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t.t0).count(); t.running = false; } }
static inline float cutGetTimerValue(unsigned int id) { return (float)shim_timers()[id].total_ms; }
static inline void cutDeleteTimer(unsigned int) {}
#endif
