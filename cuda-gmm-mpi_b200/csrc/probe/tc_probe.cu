// tc_probe.cu — development probe for the tcgen05 building blocks used by
// kernels_tc.cu.  Run on a B200:  ./tc_probe  (prints PASS/FAIL lines + timings)
//   T1  operand layouts: MN-major A / K-major A with SWIZZLE_NONE descriptors,
//       K-major B, M=128, checks D = A * B^T exactly (small integers).
//   T2  rounding behaviour of the FP32 accumulation in TMEM (RN vs truncation).
//   T3  MMA issue throughput for the shapes of the M-step (N=64) and E-step (N=192).
//   T4  tcgen05.ld throughput (TMEM -> registers).
//   T5  LBO = 0 aliasing of the two K chunks of a B operand step.
//   T6  do tcgen05.mma accumulation and tcgen05.ld of other TMEM columns overlap?  (round 1 result:
//       they overlap completely — 64.0 cycles per MMA and 429 B/clk of loads together vs 64.0 / 443 alone)
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../tc_ptx.cuh"

using namespace gmm::ptx;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(1); } } while (0)

struct ProbeArgs {
    const uint8_t* a_img; const uint8_t* b_img;   // smem images (bytes)
    int a_bytes, b_bytes;
    int a_lbo, a_sbo, b_lbo, b_sbo;               // descriptor byte offsets
    int a_kstep_bytes, b_kstep_bytes;             // start-address advance per K=16 step
    int a_mn_major;
    int N, ksteps, repeats, ndst;
    float* d_out;                                 // [128][N]
    long long* cycles;                            // [0] = mma cycles
};

__global__ void __launch_bounds__(128, 1) probe_mma_kernel(ProbeArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    uint8_t* sa = smem;
    uint8_t* sb = smem + ((p.a_bytes + 1023) & ~1023);
    for (int i = threadIdx.x * 16; i < p.a_bytes; i += blockDim.x * 16) *(uint4*)(sa + i) = *(const uint4*)(p.a_img + i);
    for (int i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16) *(uint4*)(sb + i) = *(const uint4*)(p.b_img + i);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(128, p.N, p.a_mn_major != 0, false);
        long long t0 = clock64();
        if (p.ksteps == 4 && p.repeats > 1) {
            // lean issue loop: descriptors precomputed, fully unrolled body (throughput measurement)
            uint64_t ad[4], bd[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                ad[k] = make_smem_desc(smem_u32(sa) + k * p.a_kstep_bytes, p.a_lbo, p.a_sbo);
                bd[k] = make_smem_desc(smem_u32(sb) + k * p.b_kstep_bytes, p.b_lbo, p.b_sbo);
            }
            const uint32_t d0 = tmem, d1 = tmem + (p.ndst > 1 ? p.N : 0), d2 = tmem + (p.ndst > 2 ? 2 * p.N : 0),
                           d3 = tmem + (p.ndst > 3 ? 3 * p.N : (p.ndst > 1 ? p.N : 0));
            mma_f16_ss(d0, ad[0], bd[0], idesc, false);
            mma_f16_ss(d1, ad[1], bd[1], idesc, p.ndst <= 1);
            mma_f16_ss(d2, ad[2], bd[2], idesc, p.ndst <= 2);
            mma_f16_ss(d3, ad[3], bd[3], idesc, p.ndst <= 3);
#pragma unroll 1
            for (int r = 1; r < p.repeats; r++) {
                mma_f16_ss(d0, ad[0], bd[0], idesc, true);
                mma_f16_ss(d1, ad[1], bd[1], idesc, true);
                mma_f16_ss(d2, ad[2], bd[2], idesc, true);
                mma_f16_ss(d3, ad[3], bd[3], idesc, true);
            }
        } else {
            for (int r = 0; r < p.repeats; r++)
                for (int k = 0; k < p.ksteps; k++) {
                    uint64_t ad = make_smem_desc(smem_u32(sa) + k * p.a_kstep_bytes, p.a_lbo, p.a_sbo);
                    uint64_t bd = make_smem_desc(smem_u32(sb) + k * p.b_kstep_bytes, p.b_lbo, p.b_sbo);
                    mma_f16_ss(tmem, ad, bd, idesc, (r | k) != 0);
                }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (p.cycles) p.cycles[0] = t1 - t0;
    }
    __syncthreads();
    tc_fence_after();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32 && c0 + j < p.N; j++) p.d_out[(size_t)(warp * 32 + lane) * p.N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

// T9: A operand from TENSOR MEMORY (tcgen05.mma ... [d], [a_tmem], b_desc): every thread = one row of A writes its K/2
// packed FP16 pairs with tcgen05.st, then one thread issues the k-steps.  repeats > 1: throughput (cycles per MMA) with
// the B operand in shared memory only — against 64 cycles for the smem-smem form at N = 128.
struct ProbeTsArgs { const uint32_t* a_pairs; const uint8_t* b_img; int b_bytes, b_lbo, b_sbo, b_kstep_bytes, N, ksteps, repeats; float* d_out; long long* cycles; };
__global__ void __launch_bounds__(128, 1) probe_mma_ts_kernel(ProbeTsArgs p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    for (int i = threadIdx.x * 16; i < p.b_bytes; i += blockDim.x * 16) *(uint4*)(smem + i) = *(const uint4*)(p.b_img + i);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ACOL = 256;
    {   // row = thread: 8 columns per k-step
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        for (int k = 0; k < p.ksteps; k++) {
            uint32_t v[8];
            for (int j = 0; j < 8; j++) v[j] = p.a_pairs[(size_t)(warp * 32 + lane) * (p.ksteps * 8) + k * 8 + j];
            tmem_st_32x8(tmem + lane_base + ACOL + k * 8, v);
        }
        tmem_st_wait();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = make_idesc_f16(128, p.N, false, false);
        long long t0 = clock64();
        for (int r = 0; r < p.repeats; r++)
            for (int k = 0; k < p.ksteps; k++) {
                uint64_t bd = make_smem_desc(smem_u32(smem) + k * p.b_kstep_bytes, p.b_lbo, p.b_sbo);
                mma_f16_ts(tmem, tmem + ACOL + k * 8, bd, idesc, (r | k) != 0);
            }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        if (p.cycles) p.cycles[0] = clock64() - t0;
    }
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < p.N; c0 += 32) {
        uint32_t r[32];
        tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        tmem_ld_wait();
        for (int j = 0; j < 32 && c0 + j < p.N; j++) p.d_out[(size_t)(warp * 32 + lane) * p.N + c0 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

// T4: TMEM load throughput
__global__ void __launch_bounds__(128, 1) probe_ld_kernel(int iters, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_base_s;
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int warp = threadIdx.x >> 5;
    float acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c0 = 0; c0 < 192; c0 += 32) {
            uint32_t r[32];
            tmem_ld_32x32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) acc = fmaf(__uint_as_float(r[j]), __uint_as_float(r[j]), acc);
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (acc == 123.456f) sink[0] = acc;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

// T4b: TMEM load throughput, better overlapped.  mode 0 loads only; 1 = + FFMA squares (4 chains); 2 = + fma.rn.f32x2 squares
__global__ void __launch_bounds__(256, 1) probe_ld2_kernel(int iters, int mode, long long* cycles, float* sink) {
    __shared__ uint32_t tmem_base_s;
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int warp = threadIdx.x >> 5;
    const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
    float a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    unsigned long long p0 = 0, p1 = 0;
    uint32_t x = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int c0 = 0; c0 < 192; c0 += 64) {
            uint32_t r[32], q[32];
            tmem_ld_32x32(tmem + lane_base + c0, r);
            tmem_ld_32x32(tmem + lane_base + c0 + 32, q);
            tmem_ld_wait();
            if (mode == 0) {
#pragma unroll
                for (int j = 0; j < 32; j++) x ^= r[j] ^ q[j];
            } else if (mode == 1) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    a0 = fmaf(__uint_as_float(r[j]), __uint_as_float(r[j]), a0);
                    a1 = fmaf(__uint_as_float(r[j + 1]), __uint_as_float(r[j + 1]), a1);
                    a2 = fmaf(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 2]), a2);
                    a3 = fmaf(__uint_as_float(r[j + 3]), __uint_as_float(r[j + 3]), a3);
                    a0 = fmaf(__uint_as_float(q[j]), __uint_as_float(q[j]), a0);
                    a1 = fmaf(__uint_as_float(q[j + 1]), __uint_as_float(q[j + 1]), a1);
                    a2 = fmaf(__uint_as_float(q[j + 2]), __uint_as_float(q[j + 2]), a2);
                    a3 = fmaf(__uint_as_float(q[j + 3]), __uint_as_float(q[j + 3]), a3);
                }
            } else {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    unsigned long long v0 = ((unsigned long long)r[j + 1] << 32) | r[j];
                    unsigned long long v1 = ((unsigned long long)r[j + 3] << 32) | r[j + 2];
                    unsigned long long w0 = ((unsigned long long)q[j + 1] << 32) | q[j];
                    unsigned long long w1 = ((unsigned long long)q[j + 3] << 32) | q[j + 2];
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p0) : "l"(v0));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p1) : "l"(v1));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p0) : "l"(w0));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p1) : "l"(w1));
                }
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
    if (a0 + a1 + a2 + a3 == 123.456f || x == 0x12345678u || p0 + p1 == 0x1234ull) sink[0] = a0;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}


// T6: do tcgen05.mma (accumulating into TMEM) and tcgen05.ld (reading OTHER TMEM columns) overlap?
//   what = 1: MMAs only (warp 0, one thread: `mmas` x (M=128, N=128, K=16), round-robin over 2 accumulators in columns 0..255)
//   what = 2: loads only (warps 4..11: `loads` x tcgen05.ld.32x32b.x32 pairs of columns 256..511 + packed squares)
//   what = 3: both at once.  cycles[0] = MMA span, cycles[1] = load span (max over the load warps' leaders).
__global__ void __launch_bounds__(384, 1) probe_overlap_kernel(int what, int mmas, int loads, long long* cycles, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];            // 2 x 4 KB operand images, zero
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    __shared__ long long ld_span[8];
    for (int i = threadIdx.x * 16; i < 8192; i += blockDim.x * 16) *(uint4*)(smem + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&tmem_base_s);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned long long p0 = 0, p1 = 0;
    if (warp == 0) {
        if (lane == 0 && (what & 1)) {
            const uint32_t idesc = make_idesc_f16(128, 128, false, false);
            const uint64_t ad = make_smem_desc(smem_u32(smem), 2048, 128);
            const uint64_t bd = make_smem_desc(smem_u32(smem + 4096), 2048, 128);
            long long t0 = clock64();
            mma_f16_ss(tmem, ad, bd, idesc, false);
            mma_f16_ss(tmem + 128, ad, bd, idesc, false);
#pragma unroll 1
            for (int r = 2; r < mmas; r += 2) {
                mma_f16_ss(tmem, ad, bd, idesc, true);
                mma_f16_ss(tmem + 128, ad, bd, idesc, true);
            }
            mma_commit(&bar);
            mbar_wait(&bar, 0);
            cycles[0] = clock64() - t0;
        }
    } else if (warp >= 4 && (what & 2)) {
        const uint32_t lane_base = ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t col0 = 256 + ((warp - 4) >> 2) * 128;       // warpgroup 0: columns 256..383, warpgroup 1: 384..511
        long long t0 = clock64();
#pragma unroll 1
        for (int it = 0; it < loads; it++) {
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 64) {
                uint32_t r[32], q[32];
                tmem_ld_32x32(tmem + lane_base + col0 + c0, r);
                tmem_ld_32x32(tmem + lane_base + col0 + c0 + 32, q);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    unsigned long long v0 = ((unsigned long long)r[j + 1] << 32) | r[j];
                    unsigned long long v1 = ((unsigned long long)r[j + 3] << 32) | r[j + 2];
                    unsigned long long w0 = ((unsigned long long)q[j + 1] << 32) | q[j];
                    unsigned long long w1 = ((unsigned long long)q[j + 3] << 32) | q[j + 2];
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p0) : "l"(v0));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p1) : "l"(v1));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p0) : "l"(w0));
                    asm volatile("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(p1) : "l"(w1));
                }
            }
        }
        long long t1 = clock64();
        if (lane == 0) ld_span[warp - 4] = t1 - t0;
    }
    if (p0 + p1 == 0x1234ull) sink[0] = 1.0f;
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0 && (what & 2)) {
        long long m = 0;
        for (int w = 0; w < 8; w++) m = ld_span[w] > m ? ld_span[w] : m;
        cycles[1] = m;
    }
    if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

// ---- host-side layout builders (the formulas of tc_ptx.cuh) -----------------
static void put_h(std::vector<uint8_t>& img, size_t off, float v) {
    __half h = __float2half_rn(v);
    memcpy(&img[off], &h, 2);
}
// K-major, no swizzle: [kchunk][row][16B]; LBO = rows*16, SBO = 128
static std::vector<uint8_t> build_kmajor(const std::vector<float>& m, int rows, int K, int& lbo, int& sbo) {
    lbo = rows * 16; sbo = 128;
    std::vector<uint8_t> img((size_t)rows * K * 2, 0);
    for (int r = 0; r < rows; r++)
        for (int k = 0; k < K; k++) {
            size_t off = (size_t)(r / 8) * sbo + (size_t)((k * 2) / 16) * lbo + (r % 8) * 16 + (k * 2) % 16;
            put_h(img, off, m[(size_t)r * K + k]);
        }
    return img;
}
// MN-major, no swizzle: [mn-chunk of 8][k][16B]; LBO = 128 (next 8 k's), SBO = K*16 (next 8 mn)
static std::vector<uint8_t> build_mnmajor(const std::vector<float>& m, int rows, int K, int& lbo, int& sbo) {
    lbo = 128; sbo = K * 16;
    std::vector<uint8_t> img((size_t)rows * K * 2, 0);
    for (int r = 0; r < rows; r++)
        for (int k = 0; k < K; k++) {
            size_t off = (size_t)((r * 2) / 16) * sbo + (size_t)(k / 8) * lbo + (k % 8) * 16 + (r * 2) % 16;
            put_h(img, off, m[(size_t)r * K + k]);
        }
    return img;
}

struct Result { std::vector<float> D; long long cycles; };

static Result run_mma(const std::vector<float>& A, const std::vector<float>& B, int N, int K, bool a_mn, int repeats, int ndst = 1) {
    ProbeArgs p{};
    int albo, asbo, blbo, bsbo;
    std::vector<uint8_t> ai = a_mn ? build_mnmajor(A, 128, K, albo, asbo) : build_kmajor(A, 128, K, albo, asbo);
    std::vector<uint8_t> bi = build_kmajor(B, N, K, blbo, bsbo);
    uint8_t *da, *db; float* dout; long long* dcyc;
    CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size()));
    CK(cudaMalloc(&dout, sizeof(float) * 128 * N)); CK(cudaMalloc(&dcyc, 8));
    CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
    p.a_img = da; p.b_img = db; p.a_bytes = (int)ai.size(); p.b_bytes = (int)bi.size();
    p.a_lbo = albo; p.a_sbo = asbo; p.b_lbo = blbo; p.b_sbo = bsbo;
    p.a_kstep_bytes = 2 * albo;       // one K=16 step = two 16-byte K chunks (K-major) or two 8-k groups (MN-major)
    p.b_kstep_bytes = 2 * blbo;
    p.a_mn_major = a_mn; p.N = N; p.ksteps = K / 16; p.repeats = repeats; p.ndst = ndst; p.d_out = dout; p.cycles = dcyc;
    size_t smem = ((ai.size() + 1023) & ~1023) + bi.size() + 1024;
    CK(cudaFuncSetAttribute(probe_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe_mma_kernel<<<1, 128, smem>>>(p);
    CK(cudaDeviceSynchronize());
    Result r; r.D.resize((size_t)128 * N);
    CK(cudaMemcpy(r.D.data(), dout, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&r.cycles, dcyc, 8, cudaMemcpyDeviceToHost));
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dcyc);
    return r;
}

static float h2f(float v) { return __half2float(__float2half_rn(v)); }

int main() {
    srand(1234);
    // ---- T1: layouts ----
    for (int a_mn = 0; a_mn <= 1; a_mn++)
        for (int N : {64, 192}) {
            const int K = 80 - (a_mn ? 16 : 0);          // 5 / 4 k-steps
            std::vector<float> A((size_t)128 * K), B((size_t)N * K);
            for (auto& v : A) v = (float)(rand() % 9 - 4);
            for (auto& v : B) v = (float)(rand() % 7 - 3);
            Result r = run_mma(A, B, N, K, a_mn, 1);
            double maxerr = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    double ref = 0;
                    for (int k = 0; k < K; k++) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                    maxerr = fmax(maxerr, fabs(ref - r.D[(size_t)m * N + n]));
                }
            printf("T1 layout A=%s-major N=%d K=%d : max|err| = %g  %s\n", a_mn ? "MN" : "K", N, K, maxerr, maxerr == 0 ? "PASS" : "FAIL");
        }
    // ---- T5: LBO = 0 on B: the two 8-element K chunks of one MMA step alias the same smem chunk ----
    {
        const int N = 192, K = 16;                 // one k-step: A = [a_lo8 | a_hi8], B chunk duplicated
        std::vector<float> A((size_t)128 * K), B8((size_t)N * 8), Bfull((size_t)N * K);
        for (auto& v : A) v = (float)(rand() % 9 - 4);
        for (auto& v : B8) v = (float)(rand() % 7 - 3);
        for (int n = 0; n < N; n++) for (int k = 0; k < 16; k++) Bfull[(size_t)n * 16 + k] = B8[(size_t)n * 8 + (k % 8)];
        // build images: A normal K-major; B image holds ONE chunk [N rows][16 B]; descriptor LBO = 0
        ProbeArgs p{};
        int albo, asbo, blbo, bsbo;
        std::vector<uint8_t> ai = build_kmajor(A, 128, K, albo, asbo);
        std::vector<uint8_t> bi = build_kmajor(B8, N, 8, blbo, bsbo);
        uint8_t *da, *db; float* dout; long long* dcyc;
        CK(cudaMalloc(&da, ai.size())); CK(cudaMalloc(&db, bi.size()));
        CK(cudaMalloc(&dout, sizeof(float) * 128 * N)); CK(cudaMalloc(&dcyc, 8));
        CK(cudaMemcpy(da, ai.data(), ai.size(), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
        p.a_img = da; p.b_img = db; p.a_bytes = (int)ai.size(); p.b_bytes = (int)bi.size();
        p.a_lbo = albo; p.a_sbo = asbo; p.b_lbo = 0; p.b_sbo = bsbo;
        p.a_kstep_bytes = 2 * albo; p.b_kstep_bytes = 0;
        p.a_mn_major = 0; p.N = N; p.ksteps = 1; p.repeats = 1; p.ndst = 1; p.d_out = dout; p.cycles = dcyc;
        size_t smem = ((ai.size() + 1023) & ~1023) + bi.size() + 1024;
        CK(cudaFuncSetAttribute(probe_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe_mma_kernel<<<1, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
        std::vector<float> D((size_t)128 * N);
        CK(cudaMemcpy(D.data(), dout, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost));
        double maxerr = 0;
        for (int m = 0; m < 128; m++)
            for (int n = 0; n < N; n++) {
                double ref = 0;
                for (int k = 0; k < K; k++) ref += (double)A[(size_t)m * K + k] * Bfull[(size_t)n * K + k];
                maxerr = fmax(maxerr, fabs(ref - D[(size_t)m * N + n]));
            }
        printf("T5 B descriptor with LBO = 0 (aliased K chunks): max|err| = %g  %s\n", maxerr, maxerr == 0 ? "PASS" : "FAIL");
        cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dcyc);
    }
    // ---- T2: accumulation rounding ----
    {
        const int N = 64, K = 16;
        std::vector<float> A((size_t)128 * K), B((size_t)N * K);
        for (auto& v : A) v = h2f(0.5f + (rand() % 1000) / 1000.0f);
        for (auto& v : B) v = h2f(0.5f + (rand() % 1000) / 1000.0f);
        for (int repeats : {1, 64, 1024, 16384}) {
            Result r = run_mma(A, B, N, K, false, repeats);
            double bias = 0, rms = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    double p = 0;
                    for (int k = 0; k < K; k++) p += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                    double ref = p * repeats;
                    double rel = (r.D[(size_t)m * N + n] - ref) / ref;
                    bias += rel; rms += rel * rel;
                }
            bias /= 128.0 * N; rms = sqrt(rms / (128.0 * N));
            printf("T2 accumulate %6d k-steps: mean rel err = %+.3e  rms = %.3e  (2^-24 = 5.96e-08; truncation predicts mean ~ -steps*3e-8)\n",
                   repeats, bias, rms);
        }
    }
    // ---- T7: rounding direction of the accumulation for NEGATIVE sums (toward zero or toward -inf?) ----
    {
        const int N = 64, K = 16;
        std::vector<float> A((size_t)128 * K), B((size_t)N * K);
        for (auto& v : A) v = -h2f(0.5f + (rand() % 1000) / 1000.0f);
        for (auto& v : B) v = h2f(0.5f + (rand() % 1000) / 1000.0f);
        for (int repeats : {1, 64}) {
            Result r = run_mma(A, B, N, K, false, repeats);
            double bias = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    double p = 0;
                    for (int k = 0; k < K; k++) p += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                    double ref = p * repeats;
                    bias += (r.D[(size_t)m * N + n] - ref) / fabs(ref);
                }
            printf("T7 negative sums, %d k-steps: mean (D - ref)/|ref| = %+.3e  (positive sums gave a negative value: < 0 here too = toward -inf, > 0 = toward zero)\n",
                   repeats, bias / (128.0 * N));
        }
    }
    // ---- T8: operands with few significant bits: is the accumulation exact while every partial sum fits 24 bits? ----
    for (int bits : {5, 6, 7, 8}) {
        const int N = 64, K = 16;
        const float q = 1.0f / (float)(1 << bits);
        std::vector<float> A((size_t)128 * K), B((size_t)N * K);
        for (auto& v : A) v = q * (float)(rand() % (1 << bits)) * ((rand() & 1) ? 1.f : -1.f);      // multiples of q, |v| < 1
        for (auto& v : B) v = q * (float)(rand() % (1 << bits));
        for (int repeats : {8, 128}) {
            Result r = run_mma(A, B, N, K, false, repeats);
            double maxerr = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    double p = 0;
                    for (int k = 0; k < K; k++) p += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                    maxerr = fmax(maxerr, fabs(r.D[(size_t)m * N + n] - p * repeats) / (q * q));
                }
            printf("T8 %d-bit operands, %3d k-steps (sum needs <= %d bits): max|err| = %g product quanta\n", bits, repeats,
                   2 * bits + 4 + (repeats == 8 ? 3 : 7), maxerr);
        }
    }
    // ---- T9: A operand from tensor memory ----
    for (int repeats : {1, 2048}) {
        const int N = 128, K = 64;
        std::vector<float> A((size_t)128 * K), B((size_t)N * K);
        for (auto& v : A) v = repeats == 1 ? (float)(rand() % 9 - 4) : 1.0f;
        for (auto& v : B) v = repeats == 1 ? (float)(rand() % 7 - 3) : 1.0f;
        std::vector<uint32_t> ap((size_t)128 * K / 2);
        for (int m = 0; m < 128; m++)
            for (int k = 0; k < K; k += 2) {
                __half lo = __float2half_rn(A[(size_t)m * K + k]), hi = __float2half_rn(A[(size_t)m * K + k + 1]);
                ap[(size_t)m * (K / 2) + k / 2] = (uint32_t)(*(unsigned short*)&lo) | ((uint32_t)(*(unsigned short*)&hi) << 16);
            }
        int blbo, bsbo;
        std::vector<uint8_t> bi = build_kmajor(B, N, K, blbo, bsbo);
        uint32_t* da; uint8_t* db; float* dout; long long* dcyc;
        CK(cudaMalloc(&da, ap.size() * 4)); CK(cudaMalloc(&db, bi.size())); CK(cudaMalloc(&dout, sizeof(float) * 128 * N)); CK(cudaMalloc(&dcyc, 8));
        CK(cudaMemcpy(da, ap.data(), ap.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(db, bi.data(), bi.size(), cudaMemcpyHostToDevice));
        ProbeTsArgs p{da, db, (int)bi.size(), blbo, bsbo, 2 * blbo, N, K / 16, repeats, dout, dcyc};
        size_t smem = bi.size() + 1024;
        CK(cudaFuncSetAttribute(probe_mma_ts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        probe_mma_ts_kernel<<<1, 128, smem>>>(p);
        CK(cudaDeviceSynchronize());
        std::vector<float> Dv((size_t)128 * N);
        long long cyc;
        CK(cudaMemcpy(Dv.data(), dout, sizeof(float) * 128 * N, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost));
        if (repeats == 1) {
            double maxerr = 0;
            for (int m = 0; m < 128; m++)
                for (int n = 0; n < N; n++) {
                    double ref = 0;
                    for (int k = 0; k < K; k++) ref += (double)A[(size_t)m * K + k] * B[(size_t)n * K + k];
                    maxerr = fmax(maxerr, fabs(ref - Dv[(size_t)m * N + n]));
                }
            printf("T9 A operand from TMEM (tcgen05.st rows, 8 columns per k-step), N=%d K=%d: max|err| = %g  %s\n", N, K, maxerr, maxerr == 0 ? "PASS" : "FAIL");
        } else {
            printf("T9 A from TMEM, M=128 N=%d K=16 fp16: %.1f cycles per MMA (A and B from shared memory: 64.1)\n", N, (double)cyc / (2048.0 * (K / 16)));
        }
        cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dcyc);
    }
    // ---- T3: MMA throughput (single CTA) ----
    for (int N : {64, 192, 256}) {
        const int K = 64;
        std::vector<float> A((size_t)128 * K, 1.0f), B((size_t)N * K, 1.0f);
        Result r = run_mma(A, B, N, K, false, 2048);
        double per = (double)r.cycles / (2048.0 * (K / 16));
        printf("T3 M=128 N=%d K=16 fp16: %.1f cycles per MMA (floor 128*N/256 = %d)\n", N, per, 128 * N / 256);
    }
    for (int N : {64, 128})
        for (int ndst : {1, 2, 3, 4}) {
            if (N * ndst > 512) continue;
            const int K = 64;
            std::vector<float> A((size_t)128 * K, 1.0f), B((size_t)N * K, 1.0f);
            Result r = run_mma(A, B, N, K, false, 2048, ndst);
            printf("T3b M=128 N=%d round-robin over %d accumulators: %.1f cycles per MMA\n", N, ndst, (double)r.cycles / (2048.0 * (K / 16)));
        }
    {
        long long* dcyc; float* sink;
        CK(cudaMalloc(&dcyc, 8)); CK(cudaMalloc(&sink, 4));
        for (int mode = 0; mode < 3; mode++)
            for (int warps : {4, 8}) {
                probe_ld2_kernel<<<1, warps * 32>>>(2000, mode, dcyc, sink);
                CK(cudaDeviceSynchronize());
                long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
                double bytes = 2000.0 * 128 * 192 * 4 * (warps / 4);
                printf("T4b mode %d (%s) %d warps: %.1f cycles per 128x192 tile per warpgroup, %.1f B/cycle/SM\n", mode,
                       mode == 0 ? "loads only" : mode == 1 ? "loads + 4-way FFMA squares" : "loads + f32x2 squares", warps,
                       c / 2000.0, bytes / c);
            }
    }
    // ---- T4: tcgen05.ld throughput ----
    {
        long long* dcyc; float* sink;
        CK(cudaMalloc(&dcyc, 8)); CK(cudaMalloc(&sink, 4));
        probe_ld_kernel<<<1, 128>>>(2000, dcyc, sink);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dcyc, 8, cudaMemcpyDeviceToHost));
        double bytes = 2000.0 * 128 * 192 * 4;
        printf("T4 tcgen05.ld 32x32b.x32 + 32 FFMA per load, 4 warps: %.1f cycles per 128x192 tile, %.1f B/cycle/SM\n", c / 2000.0, bytes / c);
    }
    // ---- T6: overlap of tcgen05.mma accumulation and tcgen05.ld of other columns ----
    {
        long long* dcyc; float* sink;
        CK(cudaMalloc(&dcyc, 16)); CK(cudaMalloc(&sink, 4));
        CK(cudaFuncSetAttribute(probe_overlap_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192));
        const int mmas = 8192;                    // 8192 x 64 cycles at the floor
        const int loads = 2048;                   // per warpgroup: 2048 x 64 KB
        for (int what = 1; what <= 3; what++) {
            long long h[2] = {0, 0};
            CK(cudaMemcpy(dcyc, h, 16, cudaMemcpyHostToDevice));
            probe_overlap_kernel<<<1, 384, 8192>>>(what, mmas, loads, dcyc, sink);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(h, dcyc, 16, cudaMemcpyDeviceToHost));
            printf("T6 %s: MMA span %lld cycles (%.1f per MMA), load span %lld cycles (%.1f B/cycle/SM)\n",
                   what == 1 ? "MMA only  " : what == 2 ? "loads only" : "both      ", h[0], h[0] / (double)mmas, h[1],
                   h[1] ? 2.0 * loads * 128.0 * 128 * 4 / (double)h[1] : 0.0);
        }
    }
    printf("probe done\n");
    return 0;
}
