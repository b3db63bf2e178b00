// kernels_tc.cu — tcgen05 (UTCHMMA) kernels of the EM hot path for sm_100a.
//
// M-step (mstep_N + mstep_means + mstep_covariance1 of the reference,
// gaussian_kernel.cu:522-677) as ONE tensor-core contraction over the events:
//
//     S^T[f][k] = sum_n  phi_f(z_n) * g[k][n]        z = (x - shift) * inv_scale
//
// with the per-event feature vector phi = [1, z_d, z_i z_j (i>=j)] (F = 1+D+D(D+1)/2
// rows, shared by all clusters) as the A operand and the responsibilities as the
// B operand; FP32 accumulation in TMEM.  Both operands are split into FP16
// hi + lo parts (22 significant bits) and the three significant products
// (hi*hi, lo*hi, hi*lo) are accumulated, i.e. 3 MMA passes at FP16 rate.
//
// Dataflow per CTA (persistent over a contiguous range of events, 512 threads):
//   warp 0      TMA producer: raw event tile [32][D] and raw responsibility tile
//               [64 clusters][32 events] (2-D tensor maps, zero fill out of bounds)
//   warps 4-11  operand builders: centre/scale, form the products, split hi/lo,
//               write the UMMA operand images (SWIZZLE_NONE core-matrix layout)
//   warp 1      MMA issuer: 3 x 2 x MT tcgen05.mma (M=128, N=64, K=16) per 32 events
//   warps 12-15 flush: TMEM accumulators -> FP32 partial sums in an L2-resident
//               per-CTA scratch every 512 events (the TMEM accumulation truncates:
//               measured bias -1e-7 per MMA step, see profiles/tc_probe_r1.txt)
// A second tiny kernel reduces the per-CTA partials in double and un-scales.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "host_math.h"
#include "kernels_tc.cuh"
#include "tc_ptx.cuh"

namespace gmm {

using namespace ptx;

#define TC_CUDA_TRY(expr)                                                                     \
    do {                                                                                      \
        cudaError_t e_ = (expr);                                                              \
        if (e_ != cudaSuccess)                                                                \
            return fail(GMM_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));    \
    } while (0)

// ---------------------------------------------------------------------------
// M-step kernel configuration
// ---------------------------------------------------------------------------
constexpr int kTE = 32;          // events per sub-tile (MMA K extent per operand part)
constexpr int kNCL = 64;         // clusters per CTA pass (MMA N)
constexpr int kNST = 3;          // operand stages
constexpr int kNRAW = 3;         // raw (TMA) stages
constexpr int kChunkSub = 16;    // sub-tiles between TMEM flushes (512 events)
constexpr int kMThreads = 512;
constexpr float kGammaScale = 1024.0f;   // responsibilities are scaled by 2^10 before the FP16 split

template <int D> struct MCfg {
    static constexpr int F = 1 + D + D * (D + 1) / 2;
    static constexpr int NCHUNK = (F + 7) / 8;            // 16-byte feature chunks actually written
    static constexpr int MT = (F + 127) / 128;            // M tiles of 128 feature rows
    static constexpr int PHI_PART = MT * 128 * kTE * 2;   // bytes of one precision part (hi or lo)
    static constexpr int PHI_STAGE = 2 * PHI_PART;
    static constexpr int G_PART = kNCL * kTE * 2;
    static constexpr int G_STAGE = 2 * G_PART;
    static constexpr int RAWX = (kTE * D * 4 + 127) & ~127;
    static constexpr int RAWG = kNCL * kTE * 4;
    static constexpr int OFF_PHI = 0;
    static constexpr int OFF_G = OFF_PHI + kNST * PHI_STAGE;
    static constexpr int OFF_RAWX = OFF_G + kNST * G_STAGE;
    static constexpr int OFF_RAWG = OFF_RAWX + kNRAW * RAWX;
    static constexpr int OFF_BAR = OFF_RAWG + kNRAW * RAWG;
    static constexpr int SMEM_BYTES = OFF_BAR + 256;
    static constexpr int TMEM_COLS = 2 * MT * kNCL;       // two accumulator buffers
};

__host__ __device__ constexpr int tri_row(int t) {        // t = i(i+1)/2 + j, j <= i  ->  i
    int i = 0;
    while ((i + 1) * (i + 2) / 2 <= t) i++;
    return i;
}

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// value of feature f for centred/scaled event z (f is a compile-time constant after unrolling)
template <int D>
__device__ __forceinline__ float feature_value(const float (&z)[D], int f) {
    constexpr int F = MCfg<D>::F;
    if (f == 0) return 1.0f;
    if (f <= D) return z[f - 1];
    if (f < F) {
        const int t = f - 1 - D;
        const int i = tri_row(t);
        const int j = t - i * (i + 1) / 2;
        return z[i] * z[j];
    }
    return 0.0f;
}

// Builds the 16-byte chunks c = P, P+8, P+16, ... of the feature vector of one event and
// stores hi/lo parts into the MN-major operand image:
//   byte(f, e) = (f/8)*512 + (e/8)*128 + (e%8)*16 + (f%8)*2        (LBO = 128, SBO = 512)
template <int D, int P>
__device__ __forceinline__ void build_phi_chunks(const float (&z)[D], uint8_t* hi_base, uint8_t* lo_base, int e) {
    constexpr int NCHUNK = MCfg<D>::NCHUNK;
    const int eoff = (e >> 3) * 128 + (e & 7) * 16;
#pragma unroll
    for (int c = P; c < NCHUNK; c += 8) {
        float hi[8], lo[8];
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const float v = feature_value<D>(z, c * 8 + u);
            hi[u] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);     // top 11 significant bits: exact in FP16
            lo[u] = v - hi[u];                                             // exact remainder
        }
        uint4 h, l;
        h.x = pack_half2(hi[0], hi[1]); h.y = pack_half2(hi[2], hi[3]); h.z = pack_half2(hi[4], hi[5]); h.w = pack_half2(hi[6], hi[7]);
        l.x = pack_half2(lo[0], lo[1]); l.y = pack_half2(lo[2], lo[3]); l.z = pack_half2(lo[4], lo[5]); l.w = pack_half2(lo[6], lo[7]);
        *reinterpret_cast<uint4*>(hi_base + c * 512 + eoff) = h;
        *reinterpret_cast<uint4*>(lo_base + c * 512 + eoff) = l;
    }
}

template <int D>
__global__ void __launch_bounds__(kMThreads, 1)
mstep_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g, int n,
                const float* __restrict__ shift_f, const float* __restrict__ inv_scale_f, float* __restrict__ scratch,
                int events_per_cta) {
    using C = MCfg<D>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* raw_full = bars;                 // [kNRAW]
    uint64_t* raw_empty = bars + kNRAW;        // [kNRAW]
    uint64_t* op_full = bars + 2 * kNRAW;      // [kNST]
    uint64_t* op_empty = op_full + kNST;       // [kNST]
    uint64_t* acc_full = op_empty + kNST;      // [2]
    uint64_t* acc_empty = acc_full + 2;        // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int e_begin = blockIdx.x * events_per_cta;
    const int e_end = min(n, e_begin + events_per_cta);
    const int nsub = (e_end - e_begin + kTE - 1) / kTE;
    const int k0 = blockIdx.y * kNCL;

    // ---- one-time setup ----
    for (int i = threadIdx.x * 16; i < C::OFF_RAWX; i += kMThreads * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        for (int s = 0; s < kNRAW; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 8); }
        for (int s = 0; s < kNST; s++) { mbar_init(&op_full[s], 8); mbar_init(&op_empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            for (int i = 0; i < nsub; i++) {
                const int st = i % kNRAW, ph = (i / kNRAW) & 1;
                mbar_wait(&raw_empty[st], ph ^ 1);
                mbar_arrive_expect_tx(&raw_full[st], kTE * D * 4 + C::RAWG);
                const int e0 = e_begin + i * kTE;
                tma_load_2d(smem + C::OFF_RAWX + st * C::RAWX, &tm_x, 0, e0, &raw_full[st]);
                tma_load_2d(smem + C::OFF_RAWG + st * C::RAWG, &tm_g, e0, k0, &raw_full[st]);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(128, kNCL, /*A MN-major*/ true, /*B MN-major*/ false);
            for (int i = 0; i < nsub; i++) {
                const int os = i % kNST, oph = (i / kNST) & 1;
                const int chunk = i / kChunkSub, ab = chunk & 1;
                const bool first = (i % kChunkSub) == 0;
                if (first) mbar_wait(&acc_empty[ab], ((chunk >> 1) & 1) ^ 1);
                mbar_wait(&op_full[os], oph);
                tc_fence_after();
                const uint32_t phi = smem_u32(smem + C::OFF_PHI + os * C::PHI_STAGE);
                const uint32_t gam = smem_u32(smem + C::OFF_G + os * C::G_STAGE);
                const uint32_t dcol = tmem + ab * (C::MT * kNCL);
#pragma unroll
                for (int seg = 0; seg < 3; seg++) {          // (phi_hi, g_hi), (phi_lo, g_hi), (phi_hi, g_lo)
                    const uint32_t pa = phi + (seg == 1 ? C::PHI_PART : 0);
                    const uint32_t pb = gam + (seg == 2 ? C::G_PART : 0);
#pragma unroll
                    for (int ks = 0; ks < kTE / 16; ks++) {
                        const uint64_t bdesc = make_smem_desc(pb + ks * 2048, /*LBO*/ 1024, /*SBO*/ 128);
#pragma unroll
                        for (int mt = 0; mt < C::MT; mt++) {
                            const uint64_t adesc = make_smem_desc(pa + mt * 8192 + ks * 256, /*LBO*/ 128, /*SBO*/ 512);
                            mma_f16_ss(dcol + mt * kNCL, adesc, bdesc, idesc, !(first && seg == 0 && ks == 0));
                        }
                    }
                }
                mma_commit(&op_empty[os]);                    // operand stage reusable once these MMAs retire
                if ((i % kChunkSub) == kChunkSub - 1 || i == nsub - 1) mma_commit(&acc_full[ab]);
            }
        }
    } else if (warp >= 4 && warp < 12) {
        // ===================== operand builders =====================
        const int part = warp - 4;                 // feature chunks c = part (mod 8)
        const int bt = threadIdx.x - 128;          // 0..255
        float sh[D], isc[D];
#pragma unroll
        for (int d = 0; d < D; d++) { sh[d] = shift_f[d]; isc[d] = inv_scale_f[d]; }
        for (int i = 0; i < nsub; i++) {
            const int rs = i % kNRAW, rph = (i / kNRAW) & 1;
            const int os = i % kNST, oph = (i / kNST) & 1;
            mbar_wait(&raw_full[rs], rph);
            mbar_wait(&op_empty[os], oph ^ 1);
            // --- features of event `lane` ---
            float z[D];
            {
                const float4* xr = reinterpret_cast<const float4*>(smem + C::OFF_RAWX + rs * C::RAWX + lane * (D * 4));
#pragma unroll
                for (int v = 0; v < D / 4; v++) {
                    const float4 t = xr[v];
                    z[4 * v + 0] = (t.x - sh[4 * v + 0]) * isc[4 * v + 0];
                    z[4 * v + 1] = (t.y - sh[4 * v + 1]) * isc[4 * v + 1];
                    z[4 * v + 2] = (t.z - sh[4 * v + 2]) * isc[4 * v + 2];
                    z[4 * v + 3] = (t.w - sh[4 * v + 3]) * isc[4 * v + 3];
                }
            }
            uint8_t* phi_hi = smem + C::OFF_PHI + os * C::PHI_STAGE;
            uint8_t* phi_lo = phi_hi + C::PHI_PART;
            switch (part) {
                case 0: build_phi_chunks<D, 0>(z, phi_hi, phi_lo, lane); break;
                case 1: build_phi_chunks<D, 1>(z, phi_hi, phi_lo, lane); break;
                case 2: build_phi_chunks<D, 2>(z, phi_hi, phi_lo, lane); break;
                case 3: build_phi_chunks<D, 3>(z, phi_hi, phi_lo, lane); break;
                case 4: build_phi_chunks<D, 4>(z, phi_hi, phi_lo, lane); break;
                case 5: build_phi_chunks<D, 5>(z, phi_hi, phi_lo, lane); break;
                case 6: build_phi_chunks<D, 6>(z, phi_hi, phi_lo, lane); break;
                default: build_phi_chunks<D, 7>(z, phi_hi, phi_lo, lane); break;
            }
            // --- responsibilities: thread -> (cluster row k, 8-event chunk ce) ---
            {
                const int k = bt >> 2, ce = bt & 3;
                const float4* gr = reinterpret_cast<const float4*>(smem + C::OFF_RAWG + rs * C::RAWG + k * (kTE * 4) + ce * 32);
                const float4 a = gr[0], b = gr[1];
                float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                float hi[8], lo[8];
#pragma unroll
                for (int u = 0; u < 8; u++) {
                    const float v = g[u] * kGammaScale;
                    hi[u] = __uint_as_float(__float_as_uint(v) & 0xFFFFE000u);
                    lo[u] = v - hi[u];
                }
                uint4 h, l;
                h.x = pack_half2(hi[0], hi[1]); h.y = pack_half2(hi[2], hi[3]); h.z = pack_half2(hi[4], hi[5]); h.w = pack_half2(hi[6], hi[7]);
                l.x = pack_half2(lo[0], lo[1]); l.y = pack_half2(lo[2], lo[3]); l.z = pack_half2(lo[4], lo[5]); l.w = pack_half2(lo[6], lo[7]);
                // K-major B image: byte(k, e) = (e/8)*1024 + k*16 + (e%8)*2      (LBO = 1024, SBO = 128)
                uint8_t* g_hi = smem + C::OFF_G + os * C::G_STAGE;
                *reinterpret_cast<uint4*>(g_hi + ce * 1024 + k * 16) = h;
                *reinterpret_cast<uint4*>(g_hi + C::G_PART + ce * 1024 + k * 16) = l;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) { mbar_arrive(&op_full[os]); mbar_arrive(&raw_empty[rs]); }
        }
    } else if (warp >= 12) {
        // ===================== flush: TMEM -> FP32 partials in the per-CTA scratch =====================
        const int q = warp - 12;                                   // TMEM lane quadrant (= warp % 4)
        const int nchunks = (nsub + kChunkSub - 1) / kChunkSub;
        float* my = scratch + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * C::MT * 128 + q * 32 + lane) * kNCL;
        for (int c = 0; c < nchunks; c++) {
            const int ab = c & 1;
            mbar_wait(&acc_full[ab], (c >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < C::MT; mt++) {
                float4* dst = reinterpret_cast<float4*>(my + (size_t)mt * 128 * kNCL);
#pragma unroll
                for (int h = 0; h < kNCL / 32; h++) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + ab * (C::MT * kNCL) + mt * kNCL + h * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int v = 0; v < 8; v++) {
                        float4 o = (c == 0) ? make_float4(0.f, 0.f, 0.f, 0.f) : dst[h * 8 + v];
                        o.x += __uint_as_float(r[4 * v + 0]); o.y += __uint_as_float(r[4 * v + 1]);
                        o.z += __uint_as_float(r[4 * v + 2]); o.w += __uint_as_float(r[4 * v + 3]);
                        dst[h * 8 + v] = o;
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
        }
        if (nchunks == 0) {
#pragma unroll
            for (int mt = 0; mt < C::MT; mt++)
                for (int v = 0; v < kNCL / 4; v++) reinterpret_cast<float4*>(my + (size_t)mt * 128 * kNCL)[v] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// Reduce the per-CTA FP32 partials in double, undo the operand scaling and write the packed statistics.
__global__ void mstep_tc_finalize_kernel(const float* __restrict__ scratch, int ncta_x, int MT, int K, int D, int F,
                                         const double* __restrict__ scale, double* __restrict__ stats) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * F) return;
    const int k = idx / F, f = idx % F;
    const int ty = k / kNCL, col = k % kNCL, mt = f / 128, row = f % 128;
    double s = 0;
    for (int cx = 0; cx < ncta_x; cx++)
        s += (double)scratch[(((size_t)(ty * ncta_x + cx) * MT + mt) * 128 + row) * kNCL + col];
    double fac = 1.0 / (double)kGammaScale;
    if (f >= 1 && f <= D) fac *= scale[f - 1];
    else if (f > D) {
        const int t = f - 1 - D;
        const int i = tri_row(t), j = t - i * (i + 1) / 2;
        fac *= scale[i] * scale[j];
    }
    stats[(size_t)k * F + f] += s * fac;
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct TcState {
    const float* d_x = nullptr;
    float* d_memb = nullptr;
    size_t memb_pitch = 0;
    int n = 0, D = 0, Kmax = 0, num_sms = 148;
    CUtensorMap tm_x{}, tm_g{};
    bool maps_ok = false;
    float* d_shift_f = nullptr;      // [32]
    float* d_inv_scale_f = nullptr;  // [32]
    double* d_scale = nullptr;       // [32] = 1 / inv_scale_f (double)
    float* d_scratch = nullptr;
    size_t scratch_floats = 0;
    bool have_shift = false;
};

static PFN_cuTensorMapEncodeTiled_v12000 encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

static int make_map_2d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1) {
    auto fn = encode_fn();
    if (!fn) return fail(GMM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GMM_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return GMM_OK;
}

bool tc_mstep_supported(int D, int K) {
    (void)K;
    return D == 4 || D == 8 || D == 12 || D == 16 || D == 20 || D == 24;
}
bool tc_estep_supported(int, int) { return false; }

int tc_create(TcState** out, const float* d_x_aos, int n, int D, int Kmax, float* d_memb, size_t memb_pitch, int num_sms,
              cudaStream_t stream) {
    (void)stream;
    TcState* t = new TcState();
    t->d_x = d_x_aos; t->d_memb = d_memb; t->memb_pitch = memb_pitch; t->n = n; t->D = D; t->Kmax = Kmax; t->num_sms = num_sms;
    *out = t;
    if (n <= 0 || !tc_mstep_supported(D, Kmax)) return GMM_OK;
    TC_CUDA_TRY(cudaMalloc(&t->d_shift_f, sizeof(float) * GMM_MAX_DIMENSIONS));
    TC_CUDA_TRY(cudaMalloc(&t->d_inv_scale_f, sizeof(float) * GMM_MAX_DIMENSIONS));
    TC_CUDA_TRY(cudaMalloc(&t->d_scale, sizeof(double) * GMM_MAX_DIMENSIONS));
    // tensor maps: events [n][D] (dim0 = D), responsibilities [Kmax][pitch] viewed as (events, clusters)
    if (int rc = make_map_2d(&t->tm_x, d_x_aos, (uint64_t)D, (uint64_t)n, (uint64_t)D * 4, (uint32_t)D, kTE)) return rc;
    if (int rc = make_map_2d(&t->tm_g, d_memb, (uint64_t)n, (uint64_t)Kmax, (uint64_t)memb_pitch * 4, kTE, kNCL)) return rc;
    t->maps_ok = true;
    const int mt = (num_features(D) + 127) / 128;
    const int ytiles = (Kmax + kNCL - 1) / kNCL;
    t->scratch_floats = (size_t)num_sms * ytiles * mt * 128 * kNCL;
    TC_CUDA_TRY(cudaMalloc(&t->d_scratch, sizeof(float) * t->scratch_floats));
    return GMM_OK;
}

void tc_destroy(TcState* t) {
    if (!t) return;
    cudaFree(t->d_shift_f); cudaFree(t->d_inv_scale_f); cudaFree(t->d_scale); cudaFree(t->d_scratch);
    delete t;
}

int tc_set_shift_scale(TcState* t, double* shift, const double* scale, cudaStream_t stream) {
    if (!t || !t->maps_ok) return GMM_OK;
    float sf[GMM_MAX_DIMENSIONS] = {0}, isf[GMM_MAX_DIMENSIONS] = {0};
    double sc[GMM_MAX_DIMENSIONS] = {0};
    for (int d = 0; d < t->D; d++) {
        sf[d] = (float)shift[d];
        shift[d] = (double)sf[d];                       // the host finalisation must use the value the kernel used
        const double s = (scale && scale[d] > 0) ? scale[d] : 1.0;
        isf[d] = (float)(1.0 / s);
        sc[d] = 1.0 / (double)isf[d];
    }
    TC_CUDA_TRY(cudaMemcpyAsync(t->d_shift_f, sf, sizeof(sf), cudaMemcpyHostToDevice, stream));
    TC_CUDA_TRY(cudaMemcpyAsync(t->d_inv_scale_f, isf, sizeof(isf), cudaMemcpyHostToDevice, stream));
    TC_CUDA_TRY(cudaMemcpyAsync(t->d_scale, sc, sizeof(sc), cudaMemcpyHostToDevice, stream));
    TC_CUDA_TRY(cudaStreamSynchronize(stream));         // the staging arrays live on this stack frame
    t->have_shift = true;
    return GMM_OK;
}

int tc_upload_params(TcState*, const clusters_t*, int, cudaStream_t) { return GMM_OK; }
int tc_launch_estep(TcState*, int, double*, cudaStream_t) { return fail(GMM_ERR_STATE, "tensor-core E-step not available"); }

template <int D>
static int launch_mstep_d(TcState* t, int K, double* d_stats, cudaStream_t stream) {
    using C = MCfg<D>;
    static_assert(C::SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(C::TMEM_COLS <= 512, "TMEM budget");
    static bool attr = false;
    if (!attr) {
        TC_CUDA_TRY(cudaFuncSetAttribute(mstep_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        attr = true;
    }
    int gx = t->num_sms;
    int per = (t->n + gx - 1) / gx;
    per = (per + kTE - 1) / kTE * kTE;
    gx = (t->n + per - 1) / per;
    const int gy = (K + kNCL - 1) / kNCL;
    if ((size_t)gx * gy * C::MT * 128 * kNCL > t->scratch_floats) return fail(GMM_ERR_STATE, "tensor M-step scratch too small");
    dim3 grid(gx, gy);
    mstep_tc_kernel<D><<<grid, kMThreads, C::SMEM_BYTES, stream>>>(t->tm_x, t->tm_g, t->n, t->d_shift_f, t->d_inv_scale_f,
                                                                   t->d_scratch, per);
    TC_CUDA_TRY(cudaGetLastError());
    const int F = C::F;
    mstep_tc_finalize_kernel<<<(K * F + 255) / 256, 256, 0, stream>>>(t->d_scratch, gx, C::MT, K, D, F, t->d_scale, d_stats);
    TC_CUDA_TRY(cudaGetLastError());
    return GMM_OK;
}

int tc_launch_mstep(TcState* t, int K, double* d_stats, cudaStream_t stream) {
    if (!t || !t->maps_ok) return fail(GMM_ERR_STATE, "tensor-core M-step not initialised for this shape");
    if (!t->have_shift) return fail(GMM_ERR_STATE, "tensor-core M-step needs gmm_seed (shift/scale) first");
    switch (t->D) {
        case 4: return launch_mstep_d<4>(t, K, d_stats, stream);
        case 8: return launch_mstep_d<8>(t, K, d_stats, stream);
        case 12: return launch_mstep_d<12>(t, K, d_stats, stream);
        case 16: return launch_mstep_d<16>(t, K, d_stats, stream);
        case 20: return launch_mstep_d<20>(t, K, d_stats, stream);
        case 24: return launch_mstep_d<24>(t, K, d_stats, stream);
        default: return fail(GMM_ERR_ARG, "tensor-core M-step: unsupported D");
    }
}

}  // namespace gmm
