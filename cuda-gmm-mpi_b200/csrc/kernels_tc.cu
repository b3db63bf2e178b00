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
//   warp 0      TMA producer: tile [D][32 events] of the pre-standardised SoA copy z and raw
//               responsibility tile [64 clusters][32 events] (2-D tensor maps, SWIZZLE_128B for
//               the latter, zero fill out of bounds)
//   warps 4-11  operand builders (two warpgroups on alternate tiles): form the products, split
//               hi/lo, write the UMMA operand images (SWIZZLE_NONE core-matrix layout)
//   warp 1      MMA issuer: 3 x 2 x MT tcgen05.mma (M=128, N=64, K=16) per 32 events
//   warps 12-15 flush: every 128 events the TMEM accumulators are added (FP32, round to
//               nearest) into register-resident partial sums (setmaxnreg gives this
//               warpgroup 240 registers) — the TMEM accumulation itself truncates:
//               measured bias -1e-7 per MMA step (profiles/tc_probe_r1.txt), so the
//               chains are kept to 24 steps; every 2048 events the partial sums move into
//               per-thread double slots of an L2-resident scratch (RED.ADD.F64)
// A second tiny kernel reduces the per-CTA partials in double and un-scales.
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_fp16.h>
#if defined(__F16C__) && defined(__AVX__)
#include <immintrin.h>
#endif
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
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
#ifndef GMM_CHUNKSUB
#define GMM_CHUNKSUB 4
#endif
#ifndef GMM_SPILL
#define GMM_SPILL 16
#endif
constexpr int kTE = 32;          // events per sub-tile (MMA K extent per operand part)
constexpr int kNCL = 64;         // clusters per CTA pass (MMA N)
// Pipeline depth.  GMM_MSTEP_NST / GMM_MSTEP_NRAW / GMM_MSTEP_TRIM are build-time experiment knobs (defaults = the validated
// configuration): a fourth operand stage fits when the phi images are trimmed to their real rows and two raw stages
// are given up (D = 24: 223.7 KB) — the builders were measured waiting on op_empty once per tile (profiles/).
#ifndef GMM_MSTEP_NST
#define GMM_MSTEP_NST 3
#endif
#ifndef GMM_MSTEP_NRAW
#define GMM_MSTEP_NRAW 4
#endif
#ifndef GMM_MSTEP_TRIM
#define GMM_MSTEP_TRIM 0
#endif
// Operand split of the M-step (build-time knobs, defaults = the shipped configuration):
//   GMM_MSTEP_RN = 1  hi = round-to-nearest FP16 of the value, lo = FP16(value - hi): the dropped lo*lo product is then
//                     zero-mean.  With the truncating split (0: hi = leading 11 bits) hi <= |value| always, lo*lo has one
//                     sign and the three kept products carry a systematic relative bias of ~1e-7 that is different for
//                     the count row (phi = 1: none), the first and the second moments — the raw-moment cancellation
//                     (|mu - shift|^2 / sigma^2 ~ 100) turns it into ~1e-5 on covariance entries (scripts/emu_mstep.py).
//   GMM_MSTEP_P4 = 1  also issue the fourth product (phi_lo, g_lo) — diagnostic only.
#ifndef GMM_MSTEP_RN
#define GMM_MSTEP_RN 1
#endif
#ifndef GMM_MSTEP_P4
#define GMM_MSTEP_P4 0
#endif
constexpr int kNST = GMM_MSTEP_NST;      // operand stages
constexpr int kNRAW = GMM_MSTEP_NRAW;    // raw (TMA) stages
constexpr int kChunkSub = GMM_CHUNKSUB;     // sub-tiles between TMEM flushes
constexpr int kMThreads = 512;
constexpr float kGammaScale = 1024.0f;   // responsibilities are scaled by 2^10 before the FP16 split

template <int D> struct MCfg {
    static constexpr int F = 1 + D + D * (D + 1) / 2;
    static constexpr int NCHUNK = (F + 7) / 8;            // 16-byte feature chunks actually written
    static constexpr int MT = (F + 127) / 128;            // M tiles of 128 feature rows
    // bytes of one precision part (hi or lo).  Trimmed: only the NCHUNK real 8-row groups; the last M tile's descriptor then
    // reads past the part into whatever follows (finite FP16 data of the next part / stage / the responsibility images):
    // garbage only in output rows >= 8 * NCHUNK, which nobody reads.
    static constexpr int PHI_PART = GMM_MSTEP_TRIM ? NCHUNK * 8 * kTE * 2 : MT * 128 * kTE * 2;
    static constexpr int PHI_STAGE = 2 * PHI_PART;
    static constexpr int G_PART = kNCL * kTE * 2;
    static constexpr int G_STAGE = 2 * G_PART;
    static constexpr int RAWX = D * kTE * 4;              // [D][32 events] from the SoA copy
    static constexpr int RAWG = kNCL * kTE * 4;
    static constexpr int OFF_PHI = 0;
    static constexpr int OFF_G = OFF_PHI + kNST * PHI_STAGE;
    static constexpr int OFF_RAWX = OFF_G + kNST * G_STAGE;
    static constexpr int OFF_RAWG = OFF_RAWX + kNRAW * RAWX;
    static constexpr int OFF_BAR = OFF_RAWG + kNRAW * RAWG;
    static constexpr int SMEM_BYTES = OFF_BAR + 512;
    static constexpr int TMEM_COLS = 2 * MT * kNCL;       // two accumulator buffers
    static_assert(OFF_RAWG % 1024 == 0 && RAWG % 1024 == 0, "SWIZZLE_128B TMA destinations need 1024-byte alignment");
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

// FP16 hi/lo split of a pair of values, packed as half2 bits (low half = first value).
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
#if GMM_MSTEP_RN
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 f = __half22float2(h);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = pack_half2(v0 - f.x, v1 - f.y);                                   // exact remainders, rounded once
#else
    const float h0 = __uint_as_float(__float_as_uint(v0) & 0xFFFFE000u);   // top 11 significant bits: exact in FP16
    const float h1 = __uint_as_float(__float_as_uint(v1) & 0xFFFFE000u);
    hi = pack_half2(h0, h1);
    lo = pack_half2(v0 - h0, v1 - h1);
#endif
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

// Builds the 16-byte chunks c = P, P+4, P+8, ... of the feature vector of one event and
// stores hi/lo parts into the MN-major operand image:
//   byte(f, e) = (f/8)*512 + (e/8)*128 + (e%8)*16 + (f%8)*2        (LBO = 128, SBO = 512)
template <int D, int P>
__device__ __forceinline__ void build_phi_chunks(const float (&z)[D], uint8_t* hi_base, uint8_t* lo_base, int e) {
    constexpr int NCHUNK = MCfg<D>::NCHUNK;
    const int eoff = (e >> 3) * 128 + (e & 7) * 16;
#pragma unroll
    for (int c = P; c < NCHUNK; c += 4) {
        uint4 h, l;
        split_pair(feature_value<D>(z, c * 8 + 0), feature_value<D>(z, c * 8 + 1), h.x, l.x);
        split_pair(feature_value<D>(z, c * 8 + 2), feature_value<D>(z, c * 8 + 3), h.y, l.y);
        split_pair(feature_value<D>(z, c * 8 + 4), feature_value<D>(z, c * 8 + 5), h.z, l.z);
        split_pair(feature_value<D>(z, c * 8 + 6), feature_value<D>(z, c * 8 + 7), h.w, l.w);
        *reinterpret_cast<uint4*>(hi_base + c * 512 + eoff) = h;
        *reinterpret_cast<uint4*>(lo_base + c * 512 + eoff) = l;
    }
}

// GS: responsibilities enter the MMA as an FP16 hi/lo pair (three products per element).  With GS = false
// they are rounded once to FP16 (round to nearest, after the 2^10 scaling) and only (phi_hi, g) + (phi_lo, g)
// are issued: a third less tensor and shared-memory work (measured: -10 % kernel time at N=10M, D=24, K=64).  The rounding perturbs every weight by an unbiased
// relative error <= 2^-12, consistently in N_k, the first and the second moments, so the result is the exact
// M-step of weights g(1 + d): the perturbation averages over the events of a cluster (~1e-4 / sqrt(n_eff))
// and is not amplified by the centring cancellation.  Measured per-call deviation of the means on clusters
// of ~10 events: 2e-4 without the pair against 5e-6 with it, so the pair is the default.
template <int D, bool GS>
__global__ void __launch_bounds__(kMThreads, 1)
mstep_tc_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_g, int n,
                double* __restrict__ scratch, int events_per_cta) {
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
        for (int s = 0; s < kNRAW; s++) { mbar_init(&raw_full[s], 1); mbar_init(&raw_empty[s], 4); }
        for (int s = 0; s < kNST; s++) { mbar_init(&op_full[s], 4); mbar_init(&op_empty[s], 1); }
        for (int s = 0; s < 2; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], 4); }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // register re-partition between the warpgroups (64K registers per SM):
    //   WG0 (TMA / MMA / alloc) 40, WG1-2 (builders) 112, WG3 (flush accumulators) 240
    if (warp < 4) {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 0) {
        // ===================== TMA producer =====================
        if (elect_one()) {
            for (int i = 0; i < nsub; i++) {
                const int st = i % kNRAW, ph = (i / kNRAW) & 1;
                mbar_wait_parked(&raw_empty[st], ph ^ 1, 500);
                mbar_arrive_expect_tx(&raw_full[st], kTE * D * 4 + C::RAWG);
                const int e0 = e_begin + i * kTE;
                tma_load_2d(smem + C::OFF_RAWX + st * C::RAWX, &tm_x, e0, 0, &raw_full[st]);
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
                if (first) mbar_wait_parked(&acc_empty[ab], ((chunk >> 1) & 1) ^ 1, 100);
                mbar_wait_parked(&op_full[os], oph, 100);
                tc_fence_after();
                const uint32_t phi = smem_u32(smem + C::OFF_PHI + os * C::PHI_STAGE);
                const uint32_t gam = smem_u32(smem + C::OFF_G + os * C::G_STAGE);
                const uint32_t dcol = tmem + ab * (C::MT * kNCL);
#pragma unroll
                for (int seg = 0; seg < (GS ? 3 + GMM_MSTEP_P4 : 2); seg++) {   // (phi_hi, g_hi), (phi_lo, g_hi), (phi_hi, g_lo) [, (phi_lo, g_lo)]
                    const uint32_t pa = phi + ((seg & 1) ? C::PHI_PART : 0);
                    const uint32_t pb = gam + (seg >= 2 ? C::G_PART : 0);
#pragma unroll
                    for (int ks = 0; ks < kTE / 16; ks++) {
                        const uint64_t bdesc = make_smem_desc(pb + ks * 256, /*LBO*/ 128, /*SBO*/ 512);
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
      }
    } else if (warp < 12) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 112;");
        // ===================== operand builders =====================
        // two builder warpgroups work on alternate sub-tiles (two sub-tiles in flight), the four warps
        // of a group split the 16-byte feature chunks (c = part mod 4)
        const int bwg = (warp - 4) >> 2;           // sub-tiles i = bwg (mod 2)
        const int part = (warp - 4) & 3;
        const int bt = threadIdx.x - 128 - bwg * 128;   // 0..127 inside the group
        for (int i = bwg; i < nsub; i += 2) {
            const int rs = i % kNRAW, rph = (i / kNRAW) & 1;
            const int os = i % kNST, oph = (i / kNST) & 1;
            mbar_wait_parked(&raw_full[rs], rph, 200);
            // --- features of event `lane` ---
            float z[D];                                // already centred and scaled (tc_set_shift_scale writes the z copy)
            {
                const float* xr = reinterpret_cast<const float*>(smem + C::OFF_RAWX + rs * C::RAWX) + lane;   // [d][32]: conflict-free
#pragma unroll
                for (int d = 0; d < D; d++) z[d] = xr[d * kTE];
            }
            // --- responsibilities: thread -> (cluster row k, 8-event chunk ce), two items per thread.
            // The raw tile is written by TMA with SWIZZLE_128B (16-byte chunk c of row r sits at chunk
            // c ^ (r & 7)), so 8 lanes reading the same chunk of 8 consecutive rows hit 8 different
            // bank groups; the operand image puts the 4 K-chunks of an 8-row group next to each other
            // (LBO = 128, SBO = 512), so a warp stores 512 contiguous bytes: no bank conflicts either way.
            uint4 gh[2], gl[2];
            float gdep = 0.0f;
#pragma unroll
            for (int it2 = 0; it2 < 2; it2++) {
                const int item = bt + it2 * 128;
                const int kg = item >> 5, l = item & 31;
                const int k = kg * 8 + (l & 7), ce = l >> 3;
                const uint8_t* grow = smem + C::OFF_RAWG + rs * C::RAWG + k * (kTE * 4);
                const float4 a = *reinterpret_cast<const float4*>(grow + (((2 * ce) ^ (k & 7)) << 4));
                const float4 b = *reinterpret_cast<const float4*>(grow + (((2 * ce + 1) ^ (k & 7)) << 4));
                float g[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
                gdep += a.x + b.x;
                if (GS) {
#pragma unroll
                    for (int u = 0; u < 8; u++) g[u] *= kGammaScale;
                    split_pair(g[0], g[1], gh[it2].x, gl[it2].x);
                    split_pair(g[2], g[3], gh[it2].y, gl[it2].y);
                    split_pair(g[4], g[5], gh[it2].z, gl[it2].z);
                    split_pair(g[6], g[7], gh[it2].w, gl[it2].w);
                } else {
#pragma unroll
                    for (int u = 0; u < 8; u++) g[u] *= kGammaScale;
                    gh[it2] = make_uint4(pack_half2(g[0], g[1]), pack_half2(g[2], g[3]), pack_half2(g[4], g[5]), pack_half2(g[6], g[7]));
                }
            }
            // The raw tiles must BE in registers before the stage goes back to the TMA producer: an mbarrier arrive does
            // not wait for the warp's outstanding LDS (measured: with nothing consuming the z loads before the arrive, the
            // refill of the stage overtook the loads of the last dimensions).  The arrive is therefore made data-dependent
            // on every load of this thread: one FADD chain over z and one component of each responsibility vector; the
            // compared bit pattern (a signalling NaN) is never the result of an addition, whatever the data.
            float dep = gdep;
#pragma unroll
            for (int d = 0; d < D; d++) dep += z[d];
            const bool never = __float_as_uint(dep) == 0xff800001u;
            __syncwarp();
            if (lane == 0 || never) mbar_arrive(&raw_empty[rs]);
            mbar_wait_parked(&op_empty[os], oph ^ 1, 200);
            uint8_t* phi_hi = smem + C::OFF_PHI + os * C::PHI_STAGE;
            uint8_t* phi_lo = phi_hi + C::PHI_PART;
            switch (part) {
                case 0: build_phi_chunks<D, 0>(z, phi_hi, phi_lo, lane); break;
                case 1: build_phi_chunks<D, 1>(z, phi_hi, phi_lo, lane); break;
                case 2: build_phi_chunks<D, 2>(z, phi_hi, phi_lo, lane); break;
                default: build_phi_chunks<D, 3>(z, phi_hi, phi_lo, lane); break;
            }
            {
                // K-major B image: byte(k, e) = (k/8)*512 + (e/8)*128 + (k%8)*16 + (e%8)*2      (LBO = 128, SBO = 512)
                uint8_t* g_hi = smem + C::OFF_G + os * C::G_STAGE;
#pragma unroll
                for (int it2 = 0; it2 < 2; it2++) {
                    const int item = bt + it2 * 128;
                    const int kg = item >> 5, l = item & 31;
                    *reinterpret_cast<uint4*>(g_hi + kg * 512 + l * 16) = gh[it2];
                    if (GS) *reinterpret_cast<uint4*>(g_hi + C::G_PART + kg * 512 + l * 16) = gl[it2];
                }
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&op_full[os]);
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 240;");
        // ===================== flush: TMEM -> register-resident FP32 partial sums =====================
        const int q = warp - 12;                                   // TMEM lane quadrant (= warp % 4)
        const int nchunks = (nsub + kChunkSub - 1) / kChunkSub;
        float racc[C::MT * kNCL];
#pragma unroll
        for (int j = 0; j < C::MT * kNCL; j++) racc[j] = 0.0f;
        // second level: every kSpill chunks the FP32 partial sums move into this thread's double
        // partials in the (L2-resident) per-CTA scratch, bounding the FP32 random walk to kSpill adds
        constexpr int kSpill = GMM_SPILL;
        double* my = scratch + ((size_t)(blockIdx.y * gridDim.x + blockIdx.x) * C::MT * 128 + q * 32 + lane) * kNCL;
        for (int c = 0; c < nchunks; c++) {
            const int ab = c & 1;
            mbar_wait_parked(&acc_full[ab], (c >> 1) & 1, 500);
            tc_fence_after();
#pragma unroll
            for (int mt = 0; mt < C::MT; mt++) {
#pragma unroll
                for (int h = 0; h < kNCL / 32; h++) {
                    uint32_t r[32];
                    tmem_ld_32x32(tmem + ((uint32_t)(q * 32) << 16) + ab * (C::MT * kNCL) + mt * kNCL + h * 32, r);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; j++) racc[(mt * 2 + h) * 32 + j] += __uint_as_float(r[j]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty[ab]);
            if ((c % kSpill) == kSpill - 1 || c == nchunks - 1) {
                // fire-and-forget double reductions (RED.ADD.F64) into this thread's private slots of the
                // zero-initialised scratch: no load latency on the flush path
#pragma unroll
                for (int mt = 0; mt < C::MT; mt++)
#pragma unroll
                    for (int j = 0; j < kNCL; j++) {
                        atomicAdd(my + (size_t)mt * 128 * kNCL + j, (double)racc[mt * kNCL + j]);
                        racc[mt * kNCL + j] = 0.0f;
                    }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// z = (x - shift) * inv_scale over the SoA event copy, once per data set: the centred/scaled copy the M-step
// tiles are cut from (the E-step converters apply the same two operations to the AoS rows, so both kernels
// see bit-identical z).
__global__ void standardise_soa_kernel(const float* __restrict__ xs, float* __restrict__ zs, size_t pitch, int n, int D,
                                       const float* __restrict__ shift_f, const float* __restrict__ inv_scale_f) {
    const int d = blockIdx.y;
    const float s = shift_f[d], isc = inv_scale_f[d];
    const float* x = xs + (size_t)d * pitch;
    float* z = zs + (size_t)d * pitch;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        z[e] = __fmul_rn(__fsub_rn(x[e], s), isc);
}

// Reduce the per-CTA FP32 partials in double, undo the operand scaling and write the packed statistics.
__global__ void __launch_bounds__(256)
mstep_tc_finalize_kernel(const double* __restrict__ scratch, int ncta_x, int MT, int K, int D, int F,
                         const double* __restrict__ scale, double* __restrict__ stats) {
    // one block per feature row f; thread -> (cluster column, quarter of the CTAs): 512-byte coalesced reads
    __shared__ double part[4][kNCL];
    const int f = blockIdx.x, mt = f / 128, row = f % 128;
    const int col = threadIdx.x & (kNCL - 1), q = threadIdx.x / kNCL;
    double fac = 1.0 / (double)kGammaScale;
    if (f >= 1 && f <= D) fac *= scale[f - 1];
    else if (f > D) {
        const int t = f - 1 - D;
        const int i = tri_row(t), j = t - i * (i + 1) / 2;
        fac *= scale[i] * scale[j];
    }
    for (int ty = 0; ty * kNCL < K; ty++) {
        double s = 0;
        for (int cx = q; cx < ncta_x; cx += 4)
            s += scratch[(((size_t)(ty * ncta_x + cx) * MT + mt) * 128 + row) * kNCL + col];
        part[q][col] = s;
        __syncthreads();
        const int k = ty * kNCL + col;
        if (q == 0 && k < K) stats[(size_t)k * F + f] += (part[0][col] + part[1][col] + part[2][col] + part[3][col]) * fac;
        __syncthreads();
    }
}


// ===========================================================================
// E-step (estep1 + estep2 of the reference, gaussian_kernel.cu:383-512) on
// tensor cores.  With Rinv = W^T W (W upper triangular, from the Cholesky
// factor of Rinv computed on the host) the quadratic form is
//     q_k(x) = || W_k (x - mu_k) ||^2 = || W'_k z + v_k ||^2 ,   z = (x - shift) * inv_scale,
// i.e. ONE GEMM  Y[n][(k,d)] = Z~[n][:] . B[(k,d)][:]  with the constant folded in
// through a ones column, followed by a square-and-sum epilogue, the log-sum-exp
// over the clusters and the log-likelihood reduction.  Operands are FP16 hi/lo
// split (z = zh + zl, W' = Wh + Wl); the K dimension concatenates
//     [ zh_c zl_c ]_c | [ zh | 1 1 0.. ]     x     [ Wh_c Wh_c ]_c | [ Wl | vh vl 0.. ]
// (the lo*lo product is dropped), FP32 accumulation in TMEM.  The duplicated Wh
// chunk is not stored twice: the B descriptor of those MMA steps uses a leading
// byte offset of 0, so both 8-element K chunks alias the same shared-memory chunk
// (verified by csrc/probe/tc_probe.cu, test T5).  That makes the whole B operand
// (all clusters: 172 KB at K=64, D=24) RESIDENT in the shared memory of one CTA;
// only the event tiles stream.
//
// One persistent CTA per SM, 512 threads:
//   warp 1      MMA issuer: per 128-event tile, per supergroup of 16 clusters and per block c of 8
//               output dimensions, the k-steps that block needs (tcgen05.mma M=128, N=128, K=16)
//   warp 2      TMEM allocation (4 accumulator buffers x 128 columns)
//   warps 4-7   converters: coalesced loads of the event rows, centre/scale, FP16 hi/lo
//               split, K-major SWIZZLE_NONE operand image (2 stages)
//   warps 8-15  epilogue (two warpgroups, each takes 8 of the 16 clusters of every supergroup):
//               tcgen05.ld -> packed squares, carried over the blocks -> base-2 logits ->
//               max / sum-exp2 (+ exchange between the warpgroups) -> responsibilities
//               (coalesced 128-byte row segments) + log-likelihood (double)
// ===========================================================================
constexpr int kEThreads = 512;

// GMM_ESTEP_PROF=1 (build-time, diagnostic variant only): the epilogue of CTA 0 / warp 8 accumulates clock64() spans of its
// phases and prints them at the end of the kernel — where do the cycles of a tile go (waiting for accumulators, tcgen05.ld,
// squares, logits + log-sum-exp + exchange, stores)?  Off by default: no code is generated.
#ifndef GMM_ESTEP_PROF
#define GMM_ESTEP_PROF 0
#endif
#if GMM_ESTEP_PROF
#define EPROF(stmt) stmt
#else
#define EPROF(stmt)
#endif

// Block structure.  W is upper triangular, so the 8 output columns d in [8c, 8c+8) of a cluster
// ("block" c) only need the K chunks z_j with j >= c.  Columns are therefore grouped by block:
// one MMA N tile = block c of 16 clusters (N = 128), and block c issues only the k-steps it needs —
// 5 + 4 + 2 = 11 instead of 15 at D = 24 (-27 % tensor work and TMEM accumulator traffic).
template <int D, int NWG = 2> struct ECfg {
    static_assert(D % 8 == 0, "tensor E-step: D must be a multiple of 8");
    static constexpr int CP = D / 8;                          // 8-wide chunks of z / blocks of output columns
    static constexpr int NLO = (CP + 1 + 1) / 2 * 2;          // chunks of the [zh | ones (| pad)] x [Wl | v] part
    static constexpr int NCHKA = 2 * CP + NLO;                // A image chunks: (zh_c, zl_c) pairs, then zh.., ones, pad
    static constexpr int NCHKB = CP + NLO;                    // B image chunks: Wh_c, then Wl.., v, pad
    static constexpr int GB = 16;                             // clusters per supergroup
    static constexpr int N = GB * 8;                          // MMA N = one block of a supergroup (128 columns)
    static constexpr int MAXSG = 64 / GB;                     // up to 64 clusters resident
    static constexpr int NBUF = 512 / N;                      // TMEM accumulator buffers (4)
    static constexpr int CW = GB / NWG;                       // clusters per epilogue warpgroup per supergroup
    static constexpr int LPT = MAXSG * CW;                    // logits held per epilogue thread (32, or 16 with 4 warpgroups)
    static constexpr int A_STAGE = NCHKA * 128 * 16;
    static constexpr int B_BLOCK = NCHKB * N * 16;            // one block of one supergroup
    static constexpr int B_SG = CP * B_BLOCK;
    static constexpr int OFF_B = 0;
    static constexpr int OFF_A = OFF_B + MAXSG * B_SG;
    static constexpr int OFF_CK = OFF_A + 2 * A_STAGE;        // float2[64]: (constant + ln(pi)) * log2(e), -0.5 * log2(e) / scale_k^2
    static constexpr int OFF_EX = OFF_CK + 512;               // exchange: [2 parity][NWG][128] x (max, sum)
    static constexpr int OFF_BAR = OFF_EX + 2 * NWG * 128 * 8;
    static constexpr int SMEM_BYTES = OFF_BAR + 512;
    static constexpr int THREADS = 256 + 128 * NWG;           // warpgroup 0, converters, NWG epilogue warpgroups
    static_assert(NWG == 2 || NWG == 4, "epilogue warpgroups");
};

// ALT = false: both epilogue warpgroups work on every tile (each takes 8 of the 16 clusters of a supergroup and they
// exchange (max, sum) through shared memory).  ALT = true (experimental, GMM_ESTEP_ALT=1): the warpgroups take alternate
// tiles and each handles all 16 clusters — no exchange, and the log-sum-exp / store phase of one tile overlaps the
// accumulator reads of the next, so the MMA issuer is not held up by full TMEM buffers during that phase.
// NWG = 4 (experimental, GMM_ESTEP_WG4=1): four epilogue warpgroups, each with 4 of the 16 clusters of a supergroup (one
// tcgen05.ld.x32 per block and warp, 16 logits per thread, 96 registers) — four instead of two epilogue warps per sub-core
// to hide the load / mbarrier / MUFU latencies the two-warp version exposes (tc_probe T6: the hardware floor is the MMA time).
template <int D, bool ALT, int NWG = 2>
__global__ void __launch_bounds__(256 + 128 * NWG, 1)
estep_tc_kernel(const float* __restrict__ x_aos, const uint8_t* __restrict__ b_img, const float* __restrict__ ck,
                const float* __restrict__ shift_f, const float* __restrict__ inv_scale_f, float* __restrict__ memb,
                size_t pitch, int n, int K, int NSG, double* __restrict__ ll_out, float* __restrict__ den_out) {
    // K / NSG / b_img / ck / memb describe ONE pass of at most 64 clusters.  With more than 64 clusters the host
    // launches one pass per 64 (den_out != nullptr): each pass normalises within itself and records its
    // log-denominator per event; estep_tc_combine_kernel then rescales the passes against each other.
    using C = ECfg<D, NWG>;
    static_assert(!ALT || NWG == 2, "alternate-tile epilogue is written for two warpgroups");
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
    uint64_t* a_full = bars;            // [2]  4 converter warps
    uint64_t* a_empty = bars + 2;       // [2]  tcgen05.commit
    uint64_t* b_full = bars + 4;        // [1]
    uint64_t* acc_full = bars + 5;      // [NBUF]  tcgen05.commit
    uint64_t* acc_empty = bars + 5 + C::NBUF;     // [NBUF]  8 epilogue warps
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5 + 2 * C::NBUF);
    // ALT: the odd tiles' accumulators are announced on a second set of barriers, so that each epilogue warpgroup only
    // ever waits for the NEXT phase of a barrier (a parity wait cannot tell phase p from phase p + 2)
    uint64_t* acc_full_odd = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR + 384);   // [NBUF]
    float2* ck_s = reinterpret_cast<float2*>(smem + C::OFF_CK);
    float2* ex = reinterpret_cast<float2*>(smem + C::OFF_EX);
    float* sh_s = reinterpret_cast<float*>(smem + C::OFF_BAR + 128);   // [32] shift, [32] inverse scale
    float* isc_s = sh_s + 32;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = (n + 127) / 128;
    const int my_tiles = (int)blockIdx.x < ntiles ? (ntiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(&a_full[s], 4); mbar_init(&a_empty[s], 1); }
        for (int s = 0; s < C::NBUF; s++) { mbar_init(&acc_full[s], 1); mbar_init(&acc_empty[s], ALT ? 4 : 4 * NWG); }
        mbar_init(b_full, 1);
        if (ALT) for (int s = 0; s < C::NBUF; s++) mbar_init(&acc_full_odd[s], 1);
        fence_mbar_init();
    }
    // base-2 logits in the epilogue: l2 = ck * log2(e) + (-0.5 * log2(e) / scale_k^2) * |scale_k * y|^2   (the per-cluster
    // power-of-two scale_k keeps the FP16 whitening factors in range whatever the cluster's width, see bimg_cluster)
    if (threadIdx.x < 64) ck_s[threadIdx.x] = make_float2(ck[threadIdx.x] * 1.4426950408889634f, ck[64 + threadIdx.x]);
    if (threadIdx.x < D) { sh_s[threadIdx.x] = shift_f[threadIdx.x]; isc_s[threadIdx.x] = inv_scale_f[threadIdx.x]; }
    __syncthreads();
    if (threadIdx.x == 0) {                    // resident B operand: one TMA bulk copy per block
        mbar_arrive_expect_tx(b_full, (uint32_t)NSG * C::B_SG);
        for (int g = 0; g < NSG * C::CP; g++) tma_load_1d(smem + C::OFF_B + g * C::B_BLOCK, b_img + (size_t)g * C::B_BLOCK, C::B_BLOCK, b_full);
    }
    if (warp == 2) tmem_alloc<512>(tmem_slot);
    mbar_wait(b_full, 0);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // register re-partition inside the CTA's launch allocation (512 x 128): WG0 40, converters 72, epilogue 2 x 200
    if (warp < 4) {
      // register pools: NWG = 2 launches 512 x 128, NWG = 4 launches 768 x 80 (= 61440): 128 x (24 + 64) + 512 x 96 = 60416
      if constexpr (NWG == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 24;");
      else if constexpr (ALT) asm volatile("setmaxnreg.dec.sync.aligned.u32 32;");
      else asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
      if (warp == 1) {
        // ===================== MMA issuer =====================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_f16(128, C::N, false, false);
            uint32_t pe = (1u << C::NBUF) - 1u;                // wait parity of acc_empty[b], one bit per buffer
            for (int it = 0; it < my_tiles; it++) {
                const int as = it & 1, aph = (it >> 1) & 1;
                mbar_wait_parked(&a_full[as], aph, 200);
                tc_fence_after();
                const uint32_t abase = smem_u32(smem + C::OFF_A + as * C::A_STAGE);
                uint32_t buf = 0;                              // the buffer sequence restarts with every tile (epilogue: same rule)
                for (int sg = 0; sg < NSG; sg++) {
#pragma unroll
                    for (int c = 0; c < C::CP; c++) {
                        mbar_wait_parked(&acc_empty[buf], (pe >> buf) & 1u, 100);
                        pe ^= 1u << buf;
                        tc_fence_after();
                        const uint32_t bbase = smem_u32(smem + C::OFF_B + (sg * C::CP + c) * C::B_BLOCK);
                        bool acc = false;
#pragma unroll
                        for (int j = c; j < C::CP; j++) {       // (zh_j, zl_j) x (Wh_j, Wh_j): B chunk aliased through LBO = 0
                            const uint64_t adesc = make_smem_desc(abase + (2 * j) * 2048, /*LBO*/ 2048, /*SBO*/ 128);
                            const uint64_t bdesc = make_smem_desc(bbase + j * (C::N * 16), /*LBO*/ 0, /*SBO*/ 128);
                            mma_f16_ss(tmem + buf * C::N, adesc, bdesc, idesc, acc);
                            acc = true;
                        }
#pragma unroll
                        for (int t = 0; t < C::NLO / 2; t++) {  // (zh.., ones) x (Wl.., v): needed iff it holds a chunk index >= c
                            if (2 * t + 1 >= c) {
                                const uint64_t adesc = make_smem_desc(abase + (2 * C::CP + 2 * t) * 2048, /*LBO*/ 2048, /*SBO*/ 128);
                                const uint64_t bdesc = make_smem_desc(bbase + (C::CP + 2 * t) * (C::N * 16), /*LBO*/ C::N * 16, /*SBO*/ 128);
                                mma_f16_ss(tmem + buf * C::N, adesc, bdesc, idesc, acc);
                                acc = true;
                            }
                        }
                        mma_commit(ALT && (it & 1) ? &acc_full_odd[buf] : &acc_full[buf]);
                        buf = (buf + 1) % C::NBUF;
                    }
                }
                mma_commit(&a_empty[as]);
            }
        }
      }
    } else if (warp < 8) {
        if constexpr (NWG == 4) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        else if constexpr (ALT) asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 72;");
        // ===================== converters =====================
        const int row = threadIdx.x - 128;
        for (int it = 0; it < my_tiles; it++) {
            const int st = it & 1, ph = (it >> 1) & 1;
            const long long e = (long long)((int)blockIdx.x + it * (int)gridDim.x) * 128 + row;
            float4 xv[D / 4];
            if (e < n) {
                const float4* xr = reinterpret_cast<const float4*>(x_aos + (size_t)e * D);
#pragma unroll
                for (int v = 0; v < D / 4; v++) xv[v] = __ldg(xr + v);
            } else {
#pragma unroll
                for (int v = 0; v < D / 4; v++) xv[v] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            uint32_t hi[D / 2], lo[D / 2];
#pragma unroll
            for (int v = 0; v < D / 4; v++) {
                const float4 t = xv[v];
                const float4 s4 = reinterpret_cast<const float4*>(sh_s)[v], i4 = reinterpret_cast<const float4*>(isc_s)[v];
                const float z0 = (t.x - s4.x) * i4.x, z1 = (t.y - s4.y) * i4.y;
                const float z2 = (t.z - s4.z) * i4.z, z3 = (t.w - s4.w) * i4.w;
                const __half2 h01 = __floats2half2_rn(z0, z1), h23 = __floats2half2_rn(z2, z3);
                const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                hi[2 * v] = *reinterpret_cast<const uint32_t*>(&h01);
                hi[2 * v + 1] = *reinterpret_cast<const uint32_t*>(&h23);
                lo[2 * v] = pack_half2(z0 - f01.x, z1 - f01.y);
                lo[2 * v + 1] = pack_half2(z2 - f23.x, z3 - f23.y);
            }
            mbar_wait_parked(&a_empty[st], ph ^ 1, 1000);
            uint8_t* a = smem + C::OFF_A + st * C::A_STAGE + row * 16;     // K-major: [chunk][row][16 B]
#pragma unroll
            for (int c = 0; c < C::CP; c++) {
                const uint4 h = make_uint4(hi[4 * c], hi[4 * c + 1], hi[4 * c + 2], hi[4 * c + 3]);
                const uint4 l = make_uint4(lo[4 * c], lo[4 * c + 1], lo[4 * c + 2], lo[4 * c + 3]);
                *reinterpret_cast<uint4*>(a + (2 * c) * 2048) = h;
                *reinterpret_cast<uint4*>(a + (2 * c + 1) * 2048) = l;
                *reinterpret_cast<uint4*>(a + (2 * C::CP + c) * 2048) = h;
            }
            *reinterpret_cast<uint4*>(a + (3 * C::CP) * 2048) = make_uint4(0x3C003C00u, 0u, 0u, 0u);      // {1, 1, 0...}
            if (C::NLO > C::CP + 1) *reinterpret_cast<uint4*>(a + (3 * C::CP + 1) * 2048) = make_uint4(0u, 0u, 0u, 0u);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&a_full[st]);
        }
    } else {
        if constexpr (NWG == 4) asm volatile("setmaxnreg.inc.sync.aligned.u32 96;");
        else if constexpr (ALT) asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");   // 128 x (32 + 64 + 2 x 208) = 64K registers
        else asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        // ===================== epilogue =====================
      if constexpr (ALT) {
        // warpgroup g takes the tiles it = g, g+2, ...; thread = event row, all clusters of the pass in this thread
        const int g = (warp - 8) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        const int NB = NSG * C::CP;                            // blocks per tile; buffer of block j: j % NBUF (restarts per tile)
        constexpr float kLn2 = 0.6931471805599453f;
        double ll_acc = 0.0;
        for (int it = g; it < my_tiles; it += 2) {
            const long long e = (long long)((int)blockIdx.x + it * (int)gridDim.x) * 128 + row;
            uint32_t pf[C::NBUF];                              // parity of acc_full[b] at its first use in this tile
#pragma unroll
            for (int b = 0; b < C::NBUF; b++) {
                const int uses = b < NB ? (NB - b + C::NBUF - 1) / C::NBUF : 0;   // uses of buffer b per tile
                pf[b] = (uint32_t)((it >> 1) * uses) & 1u;     // this warpgroup's own barrier set: its (it / 2)-th tile
            }
            float lg[C::MAXSG * C::GB];                        // base-2 logits of all clusters of the pass
            float mx = -INFINITY;
#pragma unroll
            for (int sg = 0; sg < C::MAXSG; sg++) {
                if (sg < NSG) {
                    uint64_t qs[C::GB];                        // packed partial sums of squares, one pair per cluster
#pragma unroll
                    for (int i = 0; i < C::GB; i++) qs[i] = 0ull;
#pragma unroll
                    for (int c = 0; c < C::CP; c++) {
                        const int buf = (sg * C::CP + c) % C::NBUF;
                        mbar_wait_parked(g ? &acc_full_odd[buf] : &acc_full[buf], pf[buf], 200);
                        pf[buf] ^= 1u;
                        tc_fence_after();
#pragma unroll
                        for (int half = 0; half < 2; half++) {
                            const uint32_t tcol = tmem + lane_base + buf * C::N + half * 64;
                            uint32_t v[64];                    // 8 clusters x 8 columns
                            tmem_ld_32x32(tcol, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                            tmem_ld_32x32(tcol + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
                            tmem_ld_wait();
                            if (half == 1) {                   // the whole block is in registers: hand the buffer back
                                tc_fence_before();
                                __syncwarp();
                                if (lane == 0) mbar_arrive(&acc_empty[buf]);
                            }
#pragma unroll
                            for (int i = 0; i < 8; i++) {
                                sq_acc2(qs[half * 8 + i], v[i * 8 + 0], v[i * 8 + 1]);
                                sq_acc2(qs[half * 8 + i], v[i * 8 + 2], v[i * 8 + 3]);
                                sq_acc2(qs[half * 8 + i], v[i * 8 + 4], v[i * 8 + 5]);
                                sq_acc2(qs[half * 8 + i], v[i * 8 + 6], v[i * 8 + 7]);
                            }
                        }
                    }
#pragma unroll
                    for (int i = 0; i < C::GB; i++) {
                        float lo, hi;
                        asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(qs[i]));
                        const float2 cm = ck_s[sg * C::GB + i];
                        const float l = fmaf(cm.y, lo + hi, cm.x);
                        lg[sg * C::GB + i] = l;
                        mx = fmaxf(mx, l);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < C::GB; i++) lg[sg * C::GB + i] = -INFINITY;
                }
            }
            float sm = 0.f;                                    // estep2, gaussian_kernel.cu:481-503
#pragma unroll
            for (int j = 0; j < C::MAXSG * C::GB; j++) { lg[j] = ex2_approx(lg[j] - mx); sm += lg[j]; }
            const float denom = fmaf(mx, kLn2, logf(sm));
            const float scale = 1.0f / sm;
            if (e < n) {
                if (den_out) den_out[e] = denom;
                else ll_acc += (double)denom;
                float* gp = memb + e;
#pragma unroll
                for (int kg = 0; kg < C::MAXSG * C::GB / 8; kg++) {          // whole 8-cluster groups (rows padded to 8)
                    if (kg * 8 < K) {
                        float* gq = gp + (size_t)(kg * 8) * pitch;
#pragma unroll
                        for (int i = 0; i < 8; i++) {
                            *gq = lg[kg * 8 + i] * scale;                    // :498-501
                            gq += pitch;
                        }
                    }
                }
            }
        }
        if (den_out == nullptr) {
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 16);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 8);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 4);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 2);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 1);
            if (lane == 0) atomicAdd(ll_out, ll_acc);
        }
      } else {
        const int wg = (warp - 8) >> 2, q = warp & 3;
        const int row = q * 32 + lane;
        const uint32_t lane_base = (uint32_t)(q * 32) << 16;
        double ll_acc = 0.0;
        EPROF(long long pr_wait = 0; long long pr_ld = 0; long long pr_sq = 0; long long pr_lse = 0; long long pr_st = 0; long long pr_t0 = clock64();)
        uint32_t pf[C::NBUF];                                  // wait parity of acc_full[b]; b is static after unrolling
#pragma unroll
        for (int b = 0; b < C::NBUF; b++) pf[b] = 0u;
        constexpr float kLn2 = 0.6931471805599453f;
        for (int it = 0; it < my_tiles; it++) {
            const long long e = (long long)((int)blockIdx.x + it * (int)gridDim.x) * 128 + row;
            float lg[C::LPT];                                  // logits in base 2 (ck_s is pre-multiplied by log2 e)
            float mx = -INFINITY;
#pragma unroll
            for (int sg = 0; sg < C::MAXSG; sg++) {
                if (sg < NSG) {
                    uint64_t qa[C::CW], qb[C::CW];             // packed partial sums of squares of this warpgroup's 8 clusters
#pragma unroll
                    for (int i = 0; i < C::CW; i++) { qa[i] = 0ull; qb[i] = 0ull; }
#pragma unroll
                    for (int c = 0; c < C::CP; c++) {
                        const int buf = (sg * C::CP + c) % C::NBUF;   // compile-time: the sequence restarts with every tile
                        EPROF(const long long p0 = clock64();)
                        mbar_wait_parked(&acc_full[buf], pf[buf], 200);
                        pf[buf] ^= 1u;
                        tc_fence_after();
                        EPROF(const long long p1 = clock64(); pr_wait += p1 - p0;)
                        const uint32_t tcol = tmem + lane_base + buf * C::N + wg * (C::CW * 8);
                        uint32_t v[C::CW * 8];                 // CW clusters x 8 columns
                        tmem_ld_32x32(tcol, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
                        if constexpr (C::CW * 8 > 32) tmem_ld_32x32(tcol + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[C::CW * 8 - 32]));
                        tmem_ld_wait();
                        EPROF(const long long p2 = clock64(); pr_ld += p2 - p1;)
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_empty[buf]);   // the block is in registers: hand the buffer back
#pragma unroll
                        for (int i = 0; i < C::CW; i++) {
                            if constexpr (NWG == 2) {
                                sq_acc2(qa[i], v[i * 8 + 0], v[i * 8 + 1]);
                                sq_acc2(qb[i], v[i * 8 + 2], v[i * 8 + 3]);
                                sq_acc2(qa[i], v[i * 8 + 4], v[i * 8 + 5]);
                                sq_acc2(qb[i], v[i * 8 + 6], v[i * 8 + 7]);
                            } else {                           // 96-register budget: one accumulator pair per cluster
                                sq_acc2(qa[i], v[i * 8 + 0], v[i * 8 + 1]);
                                sq_acc2(qa[i], v[i * 8 + 2], v[i * 8 + 3]);
                                sq_acc2(qa[i], v[i * 8 + 4], v[i * 8 + 5]);
                                sq_acc2(qa[i], v[i * 8 + 6], v[i * 8 + 7]);
                                asm volatile("" : "+l"(qa[i]));   // keep this block's squares ahead of the next block's load (register budget)
                            }
                        }
                        EPROF(asm volatile("" : "+l"(qa[0])); pr_sq += clock64() - p2;)
                    }
#pragma unroll
                    for (int i = 0; i < C::CW; i++) {
                        const float2 cm = ck_s[sg * C::GB + wg * C::CW + i];
                        const float l = fmaf(cm.y, hsum2(qa[i], qb[i]), cm.x);
                        lg[sg * C::CW + i] = l;
                        mx = fmaxf(mx, l);
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < C::CW; i++) lg[sg * C::CW + i] = -INFINITY;
                }
            }
            // log-sum-exp over the clusters (estep2, gaussian_kernel.cu:481-503): local part, then the two warpgroups combine
            EPROF(const long long p3 = clock64();)
            float sm = 0.f;
#pragma unroll
            for (int j = 0; j < C::LPT; j++) { lg[j] = ex2_approx(lg[j] - mx); sm += lg[j]; }
            float2* exb = ex + (it & 1) * (NWG * 128);
            exb[wg * 128 + row] = make_float2(mx, sm);
            named_bar_sync(1, NWG * 128);
            float M, own, S;
            if constexpr (NWG == 2) {
                const float2 o = exb[(wg ^ 1) * 128 + row];
                M = fmaxf(mx, o.x);
                own = ex2_approx(mx - M);
                S = sm * own + o.y * ex2_approx(o.x - M);
            } else {
                float2 o[NWG];
                M = mx;
#pragma unroll
                for (int w = 0; w < NWG; w++) { o[w] = exb[w * 128 + row]; M = fmaxf(M, o[w].x); }
                own = ex2_approx(mx - M);
                S = 0.f;
#pragma unroll
                for (int w = 0; w < NWG; w++) S += o[w].y * ex2_approx(o[w].x - M);
            }
            const float denom = fmaf(M, kLn2, logf(S));              // :490-494, back in natural units
            const float scale = own / S;                             // exp(l - denom) = 2^(l2 - mx) * 2^(mx - M) / S
            EPROF(const long long p4 = clock64(); pr_lse += p4 - p3;)
            if (e < n) {
                if (wg == 0) {
                    if (den_out) den_out[e] = denom;
                    else ll_acc += (double)denom;
                }
                // Rows [K, 8*ceil(K/8)) are written too (zeros of the padding clusters): the buffer is allocated in
                // multiples of 8 rows, which keeps the 8 stores of a group unpredicated.
                float* gp = memb + (size_t)(wg * C::CW) * pitch + e;      // row of this warpgroup's first cluster
#pragma unroll
                for (int sg = 0; sg < C::MAXSG; sg++) {
                    if (sg * C::GB + wg * C::CW < K) {
                        float* gq = gp + (size_t)(sg * C::GB) * pitch;
#pragma unroll
                        for (int i = 0; i < C::CW; i++) {
                            *gq = lg[sg * C::CW + i] * scale;             // :498-501
                            gq += pitch;
                        }
                    }
                }
            }
            EPROF(pr_st += clock64() - p4;)
        }
        EPROF(if (blockIdx.x == 0 && warp == 8 && lane == 0)
                  printf("estep epilogue profile (CTA 0, warp 8): %d tiles, cycles per tile: total %.0f = wait acc_full %.0f + tcgen05.ld %.0f + squares %.0f + "
                         "lse/exchange %.0f + stores %.0f + rest\n", my_tiles, (double)(clock64() - pr_t0) / my_tiles, (double)pr_wait / my_tiles,
                         (double)pr_ld / my_tiles, (double)pr_sq / my_tiles, (double)pr_lse / my_tiles, (double)pr_st / my_tiles);)
        if (wg == 0 && den_out == nullptr) {
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 16);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 8);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 4);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 2);
            ll_acc = ll_acc + __shfl_down_sync(0xffffffffu, ll_acc, 1);
            if (lane == 0) atomicAdd(ll_out, ll_acc);
        }
          }
}
    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc<512>(tmem);
}

// More than 64 clusters: joins the per-pass normalisations.  den[p][e] = ln sum_{k in pass p} exp(logit);
// the event's denominator is the log-sum-exp over the passes (estep2, gaussian_kernel.cu:481-503) and every
// responsibility of pass p is multiplied by exp(den_p - denom).
constexpr int kEMaxPass = GMM_MAX_CLUSTERS / 64;
__global__ void __launch_bounds__(256)
estep_tc_combine_kernel(float* __restrict__ memb, size_t pitch, int n, int K, int NP, const float* __restrict__ den,
                        double* __restrict__ ll_out) {
    __shared__ double red[8];
    double ll = 0.0;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float dp[kEMaxPass];
        float M = -INFINITY;
#pragma unroll
        for (int p = 0; p < kEMaxPass; p++) {
            dp[p] = p < NP ? den[(size_t)p * pitch + e] : -INFINITY;
            M = fmaxf(M, dp[p]);
        }
        float S = 0.f;
#pragma unroll
        for (int p = 0; p < kEMaxPass; p++) S += __expf(dp[p] - M);
        const float denom = M + logf(S);
        ll += (double)denom;
#pragma unroll
        for (int p = 0; p < kEMaxPass; p++) {
            if (p < NP) {
                const float f = __expf(dp[p] - denom);
                const int kend = min(K, (p + 1) * 64);
                float* g = memb + (size_t)(p * 64) * pitch + e;
                for (int k = p * 64; k < kend; k++, g += pitch) *g *= f;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) ll += __shfl_down_sync(0xffffffffu, ll, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ll;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; w++) t += red[w];
        atomicAdd(ll_out, t);
    }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
struct TcState {
    const float* d_x = nullptr;
    const float* d_x_soa = nullptr;
    float* d_z_soa = nullptr;        // [D][memb_pitch] centred/scaled SoA copy (M-step TMA source)
    float* d_memb = nullptr;
    size_t memb_pitch = 0;
    int n = 0, D = 0, Kmax = 0, num_sms = 148;
    CUtensorMap tm_x{}, tm_g{};
    bool maps_ok = false;
    float* d_shift_f = nullptr;      // [32]
    float* d_inv_scale_f = nullptr;  // [32]
    double* d_scale = nullptr;       // [32] = 1 / inv_scale_f (double)
    double* d_scratch = nullptr;
    size_t scratch_floats = 0;
    size_t scratch_clean_bytes = 0;  // leading bytes of d_scratch known to be zero
    size_t scratch_dirty_bytes = 0;  // bytes the last M-step launch wrote
    bool have_shift = false;
    // E-step
    CUtensorMap tm_x128{};
    bool emap_ok = false;
    uint8_t* d_bimg = nullptr;       // [MAXNG * B_GROUP] resident B operand image
    uint8_t* h_bimg = nullptr;       // pinned
    size_t bimg_bytes = 0;
    uint8_t* d_opnd = nullptr;       // [ck (e_ck_len floats) | B image]; d_ck / d_bimg point into it
    uint8_t* h_opnd = nullptr;       // pinned mirror
    cudaEvent_t ev_h2d = nullptr;    // the last operand copy has left the pinned buffer
    bool h2d_pending = false;
    float* d_ck = nullptr;           // [passes][ck 64 | mult 64]: additive constant and quadratic-form multiplier per cluster
    float* h_ck = nullptr;           // pinned mirror
    float* d_den = nullptr;          // [passes][memb_pitch] per-pass log-denominators (Kmax > 64 only)
    int e_ck_len = 0;                // Kmax rounded up to whole passes of 64
    int e_NG = 0;
    int host_threads = 8;
    bool estep_alt = false;          // experimental alternating-warpgroup E-step epilogue (GMM_ESTEP_ALT=1)
    bool estep_wg4 = false;          // experimental four-warpgroup E-step epilogue (GMM_ESTEP_WG4=1)
    int gamma_split = 2;             // M-step: FP16 hi/lo pair for the responsibilities: 0 never, 1 always, 2 by cluster size
    // cudaFuncAttributeMaxDynamicSharedMemorySize is per device: the "already set" flags live with the (per-device) state
    bool attr_estep = false, attr_estep4 = false, attr_mstep = false;
    double h_shift[GMM_MAX_DIMENSIONS] = {0}, h_scale[GMM_MAX_DIMENSIONS] = {0};
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

static int make_map_2d(CUtensorMap* m, const void* base, uint64_t dim0, uint64_t dim1, uint64_t stride1_bytes, uint32_t box0, uint32_t box1,
                       bool swizzle128 = false) {
    auto fn = encode_fn();
    if (!fn) return fail(GMM_ERR_CUDA, "cuTensorMapEncodeTiled not available from the driver");
    cuuint64_t dims[2] = {dim0, dim1};
    cuuint64_t strides[1] = {stride1_bytes};
    cuuint32_t box[2] = {box0, box1};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(GMM_ERR_CUDA, "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return GMM_OK;
}

bool tc_mstep_supported(int D, int K) {
    (void)K;
    return D == 4 || D == 8 || D == 12 || D == 16 || D == 20 || D == 24;
}
bool tc_estep_supported(int D, int K) { return (D == 8 || D == 16 || D == 24) && K >= 1 && K <= GMM_MAX_CLUSTERS; }

// bytes of the B image of one pass (64 clusters = MAXSG supergroups)
template <int D> static size_t ecfg_pass_bytes() { return (size_t)ECfg<D>::MAXSG * ECfg<D>::B_SG; }
static size_t pass_bytes_for(int D) {
    switch (D) { case 8: return ecfg_pass_bytes<8>(); case 16: return ecfg_pass_bytes<16>(); case 24: return ecfg_pass_bytes<24>(); default: return 0; }
}

int tc_create(TcState** out, const float* d_x_aos, const float* d_x_soa, int n, int D, int Kmax, float* d_memb, size_t memb_pitch, int num_sms,
              cudaStream_t stream) {
    (void)stream;
    TcState* t = new TcState();
    if (const char* alt = getenv("GMM_ESTEP_ALT")) t->estep_alt = atoi(alt) != 0;
    if (const char* wg4 = getenv("GMM_ESTEP_WG4")) t->estep_wg4 = atoi(wg4) != 0;
    t->d_x = d_x_aos; t->d_x_soa = d_x_soa; t->d_memb = d_memb; t->memb_pitch = memb_pitch; t->n = n; t->D = D; t->Kmax = Kmax; t->num_sms = num_sms;
    *out = t;
    if (n <= 0 || !tc_mstep_supported(D, Kmax)) return GMM_OK;
    TC_CUDA_TRY(cudaMalloc(&t->d_shift_f, sizeof(float) * GMM_MAX_DIMENSIONS));
    TC_CUDA_TRY(cudaMalloc(&t->d_inv_scale_f, sizeof(float) * GMM_MAX_DIMENSIONS));
    TC_CUDA_TRY(cudaMalloc(&t->d_scale, sizeof(double) * GMM_MAX_DIMENSIONS));
    // tensor maps: SoA events [D][pitch] viewed as (events, dims) -> smem tile [D][32 events];
    // responsibilities [Kmax][pitch] viewed as (events, clusters)
    TC_CUDA_TRY(cudaMalloc(&t->d_z_soa, sizeof(float) * memb_pitch * D));
    if (int rc = make_map_2d(&t->tm_x, t->d_z_soa, (uint64_t)n, (uint64_t)D, (uint64_t)memb_pitch * 4, kTE, (uint32_t)D)) return rc;
    if (int rc = make_map_2d(&t->tm_g, d_memb, (uint64_t)n, (uint64_t)Kmax, (uint64_t)memb_pitch * 4, kTE, kNCL, /*swizzle128=*/true)) return rc;
    t->maps_ok = true;
    if (D == 8 || D == 16 || D == 24) {
        const int passes = (Kmax + 63) / 64;
        t->e_ck_len = passes * 64;
        t->bimg_bytes = (size_t)passes * pass_bytes_for(D);
        // one staging / device buffer [ck | B image]: the operand of an iteration travels in ONE H2D copy
        const size_t ck_bytes = sizeof(float) * 2 * t->e_ck_len;     // 512 B per pass: keeps the image 16-byte aligned
        TC_CUDA_TRY(cudaMalloc(&t->d_opnd, ck_bytes + t->bimg_bytes));
        TC_CUDA_TRY(cudaMallocHost(&t->h_opnd, ck_bytes + t->bimg_bytes));
        t->d_ck = reinterpret_cast<float*>(t->d_opnd);
        t->h_ck = reinterpret_cast<float*>(t->h_opnd);
        t->d_bimg = t->d_opnd + ck_bytes;
        t->h_bimg = t->h_opnd + ck_bytes;
        TC_CUDA_TRY(cudaEventCreateWithFlags(&t->ev_h2d, cudaEventDisableTiming));
        if (passes > 1) TC_CUDA_TRY(cudaMalloc(&t->d_den, sizeof(float) * (size_t)passes * memb_pitch));
        t->emap_ok = true;
    }
    const int mt = (num_features(D) + 127) / 128;
    const int ytiles = (Kmax + kNCL - 1) / kNCL;
    t->scratch_floats = (size_t)num_sms * ytiles * mt * 128 * kNCL;
    TC_CUDA_TRY(cudaMalloc(&t->d_scratch, sizeof(double) * t->scratch_floats));
    return GMM_OK;
}

void tc_set_host_threads(TcState* t, int n) { if (t) t->host_threads = n < 1 ? 1 : n; }
void tc_set_gamma_split(TcState* t, int mode) { if (t) t->gamma_split = mode < 0 ? 0 : (mode > 2 ? 2 : mode); }

void tc_destroy(TcState* t) {
    if (!t) return;
    cudaFree(t->d_shift_f); cudaFree(t->d_inv_scale_f); cudaFree(t->d_scale); cudaFree(t->d_scratch);
    cudaFree(t->d_opnd); cudaFree(t->d_den); cudaFree(t->d_z_soa);
    if (t->h_opnd) cudaFreeHost(t->h_opnd);
    if (t->ev_h2d) cudaEventDestroy(t->ev_h2d);
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
        t->h_shift[d] = shift[d];
        t->h_scale[d] = sc[d];
    }
    TC_CUDA_TRY(cudaMemcpyAsync(t->d_shift_f, sf, sizeof(sf), cudaMemcpyHostToDevice, stream));
    TC_CUDA_TRY(cudaMemcpyAsync(t->d_inv_scale_f, isf, sizeof(isf), cudaMemcpyHostToDevice, stream));
    TC_CUDA_TRY(cudaMemcpyAsync(t->d_scale, sc, sizeof(sc), cudaMemcpyHostToDevice, stream));
    {
        dim3 grid((unsigned)std::min<long long>(4LL * t->num_sms, ((long long)t->n + 255) / 256), (unsigned)t->D);
        standardise_soa_kernel<<<grid, 256, 0, stream>>>(t->d_x_soa, t->d_z_soa, t->memb_pitch, t->n, t->D, t->d_shift_f, t->d_inv_scale_f);
        TC_CUDA_TRY(cudaGetLastError());
    }
    TC_CUDA_TRY(cudaStreamSynchronize(stream));         // the staging arrays live on this stack frame
    t->have_shift = true;
    return GMM_OK;
}

// Host side of the tensor E-step operand: per cluster the upper-triangular factor W of
// Rinv = W^T W (Cholesky of the symmetrised inverse covariance, double), expressed in the
// centred/scaled coordinates of the kernel, FP16 hi/lo split, laid out as the resident
// K-major B image ([supergroup][block][chunk][128 rows][16 B]).  Fails (GMM_ERR_STATE) when Rinv
// is not positive definite or the factor overflows FP16; the caller then uses the SIMT kernel
// for this state.

// float -> IEEE half bits, round to nearest even (normal, subnormal and zero; callers check the range)
static inline uint16_t f2h_bits(float f) {
    uint32_t x;
    std::memcpy(&x, &f, 4);
    const uint32_t sign = (x >> 16) & 0x8000u;
    x &= 0x7fffffffu;
    if (x >= 0x47800000u) return (uint16_t)(sign | 0x7c00u);
    if (x < 0x38800000u) {                      // below the smallest normal half: value * 2^24, rounded
        float a;
        std::memcpy(&a, &x, 4);
        return (uint16_t)(sign | (uint32_t)lrintf(a * 16777216.0f));
    }
    x += ((x >> 13) & 1u) + 0xfffu;
    return (uint16_t)(sign | ((x - 0x38000000u) >> 13));
}
static inline float h2f_bits(uint16_t h) {
    const uint32_t sign = (uint32_t)(h & 0x8000u) << 16;
    const uint32_t em = h & 0x7fffu;
    float r;
    if (em >= 0x0400u) {                         // normal (inf/nan never produced here)
        const uint32_t x = sign | ((em << 13) + 0x38000000u);
        std::memcpy(&r, &x, 4);
    } else {
        r = (float)em * (1.0f / 16777216.0f);
        if (sign) r = -r;
    }
    return r;
}

// Operand rows of cluster k (k >= K: padding cluster of the last supergroup).  Returns 0, 1 (Rinv not positive
// definite) or 2 (factor outside the FP16 range).  Clusters are independent: callers may run this in parallel.
template <int D>
static int bimg_cluster(TcState* t, const clusters_t* host, int k, int K) {
    using C = ECfg<D>;
    int bad = 0;
    {
        const int sg = k / C::GB, i = k % C::GB;
        // 16-byte K-chunk `chunk` of output column d of this cluster: K-major SWIZZLE_NONE image
        // [supergroup][block c = d/8][chunk][N = 16 clusters x 8 columns][16 B]
        auto rowp = [&](int d, int chunk) -> uint16_t* {
            const int c = d / 8, ncol = i * 8 + (d % 8);
            return reinterpret_cast<uint16_t*>(t->h_bimg + ((size_t)sg * C::CP + c) * C::B_BLOCK + (size_t)chunk * C::N * 16 + (size_t)ncol * 16);
        };
        if (k >= K) {                                    // padding cluster of the last supergroup: all-zero rows
            for (int d = 0; d < D; d++)
                for (int c = 0; c < C::NCHKB; c++) std::memset(rowp(d, c), 0, 16);
            return 0;
        }
        double A[D][D], Gc[D][D];
        const float* Ri = host->Rinv + (size_t)k * D * D;
        for (int r = 0; r < D; r++)
            for (int j = 0; j < D; j++) { A[r][j] = 0.5 * ((double)Ri[r * D + j] + (double)Ri[j * D + r]); Gc[r][j] = 0.0; }
        bool ok = true;
        for (int j = 0; j < D; j++) {                    // right-looking Cholesky A = Gc Gc^T (axpy updates vectorise)
            const double d = A[j][j];
            if (!(d > 0.0) || !std::isfinite(d)) { ok = false; break; }
            const double piv = std::sqrt(d), rp = 1.0 / piv;
            Gc[j][j] = piv;
            for (int r = j + 1; r < D; r++) Gc[r][j] = A[r][j] * rp;
            for (int r = j + 1; r < D; r++) {
                const double l = Gc[r][j];
                for (int cc = j + 1; cc <= r; cc++) A[r][cc] -= l * Gc[cc][j];
            }
        }
        if (!ok) return 1;
        // rows of W = Gc^T in the kernel's coordinates:  y_d = sum_j W'[d][j] z_j + v_d,  W'[d][j] = Gc[j][d] * scale_j (j >= d)
        alignas(32) float wrow[D][D];
        double vd[D];
        float amax = 0.f;
        for (int d = 0; d < D; d++) {
            double v = 0.0;
            for (int j = 0; j < D; j++) {
                const double w = (j >= d) ? Gc[j][d] : 0.0;
                v -= w * ((double)host->means[(size_t)k * D + j] - t->h_shift[j]);
                wrow[d][j] = (float)(w * t->h_scale[j]);
                amax = std::fmax(amax, std::fabs(wrow[d][j]));
            }
            vd[d] = v;
            amax = std::fmax(amax, (float)std::fabs(v));
        }
        if (!std::isfinite(amax)) return 2;
        // Per-cluster power-of-two scale: the largest operand entry lands in [2^12, 2^13) whatever the width of the
        // cluster (a cluster of relative width 1e-4 has factors ~1e4 - 1e5 and used to leave the FP16 range; a very wide
        // one pushed its lo parts into the FP16 subnormals).  The epilogue divides the squared norm by scale^2 (exact).
        int e2 = amax > 0.f ? 12 - std::ilogb(amax) : 0;
        e2 = e2 > 40 ? 40 : (e2 < -40 ? -40 : e2);
        for (int d = 0; d < D; d++) {
            for (int j = 0; j < D; j++) wrow[d][j] = std::ldexp(wrow[d][j], e2);
            vd[d] = std::ldexp(vd[d], e2);
            for (int c = 0; c < C::CP; c++) {
                uint16_t *ph = rowp(d, c), *pl = rowp(d, C::CP + c);      // x (zh_c, zl_c) [aliased], x zh_c
#if defined(__F16C__) && defined(__AVX__)
                const __m256 w8 = _mm256_load_ps(&wrow[d][8 * c]);
                const __m128i h8 = _mm256_cvtps_ph(w8, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC);
                const __m256 l8 = _mm256_sub_ps(w8, _mm256_cvtph_ps(h8));  // exact: hi is w rounded to 11 bits
                _mm_storeu_si128(reinterpret_cast<__m128i*>(ph), h8);
                _mm_storeu_si128(reinterpret_cast<__m128i*>(pl), _mm256_cvtps_ph(l8, _MM_FROUND_TO_NEAREST_INT | _MM_FROUND_NO_EXC));
#else
                for (int e = 0; e < 8; e++) {
                    const uint16_t wh = f2h_bits(wrow[d][8 * c + e]);
                    ph[e] = wh;
                    pl[e] = f2h_bits(wrow[d][8 * c + e] - h2f_bits(wh));
                }
#endif
            }
            const float vf = (float)vd[d];
            if (!(std::fabs(vf) < 6.0e4f)) bad = 2;
            uint16_t* pv = rowp(d, 2 * C::CP);
            const uint16_t vh = f2h_bits(vf);
            std::memset(pv, 0, 16);
            pv[0] = vh;
            pv[1] = f2h_bits((float)(vd[d] - (double)h2f_bits(vh)));
            if (C::NCHKB > 2 * C::CP + 1) std::memset(rowp(d, 2 * C::CP + 1), 0, 16);
        }
        float* ckp = t->h_ck + (size_t)(k / 64) * 128 + (k % 64);
        ckp[0] = host->constant[k] + logf(host->pi[k]);      // additive term of estep1 (gaussian_kernel.cu:442)
        ckp[64] = (float)std::ldexp(-0.5 * 1.4426950408889634, -2 * e2);
    }
    return bad;
}

static int bimg_cluster_any(TcState* t, const clusters_t* host, int k, int K) {
    switch (t->D) {
        case 8: return bimg_cluster<8>(t, host, k, K);
        case 16: return bimg_cluster<16>(t, host, k, K);
        case 24: return bimg_cluster<24>(t, host, k, K);
        default: return 3;
    }
}

int tc_params_begin(TcState* t, int K, cudaStream_t stream) {
    if (!t || !t->emap_ok) return fail(GMM_ERR_STATE, "tensor E-step not initialised for this shape");
    if (!t->have_shift) return fail(GMM_ERR_STATE, "tensor E-step needs the global moments (shift/scale) first");
    (void)stream;
    if (t->h2d_pending) {                                // the previous copy out of the pinned buffer must have finished
        TC_CUDA_TRY(cudaEventSynchronize(t->ev_h2d));
        t->h2d_pending = false;
    }
    for (int k = K; k < t->e_ck_len; k++) {               // padding clusters: never win the log-sum-exp
        float* ckp = t->h_ck + (size_t)(k / 64) * 128 + (k % 64);
        ckp[0] = -1e30f;
        ckp[64] = 0.f;
    }
    return GMM_OK;
}
int tc_params_padded(const TcState*, int K) { return (K + 15) / 16 * 16; }
int tc_params_cluster(TcState* t, const clusters_t* host, int k, int K) { return bimg_cluster_any(t, host, k, K); }
int tc_params_commit(TcState* t, int K, int bad, cudaStream_t stream) {
    if (bad == 1) return fail(GMM_ERR_STATE, "tensor E-step: inverse covariance of a cluster is not positive definite");
    if (bad == 2) return fail(GMM_ERR_STATE, "tensor E-step: whitening factor exceeds the FP16 range");
    if (bad) return fail(GMM_ERR_ARG, "tensor E-step: unsupported D");
    t->e_NG = (K + 15) / 16;
    // only the supergroups in use travel (the image is contiguous per supergroup; 4 supergroups = 64 clusters)
    const size_t used = (size_t)t->e_NG * (pass_bytes_for(t->D) / 4);
    TC_CUDA_TRY(cudaMemcpyAsync(t->d_opnd, t->h_opnd, sizeof(float) * 2 * t->e_ck_len + used, cudaMemcpyHostToDevice, stream));
    TC_CUDA_TRY(cudaEventRecord(t->ev_h2d, stream));
    t->h2d_pending = true;
    return GMM_OK;
}


int tc_upload_params(TcState* t, const clusters_t* host, int K, cudaStream_t stream) {
    if (int rc = tc_params_begin(t, K, stream)) return rc;
    const int kp = tc_params_padded(t, K);
    const int nt = t->host_threads;
    int bad = 0;
    (void)nt;
#pragma omp parallel for schedule(static) num_threads(nt) reduction(max : bad) if (nt > 1 && K >= 8)
    for (int k = 0; k < kp; k++) {
        const int b = tc_params_cluster(t, host, k, K);
        bad = b > bad ? b : bad;
    }
    return tc_params_commit(t, K, bad, stream);
}

template <int D>
static int launch_estep_d(TcState* t, int K, double* d_ll, cudaStream_t stream) {
    using C = ECfg<D>;
    static_assert(C::SMEM_BYTES <= 232448, "shared memory budget");
    if (!t->attr_estep) {
        TC_CUDA_TRY(cudaFuncSetAttribute(estep_tc_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        TC_CUDA_TRY(cudaFuncSetAttribute(estep_tc_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        t->attr_estep = true;
    }
    auto kernel = t->estep_alt ? estep_tc_kernel<D, true> : estep_tc_kernel<D, false>;
    int threads = kEThreads, smem_bytes = C::SMEM_BYTES;
    if (t->estep_wg4) {
        using C4 = ECfg<D, 4>;
        static_assert(C4::SMEM_BYTES <= 232448, "shared memory budget");
        if (!t->attr_estep4) {
            TC_CUDA_TRY(cudaFuncSetAttribute(estep_tc_kernel<D, false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, C4::SMEM_BYTES));
            t->attr_estep4 = true;
        }
        kernel = estep_tc_kernel<D, false, 4>;
        threads = C4::THREADS;
        smem_bytes = C4::SMEM_BYTES;
    }
    const int ntiles = (t->n + 127) / 128;
    int grid = t->num_sms;
    if (grid > ntiles) grid = ntiles;
    if (grid < 1) grid = 1;
    const int NP = (K + 63) / 64;
    if (NP > 1 && !t->d_den) return fail(GMM_ERR_STATE, "tensor E-step: context was created for at most 64 clusters");
    for (int p = 0; p < NP; p++) {
        const int Kp = K - 64 * p < 64 ? K - 64 * p : 64;
        kernel<<<grid, threads, smem_bytes, stream>>>(
            t->d_x, t->d_bimg + (size_t)p * C::MAXSG * C::B_SG, t->d_ck + 128 * p, t->d_shift_f, t->d_inv_scale_f,
            t->d_memb + (size_t)(64 * p) * t->memb_pitch, t->memb_pitch, t->n, Kp, (Kp + C::GB - 1) / C::GB, d_ll,
            NP > 1 ? t->d_den + (size_t)p * t->memb_pitch : nullptr);
        TC_CUDA_TRY(cudaGetLastError());
    }
    if (NP > 1) {
        estep_tc_combine_kernel<<<t->num_sms * 8, 256, 0, stream>>>(t->d_memb, t->memb_pitch, t->n, K, NP, t->d_den, d_ll);
        TC_CUDA_TRY(cudaGetLastError());
    }
    return GMM_OK;
}

int tc_launch_estep(TcState* t, int K, double* d_ll, cudaStream_t stream) {
    if (!t || !t->emap_ok) return fail(GMM_ERR_STATE, "tensor E-step not initialised for this shape");
    switch (t->D) {
        case 8: return launch_estep_d<8>(t, K, d_ll, stream);
        case 16: return launch_estep_d<16>(t, K, d_ll, stream);
        case 24: return launch_estep_d<24>(t, K, d_ll, stream);
        default: return fail(GMM_ERR_ARG, "tensor E-step: unsupported D");
    }
}

// Clusters of at least this many (soft) events take the single-FP16 responsibilities under mode 2: the rounding is an
// unbiased relative perturbation <= 2^-12 per weight, so the statistics of a cluster move by ~1.4e-4 / sqrt(n_eff),
// n_eff >= N_k: below 3.1e-6 from here on (30x under the parity bar, the size of the FP32 noise already there).
constexpr float kGammaSplitMinN = 2048.0f;

template <int D>
static int launch_mstep_d(TcState* t, int K, double* d_stats, cudaStream_t stream, float min_nk, int* pair_out) {
    using C = MCfg<D>;
    static_assert(C::SMEM_BYTES <= 232448, "shared memory budget");
    static_assert(C::TMEM_COLS <= 512, "TMEM budget");
    if (!t->attr_mstep) {
        TC_CUDA_TRY(cudaFuncSetAttribute(mstep_tc_kernel<D, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        TC_CUDA_TRY(cudaFuncSetAttribute(mstep_tc_kernel<D, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        t->attr_mstep = true;
    }
    int gx = t->num_sms;
    int per = (t->n + gx - 1) / gx;
    per = (per + kTE - 1) / kTE * kTE;
    gx = (t->n + per - 1) / per;
    const int gy = (K + kNCL - 1) / kNCL;
    if ((size_t)gx * gy * C::MT * 128 * kNCL > t->scratch_floats) return fail(GMM_ERR_STATE, "tensor M-step scratch too small");
    dim3 grid(gx, gy);
    // the flush warps add into zeroed per-CTA slots; tc_mstep_cleanup() zeroes the scratch after the statistics have
    // left for the host (off the critical path), so only a launch that finds it dirty pays for a memset in front
    const size_t scratch_bytes = sizeof(double) * (size_t)gx * gy * C::MT * 128 * kNCL;
    if (t->scratch_clean_bytes < scratch_bytes) TC_CUDA_TRY(cudaMemsetAsync(t->d_scratch, 0, scratch_bytes, stream));
    const bool split = t->gamma_split == 1 || (t->gamma_split == 2 && !(min_nk >= kGammaSplitMinN));
    if (pair_out) *pair_out = split ? 1 : 0;
    if (split)
        mstep_tc_kernel<D, true><<<grid, kMThreads, C::SMEM_BYTES, stream>>>(t->tm_x, t->tm_g, t->n, t->d_scratch, per);
    else
        mstep_tc_kernel<D, false><<<grid, kMThreads, C::SMEM_BYTES, stream>>>(t->tm_x, t->tm_g, t->n, t->d_scratch, per);
    TC_CUDA_TRY(cudaGetLastError());
    const int F = C::F;
    mstep_tc_finalize_kernel<<<F, 256, 0, stream>>>(t->d_scratch, gx, C::MT, K, D, F, t->d_scale, d_stats);
    TC_CUDA_TRY(cudaGetLastError());
    t->scratch_dirty_bytes = scratch_bytes;
    t->scratch_clean_bytes = 0;
    return GMM_OK;
}

int tc_mstep_cleanup(TcState* t, cudaStream_t stream) {
    if (!t || !t->d_scratch || t->scratch_dirty_bytes == 0) return GMM_OK;
    TC_CUDA_TRY(cudaMemsetAsync(t->d_scratch, 0, t->scratch_dirty_bytes, stream));
    t->scratch_clean_bytes = t->scratch_dirty_bytes;
    t->scratch_dirty_bytes = 0;
    return GMM_OK;
}

int tc_launch_mstep(TcState* t, int K, double* d_stats, cudaStream_t stream, float min_nk, int* pair_out) {
    if (!t || !t->maps_ok) return fail(GMM_ERR_STATE, "tensor-core M-step not initialised for this shape");
    if (!t->have_shift) return fail(GMM_ERR_STATE, "tensor-core M-step needs gmm_seed (shift/scale) first");
    switch (t->D) {
        case 4: return launch_mstep_d<4>(t, K, d_stats, stream, min_nk, pair_out);
        case 8: return launch_mstep_d<8>(t, K, d_stats, stream, min_nk, pair_out);
        case 12: return launch_mstep_d<12>(t, K, d_stats, stream, min_nk, pair_out);
        case 16: return launch_mstep_d<16>(t, K, d_stats, stream, min_nk, pair_out);
        case 20: return launch_mstep_d<20>(t, K, d_stats, stream, min_nk, pair_out);
        case 24: return launch_mstep_d<24>(t, K, d_stats, stream, min_nk, pair_out);
        default: return fail(GMM_ERR_ARG, "tensor-core M-step: unsupported D");
    }
}

}  // namespace gmm
