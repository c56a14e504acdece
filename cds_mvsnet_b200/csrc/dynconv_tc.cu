// DynamicConv (A6) on the tensor cores: tcgen05.mma + TMEM + TMA, warp specialised.
//
// Reference: models/dynamic_conv.py:97-122.  One launch evaluates every kernel-size branch of a layer for a
// batch of images as tap GEMMs without im2col (see tc_common.cuh):
//   M = 128 consecutive pixels of an image row (a "row unit"; 128 - 2*halo of them are valid outputs),
//   N = per branch 8 feature channels + 3 curvature channels (a,b,c) (+ 3 columns carrying the fp16 rounding
//       residual of the curvature weights: the gate softmax(g/T) amplifies curvature error 100x), padded to 16;
//       ALL branches side by side (N = 32 or 48), embedded in the kmax x kmax tap grid,
//   K = kmax*kmax taps x 8 channels (one 8-channel slab per tap; two slabs per K=16 MMA).
// A CTA owns TY row units.  Warp roles (192 threads):
//   warp 0      TMA producer: one cp.async.bulk.tensor for the haloed [TY+2h][128] pixel window (2 KB rows,
//               zero fill outside the image = conv padding) + one bulk copy of the packed fp16 weights
//   all warps   optional in-place InstanceNorm + LeakyReLU of the window (the PRODUCER layer's norm, applied
//               by the consumer so it never costs a pass over HBM), then fence.proxy.async
//   warp 1      MMA issuer: per row unit, all branches into one of two TMEM accumulator stages
//   warps 2..5  epilogue: tcgen05.ld the unit's accumulators, curvature -> gate softmax -> blend the branches,
//               write fp16 [n,H,W,8] + curvature maps, accumulate InstanceNorm statistics for the next layer
// The MMA of unit u+1 overlaps the epilogue of unit u (tmem_full / tmem_empty mbarriers).
//
// Activations: channels-last fp16 [n,H,W,C], C in {8,16,32} (conv00's image is padded to 8 channels).  For C = 8 a
// pixel row is one contiguous 2 KB TMA row; for C > 8 each 8-channel chunk is fetched by its own TMA box
// (16-byte inner extent) so that shared memory still holds one slab per chunk.
// Packed weights (host: weights.py pack_dynamic_conv_tc): first the MMAs of the inner KIN x KIN taps
// ([k-chunk 2][NK*16/8][8 n][8 k] fp16 each, branch b in columns [16b, 16b+16), zero where its kernel has no such
// tap), then the MMAs of the outer ring ([k-chunk 2][2][8 n][8 k], largest kernel only).
#include <algorithm>
#include <utility>

#include "cds_common.cuh"
#include "tc_common.cuh"
#include "tma_host.h"

namespace {

// resident CTAs per SM asked of the compiler for the k=(1,3) layers with <= 16 channels (out3, out2): their epilogue is
// latency-bound (ncu: 57 % issue-active at 28 % warp occupancy), their TMEM footprint (128 columns) allows four
#ifndef CDS_K13_CTAS
#define CDS_K13_CTAS 4
#endif
#ifndef CDS_K357_CTAS
#define CDS_K357_CTAS 3   // conv01
#endif
constexpr int TX = 128;
constexpr int ROW_BYTES = TX * 16;
constexpr float kInEps = 1e-5f;

template <int K0_, int K1_, int K2_, int CIN_, int COUT_>
struct Cfg {
    static constexpr int NK = K2_ > 0 ? 3 : 2;
    static constexpr int CIN = CIN_, COUT = COUT_, C8 = CIN_ / 8;
    static constexpr int K0 = K0_, K1 = K1_, K2 = K2_;
    static constexpr int KMAX = K2_ > K1_ ? (K2_ > K0_ ? K2_ : K0_) : (K1_ > K0_ ? K1_ : K0_);
    static constexpr int HALO = (KMAX - 1) / 2;
    static constexpr int TXO = TX - 2 * HALO;    // valid outputs per row unit
    // per branch: COUT feature columns, (a,b,c) from fp16-rounded curvature weights, (a,b,c) from their residuals
    // Every layer also carries the fp16 rounding RESIDUAL of the feature weights in COUT extra columns per branch, summed in
    // the epilogue (effectively fp32-accurate weights): on the chaotic "noise" input the stage-1 depth amplifies every
    // rounding of the trunk ~9x into the final depth, so no trunk layer can afford fp16 weights
    static constexpr bool WLO = true;
    static constexpr int FCOLS = (WLO ? 2 : 1) * COUT_;         // feature columns per branch; curvature columns follow
    static constexpr int NPAD = (FCOLS + 6 + 15) / 16 * 16;
    // Branches are embedded in the KMAX x KMAX tap grid.  The tensor core re-reads the 4 KB A operand from shared
    // memory for every MMA (>= 32 cycles whatever N is), so one wide MMA per tap pair beats one narrow MMA per
    // branch and tap.  Taps inside the second-largest kernel's support (KIN x KIN) feed ALL branches (N = NK*NPAD,
    // zero weights where a smaller kernel has no tap); the outer ring only exists for the largest kernel (N = NPAD).
    static constexpr int KIN = NK == 3 ? (K1_ > K0_ ? K1_ : K0_) : K0_;
    static constexpr int NALL = NK * NPAD;
    static constexpr int NIN = KIN * KIN, NRING = KMAX * KMAX - KIN * KIN;   // taps (NIN odd, NRING even)
    // K = 16 per MMA = two 8-channel slabs: for C8 = 1 two consecutive taps (first inner MMA = [tap0, zero pad]),
    // for C8 > 1 two channel chunks of the same tap
    static constexpr int MMA_IN = C8 == 1 ? (NIN + 1) / 2 : NIN * C8 / 2;
    static constexpr int MMA_RING = C8 == 1 ? NRING / 2 : NRING * C8 / 2;
    static constexpr int NMMA = MMA_IN + MMA_RING;
    static constexpr int B_BYTES = MMA_IN * 2 * NALL * 16 + MMA_RING * 2 * NPAD * 16;
    // i-th tap (row-major order) of the outer ring, as an index into the KMAX x KMAX grid
    __host__ __device__ static constexpr int ring_tap(int i) {
        int lo = (KMAX - KIN) / 2, hi = lo + KIN, c = 0;
        for (int t = 0; t < KMAX * KMAX; ++t) {
            int y = t / KMAX, x = t % KMAX;
            bool inner = y >= lo && y < hi && x >= lo && x < hi;
            if (!inner) { if (c == i) return t; ++c; }
        }
        return 0;
    }
};

// byte offset of tap t of a k x k branch inside the window, relative to the row unit's first row
template <int HALO>
__host__ __device__ constexpr uint32_t tap_off(int k, int t) {
    return (uint32_t)((t / k + HALO - (k - 1) / 2) * ROW_BYTES + (t % k + HALO - (k - 1) / 2) * 16);
}
struct DynTcParams {
    const int* img_index;     // item n reads image img_index[n] of the tensor map (NULL: n)
    const double* in_stats;   // [n][CIN][2] (sum, sumsq) of the input, or NULL: input used as is
    const float* epipole;     // [n][2]
    const __half* wgt;        // packed fp16 B image
    const float* bias;        // [NK][COUT] or NULL
    const float* gate;        // W1f [4][NK], b1 [4], W2 [NK][4]
    __half* out_raw;          // [n][H][W][COUT]
    __half* out_lo;           // optional: fp16 rounding residual of out_raw (split-precision storage), same shape
    long long in_lo_images;   // SPLIT input: the residual plane is image index + in_lo_images of the same tensor map
    double* out_stats;        // [n][COUT][2] or NULL
    float* norm_curv;         // [n][H][W] or NULL
    float* nc_sq;             // [n][H][W] or NULL
    float* nc_abs;            // [n][H][W] or NULL
    int in_act, nc_mode, H, W;
    int pair_v, pair_b;       // GRP > 1: items are (side, v, b); the side-0 items v*pair_b + b of batch item b share one image
    float epi_scale, inv_temperature;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(tc::smem_u32(bar)) : "memory");
}

// MMA J of a tap group as a template parameter: forces every table lookup to a compile-time constant.
// RING = false: inner KIN x KIN taps, all branches (N = NALL); RING = true: outer ring, largest kernel only.
template <class C, int CHUNK, bool RING, bool SPLIT, int J>
__device__ __forceinline__ void issue_one(uint32_t a_base, uint32_t b_base, uint32_t acc_col, bool elected) {
    constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
    constexpr int N = RING ? C::NPAD : C::NALL;
    constexpr uint32_t idesc = tc::instr_desc_f16(128, N);
    constexpr uint32_t b_lo_const = ((uint32_t)(N * 16) >> 4) << 16;            // LBO of B: between its two k-chunks
    constexpr uint32_t b_off = (RING ? (uint32_t)(C::MMA_IN * 2 * C::NALL * 16) : 0u) + (uint32_t)J * (2 * N * 16);
    constexpr int k = RING ? C::KMAX : C::KIN;
    constexpr uint32_t off0 = C::C8 == 1
        ? (RING ? tap_off<C::HALO>(k, C::ring_tap(2 * J)) : (J == 0 ? tap_off<C::HALO>(k, 0) : tap_off<C::HALO>(k, 2 * J - 1)))
        : (uint32_t)((2 * J) % C::C8) * CHUNK + tap_off<C::HALO>(k, RING ? C::ring_tap((2 * J) / C::C8) : (2 * J) / C::C8);
    constexpr uint32_t lbo = C::C8 == 1
        ? (RING ? tap_off<C::HALO>(k, C::ring_tap(2 * J + 1)) - off0 : (J == 0 ? 16u : tap_off<C::HALO>(k, 2 * J) - off0))
        : (uint32_t)CHUNK;
    constexpr uint32_t a_const = (off0 >> 4) | ((lbo >> 4) << 16);
    constexpr uint32_t b_const = (b_off >> 4) | b_lo_const;
    constexpr bool first = !RING && J == 0;
    if (elected)
        tc::mma_f16(acc_col + (RING ? (C::NK - 1) * C::NPAD : 0), ((uint64_t)desc_hi << 32) | (a_base + a_const),
                    ((uint64_t)desc_hi << 32) | (b_base + b_const), idesc, !first);
    if constexpr (SPLIT) {
        // split-precision input: the same weights applied to the residual slabs, C8 chunks further on
        constexpr uint32_t lo_off = ((uint32_t)C::C8 * CHUNK) >> 4;
        if (elected)
            tc::mma_f16(acc_col + (RING ? (C::NK - 1) * C::NPAD : 0), ((uint64_t)desc_hi << 32) | (a_base + a_const + lo_off),
                        ((uint64_t)desc_hi << 32) | (b_base + b_const), idesc, true);
    }
}
template <class C, int CHUNK, bool RING, bool SPLIT, int... J>
__device__ __forceinline__ void issue_group(uint32_t a_base, uint32_t b_base, uint32_t acc_col, bool elected,
                                            std::integer_sequence<int, J...>) {
    (issue_one<C, CHUNK, RING, SPLIT, J>(a_base, b_base, acc_col, elected), ...);
}
template <class C, int CHUNK, bool SPLIT>
__device__ __forceinline__ void issue_unit(uint32_t a_base, uint32_t b_base, uint32_t acc_col, bool elected) {
    issue_group<C, CHUNK, false, SPLIT>(a_base, b_base, acc_col, elected, std::make_integer_sequence<int, C::MMA_IN>{});
    issue_group<C, CHUNK, true, SPLIT>(a_base, b_base, acc_col, elected, std::make_integer_sequence<int, C::MMA_RING>{});
}

// SPLIT: the input arrives as two fp16 planes (value + rounding residual, i.e. ~22-bit activations); both are normalised
// together, re-split and fed to the tensor cores as twice as many K slabs (same weights).  Used for the layers the
// depth output is most sensitive to (conv10, conv11).
// GRP > 1 (conv00): the batch is (side, v, b) pairs of (reference image seen with pair v's epipole, source image v).  The
// reference image's branch convolutions do not depend on the epipole -- only the gate does -- so one CTA
// runs the MMAs of a reference tile ONCE and its epilogue up to GRP times, once per pair: 5 instead of 8 images' worth of
// tensor work at N = 5.
template <class C, int TY, bool SPLIT, int GRP = 1>
__global__ void __launch_bounds__(192, (GRP > 1 ? 1 : (C::COUT <= 16 ? (SPLIT ? 2 : (C::KMAX == 3 ? CDS_K13_CTAS : (C::NK == 3 && C::KMAX == 7 ? 2 : 3))) : (SPLIT ? 1 : 2)))) dynconv_tc_kernel(const __grid_constant__ CUtensorMap tmap, DynTcParams p) {
    static_assert(GRP == 1 || (C::COUT == 8 && !SPLIT), "shared-image groups are implemented for the 8-channel image layer");
    constexpr int NK = C::NK, HALO = C::HALO, TXO = C::TXO, C8 = C::C8, CIN = C::CIN, COUT = C::COUT, NPAD = C::NPAD;
    constexpr int ROWS = TY + 2 * HALO;
    constexpr uint32_t CHUNK = ROWS * ROW_BYTES;                    // one 8-channel slab of the window
    constexpr uint32_t A_BYTES = (SPLIT ? 2 : 1) * C8 * CHUNK;
    constexpr uint32_t B_BYTES = C::B_BYTES;
    constexpr uint32_t STAGE_COLS = C::NALL;                        // accumulator columns per row unit
    // accumulator stages: two (the MMAs of unit u+1 overlap the epilogue of unit u) unless that would cost a resident CTA its
    // TMEM; measured: the MMA-heavy (3,5) layers prefer 2 stages x 2 CTAs, the k=(1,3) 16-channel layer 1 stage x 3 CTAs
    constexpr uint32_t NSTG = (2 * STAGE_COLS <= 128 || (C::KMAX >= 5 && 2 * STAGE_COLS <= 256)) ? 2 : 1;
    constexpr uint32_t TMEM_COLS = NSTG * STAGE_COLS <= 64 ? 64 : (NSTG * STAGE_COLS <= 128 ? 128 : 256);
    static_assert(NSTG * STAGE_COLS <= 256, "accumulators exceed the TMEM budget of two resident CTAs");
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
    uint64_t* bar_full = bar_load + 1;    // [2]
    uint64_t* bar_empty = bar_load + 3;   // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 5);
    float* s_norm = reinterpret_cast<float*>(bar_load + 6);   // [CIN][2] mean, rstd
    float* s_red = s_norm + 2 * CIN;                          // [GRP][4 warps][COUT][2]
    float* s_gate = s_red + GRP * 4 * COUT * 2;               // W1f [4][NK], b1 [4], W2 [NK][4]  (<= 28 floats)
    float* s_bias = s_gate + 28;                              // [NK][COUT]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // items of this CTA: n, n + nstr, ... (cnt of them); they all read image img_index[n]
    int n = blockIdx.z, cnt = 1, nstr = 0;
    if constexpr (GRP > 1) {
        const int parts = (p.pair_v + GRP - 1) / GRP, z = blockIdx.z;
        if (z < p.pair_b * parts) {
            const int v0 = (z / p.pair_b) * GRP;
            n = v0 * p.pair_b + z % p.pair_b;
            cnt = min(GRP, p.pair_v - v0);
            nstr = p.pair_b;
        } else {
            n = p.pair_v * p.pair_b + (z - p.pair_b * parts);
        }
    }
    const int x0 = max(0, min((int)blockIdx.x * TXO, p.W - TXO));   // last tile overlaps its neighbour; W < TXO: one partial tile
    const int y0 = blockIdx.y * TY;
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);

    if (warp == 0) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 32) {
        tc::mbar_init(bar_load, 1);
        tc::mbar_init(bar_full, 1);
        tc::mbar_init(bar_full + 1, 1);
        tc::mbar_init(bar_empty, 4);
        tc::mbar_init(bar_empty + 1, 4);
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
    }
    if (p.in_stats && threadIdx.x >= 64 && threadIdx.x < 64 + CIN) {
        int c = threadIdx.x - 64;
        double cnt = (double)p.H * p.W;
        double s = p.in_stats[((size_t)n * CIN + c) * 2], ss = p.in_stats[((size_t)n * CIN + c) * 2 + 1];
        double m = s / cnt, var = ss / cnt - m * m;
        if (var < 0.0) var = 0.0;
        s_norm[2 * c] = (float)m;
        s_norm[2 * c + 1] = (float)(1.0 / sqrt(var + (double)kInEps));
    }
    if (threadIdx.x >= 96 && threadIdx.x < 96 + 8 * NK + 4) s_gate[threadIdx.x - 96] = __ldg(p.gate + threadIdx.x - 96);
    if (threadIdx.x >= 128 && threadIdx.x < 128 + NK * COUT) s_bias[threadIdx.x - 128] = p.bias ? __ldg(p.bias + threadIdx.x - 128) : 0.f;
    for (int i = threadIdx.x; i < GRP * 4 * COUT * 2; i += 192) s_red[i] = 0.f;
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- producer: the haloed pixel window (one TMA per 8-channel chunk) and the weights ---------------------
    if (threadIdx.x == 0) {
        const int img = p.img_index ? __ldg(p.img_index + n) : n;
        tc::mbar_expect_tx(bar_load, A_BYTES + B_BYTES);
        if (C8 == 1) {
            tc::tma_load_4d(sA_u, &tmap, bar_load, 2 * (x0 - HALO), y0 - HALO, img, 0);
            if constexpr (SPLIT) tc::tma_load_4d(sA_u + CHUNK, &tmap, bar_load, 2 * (x0 - HALO), y0 - HALO, img + (int)p.in_lo_images, 0);
        } else {
#pragma unroll
            for (int c8 = 0; c8 < C8; ++c8) tc::tma_load_5d(sA_u + c8 * CHUNK, &tmap, bar_load, 0, c8, x0 - HALO, y0 - HALO, img);
            if constexpr (SPLIT) {
#pragma unroll
                for (int c8 = 0; c8 < C8; ++c8)
                    tc::tma_load_5d(sA_u + (C8 + c8) * CHUNK, &tmap, bar_load, 0, c8, x0 - HALO, y0 - HALO, img + (int)p.in_lo_images);
            }
        }
        tc::bulk_copy_g2s(sB_u, p.wgt, B_BYTES, bar_load);
    }
    tc::mbar_wait(bar_load, 0);

    // ---- the producer layer's InstanceNorm + activation, applied in place (zero padding stays zero) ---------
    if (p.in_stats) {
        for (int i = threadIdx.x; i < C8 * ROWS * TX; i += 192) {
            int px = i % TX, ry = (i / TX) % ROWS, c8 = i / (TX * ROWS);
            int gx = x0 - HALO + px, gy = y0 - HALO + ry;
            if (gx < 0 || gx >= p.W || gy < 0 || gy >= p.H) continue;
            uint4* q = reinterpret_cast<uint4*>(sA + (size_t)i * 16);
            uint4 raw = *q;
            __half2* h = reinterpret_cast<__half2*>(&raw);
            const float* nm = s_norm + c8 * 16;
            if constexpr (SPLIT) {
                uint4* ql = reinterpret_cast<uint4*>(sA + (size_t)C8 * CHUNK + (size_t)i * 16);
                uint4 rawl = *ql;
                __half2* hl = reinterpret_cast<__half2*>(&rawl);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 f = __half22float2(h[j]), g = __half22float2(hl[j]);
                    f.x = ((f.x + g.x) - nm[4 * j]) * nm[4 * j + 1];
                    f.y = ((f.y + g.y) - nm[4 * j + 2]) * nm[4 * j + 3];
                    if (p.in_act == 1) { f.x = f.x > 0.f ? f.x : 0.1f * f.x; f.y = f.y > 0.f ? f.y : 0.1f * f.y; }
                    h[j] = __floats2half2_rn(f.x, f.y);
                    float2 back = __half22float2(h[j]);
                    hl[j] = __floats2half2_rn(f.x - back.x, f.y - back.y);
                }
                *ql = rawl;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float2 f = __half22float2(h[j]);
                    f.x = (f.x - nm[4 * j]) * nm[4 * j + 1];
                    f.y = (f.y - nm[4 * j + 2]) * nm[4 * j + 3];
                    if (p.in_act == 1) { f.x = f.x > 0.f ? f.x : 0.1f * f.x; f.y = f.y > 0.f ? f.y : 0.1f * f.y; }
                    h[j] = __floats2half2_rn(f.x, f.y);
                }
            }
            *q = raw;
        }
        tc::fence_proxy_async();   // generic-proxy writes -> visible to the tensor core's operand reads
        __syncthreads();
    }

    if (warp == 1) {
        // ---- MMA issuer: the warp stays converged, one elected lane issues ------------------------------------------
        tc::tc_fence_after();
        const bool elected = tc::elect_one();
        const uint32_t tmem_u = tc::uniform(tmem);
#pragma unroll 1
        for (uint32_t u = 0; u < (uint32_t)TY; ++u) {
            const uint32_t s = u % NSTG;
            tc::mbar_wait(bar_empty + s, ((u / NSTG) & 1) ^ 1);   // epilogue has drained this accumulator stage
            tc::tc_fence_after();
            const uint32_t a_base = (sA_u + u * ROW_BYTES) >> 4;
            const uint32_t acc = tmem_u + s * STAGE_COLS;
            issue_unit<C, (int)CHUNK, SPLIT>(a_base, sB_u >> 4, acc, elected);
            if (elected) tc::mma_commit(bar_full + s);
            __syncwarp();
        }
    } else if (warp >= 2 && GRP > 1) {
        // ---- epilogue of a shared-image group: the accumulators of a row unit are pulled into registers once (which frees
        // the TMEM stage at once), then gated / blended / stored once per item of the group with that item's epipole ---------
        const int lg = warp & 3;
        const int r = lg * 32 + lane;
        const int gx = x0 + r;
        float ex[GRP], ey[GRP];
#pragma unroll
        for (int it = 0; it < GRP; ++it) {
            const int nn = n + min(it, cnt - 1) * nstr;
            ex[it] = __ldg(p.epipole + 2 * nn) * p.epi_scale;
            ey[it] = __ldg(p.epipole + 2 * nn + 1) * p.epi_scale;
        }
        float st_sum[GRP][COUT], st_sq[GRP][COUT];
#pragma unroll
        for (int it = 0; it < GRP; ++it)
#pragma unroll
            for (int c = 0; c < COUT; ++c) { st_sum[it][c] = 0.f; st_sq[it][c] = 0.f; }
#pragma unroll 1
        for (int u = 0; u < TY; ++u) {
            const int s = u % (int)NSTG;
            tc::mbar_wait(bar_full + s, (u / (int)NSTG) & 1);
            tc::tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)s * STAGE_COLS;
            uint32_t ar[NK][8], yr[NK][8], yl[NK][8];
#pragma unroll
            for (int b = 0; b < NK; ++b) {
                tc::tmem_ld8_nowait(taddr + b * NPAD + C::FCOLS, ar[b]);
                tc::tmem_ld8_nowait(taddr + b * NPAD, yr[b]);
                tc::tmem_ld8_nowait(taddr + b * NPAD + COUT, yl[b]);   // product with the weights' fp16 rounding residual
            }
            tc::tmem_ld_wait();
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + s);   // everything is in registers: the stage may be refilled
            const int gy = y0 + u;
            const bool valid = r < TXO && gy < p.H && gx < p.W && gx >= (int)blockIdx.x * TXO;
            float ca[NK], cb[NK], cc[NK], yb[NK][8];
#pragma unroll
            for (int b = 0; b < NK; ++b) {
                ca[b] = __uint_as_float(ar[b][0]) + __uint_as_float(ar[b][3]);
                cb[b] = __uint_as_float(ar[b][1]) + __uint_as_float(ar[b][4]);
                cc[b] = __uint_as_float(ar[b][2]) + __uint_as_float(ar[b][5]);
#pragma unroll
                for (int c = 0; c < 8; ++c) yb[b][c] = (__uint_as_float(yr[b][c]) + __uint_as_float(yl[b][c])) + s_bias[b * COUT + c];
            }
#pragma unroll
            for (int it = 0; it < GRP; ++it) {
                if (it < cnt) {   // warp-uniform
                    float uu = (float)gx - ex[it], vv = (float)gy - ey[it];
                    const float rinv = __frcp_rn(sqrtf(uu * uu + vv * vv) + 1e-6f);
                    uu *= rinv;
                    vv *= rinv;
                    float curv[NK];
#pragma unroll
                    for (int b = 0; b < NK; ++b) curv[b] = (ca[b] * (uu * uu) + cb[b] * (2.f * uu * vv)) + cc[b] * (vv * vv);
                    float hdn[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        float t = s_gate[4 * NK + j];
#pragma unroll
                        for (int b = 0; b < NK; ++b) t += s_gate[j * NK + b] * curv[b];
                        hdn[j] = fmaxf(t, 0.f);
                    }
                    float wgt[NK], mx = -INFINITY;
#pragma unroll
                    for (int b = 0; b < NK; ++b) {
                        float t = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) t += s_gate[4 * NK + 4 + b * 4 + j] * hdn[j];
                        wgt[b] = t * p.inv_temperature;
                        mx = fmaxf(mx, wgt[b]);
                    }
                    float den = 0.f, nc = 0.f;
#pragma unroll
                    for (int b = 0; b < NK; ++b) { wgt[b] = __expf(wgt[b] - mx); den += wgt[b]; }
                    const float dinv = __frcp_rn(den);
#pragma unroll
                    for (int b = 0; b < NK; ++b) { wgt[b] *= dinv; nc += curv[b] * wgt[b]; }
                    float out[8];
#pragma unroll
                    for (int c = 0; c < 8; ++c) out[c] = 0.f;
#pragma unroll
                    for (int b = 0; b < NK; ++b)
#pragma unroll
                        for (int c = 0; c < 8; ++c) out[c] += wgt[b] * yb[b][c];
                    if (valid) {
                        const size_t m = ((size_t)(n + it * nstr) * p.H + gy) * p.W + gx;
                        Vec8<__half>::store(p.out_raw + m * COUT, out);
                        if (p.out_lo) {   // split-precision storage: what fp16 rounding just dropped
                            float res[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) res[c] = out[c] - __half2float(__float2half_rn(out[c]));
                            Vec8<__half>::store(p.out_lo + m * COUT, res);
                        }
#pragma unroll
                        for (int c = 0; c < 8; ++c) { st_sum[it][c] += out[c]; st_sq[it][c] += out[c] * out[c]; }
                        if (p.norm_curv) p.norm_curv[m] = nc;
                        if (p.nc_sq) {
                            if (p.nc_mode == 0) p.nc_sq[m] = nc * nc;
                            else if (p.nc_mode == 1) p.nc_sq[m] = p.nc_sq[m] + nc * nc;
                            else p.nc_sq[m] = (p.nc_sq[m] + nc * nc) / 3.f;
                        }
                        if (p.nc_abs) p.nc_abs[m] = fabsf(nc);
                    }
                }
            }
        }
        if (p.out_stats) {
#pragma unroll
            for (int it = 0; it < GRP; ++it)
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                    const float a = warp_sum(st_sum[it][c]), q = warp_sum(st_sq[it][c]);
                    if (lane == 0) { s_red[((it * 4 + lg) * COUT + c) * 2] = a; s_red[((it * 4 + lg) * COUT + c) * 2 + 1] = q; }
                }
        }
    } else if (warp >= 2) {
        // ---- epilogue: gate + blend, one pixel per thread ---------------------------------------------------------
        const int lg = warp & 3;                 // TMEM lane group this warp may read
        const int r = lg * 32 + lane;            // MMA row = pixel x0 + r (valid while r < TXO)
        const int gx = x0 + r;
        const float ex = __ldg(p.epipole + 2 * n) * p.epi_scale, ey = __ldg(p.epipole + 2 * n + 1) * p.epi_scale;
        constexpr bool REG_STATS = true;   // per-thread statistics accumulators only while they fit in registers
        float st_sum[REG_STATS ? COUT : 1], st_sq[REG_STATS ? COUT : 1];
#pragma unroll
        for (int c = 0; c < (REG_STATS ? COUT : 1); ++c) { st_sum[c] = 0.f; st_sq[c] = 0.f; }

        // the curvature accumulator is read-modify-written per pixel: its old value is fetched one row unit ahead, so that
        // the DRAM round trip overlaps the previous unit's work instead of sitting in every unit's critical path
        auto load_nc = [&](int u) {
            const int gy = y0 + u;
            const bool ok = u < TY && r < TXO && gy < p.H && gx < p.W && gx >= (int)blockIdx.x * TXO;
            return (ok && p.nc_sq && p.nc_mode != 0) ? __ldcg(p.nc_sq + ((size_t)n * p.H + gy) * p.W + gx) : 0.f;
        };
        float nc_next = load_nc(0);
#pragma unroll 1
        for (int u = 0; u < TY; ++u) {
            const int s = u % (int)NSTG;
            const float nc_old = nc_next;
            nc_next = load_nc(u + 1);
            tc::mbar_wait(bar_full + s, (u / (int)NSTG) & 1);
            tc::tc_fence_after();
            const uint32_t taddr = tmem + ((uint32_t)(lg * 32) << 16) + (uint32_t)s * STAGE_COLS;
            // 1. curvature columns of every branch -> gate weights
            uint32_t ar[NK][8];
#pragma unroll
            for (int b = 0; b < NK; ++b) tc::tmem_ld8_nowait(taddr + b * NPAD + C::FCOLS, ar[b]);
            tc::tmem_ld_wait();

            const int gy = y0 + u;
            // tiles overlap at the right image edge (x0 is clamped): every pixel is owned by exactly one tile,
            // which matters for the read-modify-write curvature accumulators and the statistics
            const bool valid = r < TXO && gy < p.H && gx < p.W && gx >= (int)blockIdx.x * TXO;
            float uu = (float)gx - ex, vv = (float)gy - ey;
            float rinv = __frcp_rn(sqrtf(uu * uu + vv * vv) + 1e-6f);
            uu *= rinv;
            vv *= rinv;
            float curv[NK];
#pragma unroll
            for (int b = 0; b < NK; ++b) {
                // (a,b,c) from the fp16-rounded weights + from their rounding residuals
                float ca = __uint_as_float(ar[b][0]) + __uint_as_float(ar[b][3]);
                float cb = __uint_as_float(ar[b][1]) + __uint_as_float(ar[b][4]);
                float cc = __uint_as_float(ar[b][2]) + __uint_as_float(ar[b][5]);
                curv[b] = (ca * (uu * uu) + cb * (2.f * uu * vv)) + cc * (vv * vv);
            }
            float hdn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float t = s_gate[4 * NK + j];
#pragma unroll
                for (int b = 0; b < NK; ++b) t += s_gate[j * NK + b] * curv[b];
                hdn[j] = fmaxf(t, 0.f);
            }
            float wgt[NK], mx = -INFINITY;
#pragma unroll
            for (int b = 0; b < NK; ++b) {
                float t = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) t += s_gate[4 * NK + 4 + b * 4 + j] * hdn[j];
                wgt[b] = t * p.inv_temperature;
                mx = fmaxf(mx, wgt[b]);
            }
            float den = 0.f, nc = 0.f;
#pragma unroll
            for (int b = 0; b < NK; ++b) { wgt[b] = __expf(wgt[b] - mx); den += wgt[b]; }
            const float dinv = __frcp_rn(den);
#pragma unroll
            for (int b = 0; b < NK; ++b) { wgt[b] *= dinv; nc += curv[b] * wgt[b]; }

            // 2. feature columns, 8 channels at a time: blend the branches, store, statistics
            const size_t m = ((size_t)n * p.H + gy) * p.W + gx;
#pragma unroll
            for (int c8 = 0; c8 < COUT / 8; ++c8) {
                uint32_t yr[NK][8];
#pragma unroll
                for (int b = 0; b < NK; ++b) tc::tmem_ld8_nowait(taddr + b * NPAD + c8 * 8, yr[b]);
                if constexpr (C::WLO) {   // product with the feature weights' fp16 rounding residual
                    uint32_t yl[NK][8];
#pragma unroll
                    for (int b = 0; b < NK; ++b) tc::tmem_ld8_nowait(taddr + b * NPAD + COUT + c8 * 8, yl[b]);
                    tc::tmem_ld_wait();
#pragma unroll
                    for (int b = 0; b < NK; ++b)
#pragma unroll
                        for (int c = 0; c < 8; ++c) yr[b][c] = __float_as_uint(__uint_as_float(yr[b][c]) + __uint_as_float(yl[b][c]));
                }
                tc::tmem_ld_wait();
                float out[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) out[c] = 0.f;
#pragma unroll
                for (int b = 0; b < NK; ++b)
#pragma unroll
                    for (int c = 0; c < 8; ++c) out[c] += wgt[b] * (__uint_as_float(yr[b][c]) + s_bias[b * COUT + c8 * 8 + c]);
                if (valid) {
                    Vec8<__half>::store(p.out_raw + m * COUT + c8 * 8, out);
                    if (p.out_lo) {   // split-precision storage: what fp16 rounding just dropped
                        float res[8];
#pragma unroll
                        for (int c = 0; c < 8; ++c) res[c] = out[c] - __half2float(__float2half_rn(out[c]));
                        Vec8<__half>::store(p.out_lo + m * COUT + c8 * 8, res);
                    }
                }
                if (p.out_stats) {
                    if constexpr (REG_STATS) {
                        if (valid) {
#pragma unroll
                            for (int c = 0; c < 8; ++c) { st_sum[c8 * 8 + c] += out[c]; st_sq[c8 * 8 + c] += out[c] * out[c]; }
                        }
                    } else {
#pragma unroll
                        for (int c = 0; c < 8; ++c) {
                            float a = warp_sum(valid ? out[c] : 0.f), q = warp_sum(valid ? out[c] * out[c] : 0.f);
                            if (lane == 0) { s_red[(lg * COUT + c8 * 8 + c) * 2] += a; s_red[(lg * COUT + c8 * 8 + c) * 2 + 1] += q; }
                        }
                    }
                }
            }
            tc::tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_empty + s);   // every accumulator column has been read: the stage may be refilled

            if (valid) {
                if (p.norm_curv) p.norm_curv[m] = nc;
                if (p.nc_sq) {
                    if (p.nc_mode == 0) p.nc_sq[m] = nc * nc;
                    else if (p.nc_mode == 1) p.nc_sq[m] = nc_old + nc * nc;
                    else p.nc_sq[m] = (nc_old + nc * nc) / 3.f;
                }
                if (p.nc_abs) p.nc_abs[m] = fabsf(nc);
            }
        }
        if constexpr (REG_STATS) {
            if (p.out_stats) {
#pragma unroll
                for (int c = 0; c < COUT; ++c) {
                    float a = warp_sum(st_sum[c]), q = warp_sum(st_sq[c]);
                    if (lane == 0) { s_red[(lg * COUT + c) * 2] = a; s_red[(lg * COUT + c) * 2 + 1] = q; }
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (p.out_stats && threadIdx.x < cnt * COUT * 2) {
        const int it = threadIdx.x / (COUT * 2), j = threadIdx.x % (COUT * 2);
        double t = 0.0;
        for (int w = 0; w < 4; ++w) t += (double)s_red[(it * 4 + w) * COUT * 2 + j];
        atomicAdd(p.out_stats + (size_t)(n + it * nstr) * COUT * 2 + j, t);
    }
    if (warp == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
}

// fp32 planar images [n,3,H,W] -> fp16 [n,H,W,8] = (r, g, b, r_lo, g_lo, b_lo, 0, 0): each colour as an fp16 value plus the
// fp16 rounding residual.  conv00's weights are duplicated for the residual channels (its 8-channel operand slab has
// 5 spare channels anyway), so the tensor cores see the image to ~22 bits -- rounding the image itself to fp16 is the
// largest single term of the fp16 error budget.
__global__ void image_to_nhwc8_kernel(const float* __restrict__ img, long long HW, __half* __restrict__ out) {
    const int n = blockIdx.y;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < HW; i += (long long)gridDim.x * blockDim.x) {
        float v[8];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float x = __ldg(img + ((size_t)n * 3 + c) * HW + i);
            float hi = __half2float(__float2half_rn(x));
            v[c] = hi;
            v[3 + c] = x - hi;
        }
        v[6] = v[7] = 0.f;
        Vec8<__half>::store(out + ((size_t)n * HW + i) * 8, v);
    }
}

template <class C, int TY, bool SPLIT = false, int GRP = 1>
int launch_dyn_tc(const void* x, int n_images, const DynTcParams& p, int n, cudaStream_t st) {
    constexpr size_t smem = (size_t)(SPLIT ? 2 : 1) * C::C8 * (TY + 2 * C::HALO) * ROW_BYTES + (size_t)C::B_BYTES + 8 * 6 +
                            (2 * C::CIN + GRP * 8 * C::COUT + 28 + C::NK * C::COUT) * 4 + 16;
    static_assert(smem <= 227 * 1024, "tile does not fit in shared memory");
    auto kern = dynconv_tc_kernel<C, TY, SPLIT, GRP>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cds_set_error("cds_dynamic_conv_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    CUtensorMap tmap;
    const uint64_t H = p.H, W = p.W, NI = n_images;
    bool ok;
    if (C::C8 == 1) {   // 16-byte pixels: 8-byte elements so that a box row is 2 KB (128 pixels)
        const uint64_t dims[4] = {2 * W, H, (SPLIT ? 2 : 1) * NI, 1};   // SPLIT: the residual plane follows
        const uint64_t strides[4] = {0, W * 16, H * W * 16, (SPLIT ? 2 : 1) * NI * H * W * 16};
        const uint32_t box[4] = {2 * TX, (uint32_t)(TY + 2 * C::HALO), 1, 1};
        ok = tma::make_u64(&tmap, x, 4, dims, strides, box);
    } else {            // (8 ch, chunk, W, H, image): one box per 8-channel chunk lands as a slab
        const uint64_t dims[5] = {8, (uint64_t)C::C8, W, H, (SPLIT ? 2 : 1) * NI};   // SPLIT: residual plane follows
        const uint64_t strides[5] = {0, 16, (uint64_t)C::CIN * 2, W * C::CIN * 2, H * W * C::CIN * 2};
        const uint32_t box[5] = {8, 1, TX, (uint32_t)(TY + 2 * C::HALO), 1};
        ok = tma::make_f16(&tmap, x, 5, dims, strides, box);
    }
    if (!ok) return CDS_EUNSUPPORTED;
    // GRP > 1: one CTA column per group of up to GRP same-image items + one per remaining (source image) item
    const int nz = GRP > 1 ? p.pair_b * ((p.pair_v + GRP - 1) / GRP) + p.pair_v * p.pair_b : n;
    dim3 grid(cds_div_up(p.W, C::TXO), cds_div_up(p.H, TY), nz);
    kern<<<grid, 192, smem, st>>>(tmap, p);
    return cds_check_launch("cds_dynamic_conv_tc");
}

// the layer shapes of the feature extractor (models/module.py:211-234); 0 = not covered
int layer_id(int Cin, int Cout, int nk, const int* ks) {
    if (!ks) return 0;
    if (Cin == 8 && Cout == 8 && nk == 3 && ks[0] == 3 && ks[1] == 7 && ks[2] == 11) return 1;   // conv00 (image padded)
    if (Cin == 8 && Cout == 8 && nk == 3 && ks[0] == 3 && ks[1] == 5 && ks[2] == 7) return 2;    // conv01
    if (Cin == 8 && Cout == 8 && nk == 2 && ks[0] == 1 && ks[1] == 3) return 3;                  // out3
    if (Cin == 16 && Cout == 16 && nk == 2 && ks[0] == 3 && ks[1] == 5) return 4;                // conv10, conv11
    if (Cin == 16 && Cout == 16 && nk == 2 && ks[0] == 1 && ks[1] == 3) return 5;                // out2
    if (Cin == 32 && Cout == 32 && nk == 2 && ks[0] == 1 && ks[1] == 3) return 6;                // conv20, conv21, out1
    return 0;
}

}  // namespace

extern "C" {

int cds_image_to_nhwc8(const float* img, int n, int H, int W, void* out, cudaStream_t stream) {
    CDS_REQUIRE(img && out && n > 0 && n <= 65535 && H > 0 && W > 0, CDS_EARG, "cds_image_to_nhwc8: bad arguments");
    long long HW = (long long)H * W;
    dim3 grid((unsigned)std::min<long long>(148 * 8, (HW + 255) / 256), n);
    image_to_nhwc8_kernel<<<grid, 256, 0, stream>>>(img, HW, (__half*)out);
    return cds_check_launch("cds_image_to_nhwc8");
}

// 1 when the tensor-core DynamicConv covers this layer (the feature extractor's shapes, W >= 8)
int cds_dynamic_conv_tc_supported(int Cin, int Cout, int H, int W, int num_kernels, const int* ks) {
    if (W < 8 || H < 1) return 0;
    return layer_id(Cin, Cout, num_kernels, ks) != 0;
}

int cds_dynamic_conv_tc_weight_halfs(int Cin, int Cout, int num_kernels, const int* ks) {
    int kmax = 0, kin = 0, c8 = Cin / 8, npad = (2 * Cout + 6 + 15) / 16 * 16;
    for (int i = 0; i < num_kernels; ++i) kmax = ks[i] > kmax ? ks[i] : kmax;
    for (int i = 0; i < num_kernels; ++i) if (ks[i] < kmax && ks[i] > kin) kin = ks[i];
    int nin = kin * kin, nring = kmax * kmax - nin;
    int mma_in = c8 == 1 ? (nin + 1) / 2 : nin * c8 / 2, mma_ring = c8 == 1 ? nring / 2 : nring * c8 / 2;
    return (mma_in * 2 * (num_kernels * npad) + mma_ring * 2 * npad) * 8;
}

static int dynamic_conv_tc_impl(const void* x, int n_images, const int* img_index, const double* in_stats, int in_act,
                        const float* epipole, float epi_scale, const void* wgt_packed, const float* bias, const float* gate,
                        int n, int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes, float temperature,
                        int split_in, void* out_raw, void* out_lo, double* out_stats, float* norm_curv, float* nc_sq,
                        int nc_mode, float* nc_abs, int pair_v, int pair_b, cudaStream_t stream);

int cds_dynamic_conv_tc(const void* x, int n_images, const int* img_index, const double* in_stats, int in_act,
                        const float* epipole, float epi_scale, const void* wgt_packed, const float* bias, const float* gate,
                        int n, int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes, float temperature,
                        int split_in, void* out_raw, void* out_lo, double* out_stats, float* norm_curv, float* nc_sq,
                        int nc_mode, float* nc_abs, cudaStream_t stream) {
    return dynamic_conv_tc_impl(x, n_images, img_index, in_stats, in_act, epipole, epi_scale, wgt_packed, bias, gate, n, Cin, Cout, H, W,
                                num_kernels, kernel_sizes, temperature, split_in, out_raw, out_lo, out_stats, norm_curv, nc_sq, nc_mode,
                                nc_abs, 0, 0, stream);
}

// The image layer (conv00) over the (side, v, b) pair batch of the cascade: n = 2*V*B items, item (0, v, b) = the reference
// image of batch item b seen with pair v's epipole, item (1, v, b) = source image v; img_index must map the V side-0 items of
// a batch item to one image.  Same results as cds_dynamic_conv_tc; the reference image's convolutions run once per batch item.
int cds_dynamic_conv_tc_pairs(const void* x, int n_images, const int* img_index, const float* epipole, float epi_scale,
                              const void* wgt_packed, const float* bias, const float* gate, int V, int B, int Cin, int Cout, int H, int W,
                              int num_kernels, const int* kernel_sizes, float temperature, void* out_raw, void* out_lo, double* out_stats,
                              float* norm_curv, float* nc_sq, int nc_mode, float* nc_abs, cudaStream_t stream) {
    CDS_REQUIRE(V >= 1 && B >= 1 && img_index, CDS_EARG, "cds_dynamic_conv_tc_pairs: bad pair batch");
    return dynamic_conv_tc_impl(x, n_images, img_index, nullptr, 0, epipole, epi_scale, wgt_packed, bias, gate, 2 * V * B, Cin, Cout, H, W,
                                num_kernels, kernel_sizes, temperature, 0, out_raw, out_lo, out_stats, norm_curv, nc_sq, nc_mode, nc_abs,
                                V, B, stream);
}

static int dynamic_conv_tc_impl(const void* x, int n_images, const int* img_index, const double* in_stats, int in_act,
                        const float* epipole, float epi_scale, const void* wgt_packed, const float* bias, const float* gate,
                        int n, int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes, float temperature,
                        int split_in, void* out_raw, void* out_lo, double* out_stats, float* norm_curv, float* nc_sq,
                        int nc_mode, float* nc_abs, int pair_v, int pair_b, cudaStream_t stream) {
    CDS_REQUIRE(x && epipole && wgt_packed && gate && out_raw && kernel_sizes, CDS_EARG, "cds_dynamic_conv_tc: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && n_images > 0, CDS_ESHAPE, "cds_dynamic_conv_tc: bad batch");
    CDS_REQUIRE(temperature > 0.f, CDS_EARG, "cds_dynamic_conv_tc: temperature must be positive");
    CDS_REQUIRE(cds_dynamic_conv_tc_supported(Cin, Cout, H, W, num_kernels, kernel_sizes), CDS_EUNSUPPORTED,
                "cds_dynamic_conv_tc: unsupported layer (Cin=%d Cout=%d W=%d): needs a feature-extractor layer shape",
                Cin, Cout, W);
    DynTcParams p{};
    p.img_index = img_index; p.in_stats = in_stats; p.epipole = epipole; p.wgt = (const __half*)wgt_packed; p.bias = bias;
    p.gate = gate; p.out_raw = (__half*)out_raw; p.out_stats = out_stats; p.norm_curv = norm_curv; p.nc_sq = nc_sq;
    p.nc_abs = nc_abs; p.in_act = in_act; p.nc_mode = nc_mode; p.H = H; p.W = W; p.epi_scale = epi_scale;
    p.inv_temperature = 1.f / temperature;
    p.out_lo = (__half*)out_lo;
    p.in_lo_images = n_images;
    p.pair_v = pair_v; p.pair_b = pair_b;
    const int lid = layer_id(Cin, Cout, num_kernels, kernel_sizes);
    if (pair_v > 0) {
        CDS_REQUIRE(lid == 1 && !in_stats, CDS_EUNSUPPORTED, "cds_dynamic_conv_tc_pairs: implemented for the image layer (3,7,11)");
        return launch_dyn_tc<Cfg<3, 7, 11, 8, 8>, 16, false, 4>(x, n_images, p, n, stream);
    }
    if (split_in) {
        CDS_REQUIRE((lid == 2 || lid == 4 || lid == 6) && in_stats, CDS_EUNSUPPORTED,
                    "cds_dynamic_conv_tc: split-precision input is implemented for the trunk layers 8->8 (3,5,7), 16->16 (3,5) and "
                    "32->32 (1,3), with input statistics");
        if (lid == 2) return launch_dyn_tc<Cfg<3, 5, 7, 8, 8>, 8, true>(x, n_images, p, n, stream);
        if (lid == 4) return launch_dyn_tc<Cfg<3, 5, 0, 16, 16>, 4, true>(x, n_images, p, n, stream);
        return launch_dyn_tc<Cfg<1, 3, 0, 32, 32>, 4, true>(x, n_images, p, n, stream);
    }
    switch (lid) {
        case 1: return launch_dyn_tc<Cfg<3, 7, 11, 8, 8>, 8>(x, n_images, p, n, stream);
        case 2: return launch_dyn_tc<Cfg<3, 5, 7, 8, 8>, 16>(x, n_images, p, n, stream);
        case 3: return launch_dyn_tc<Cfg<1, 3, 0, 8, 8>, 16>(x, n_images, p, n, stream);
        case 4: return launch_dyn_tc<Cfg<3, 5, 0, 16, 16>, 8>(x, n_images, p, n, stream);
        case 5: return launch_dyn_tc<Cfg<1, 3, 0, 16, 16>, 8>(x, n_images, p, n, stream);
        default: return launch_dyn_tc<Cfg<1, 3, 0, 32, 32>, 4>(x, n_images, p, n, stream);
    }
}

}  // extern "C"
