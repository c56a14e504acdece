// The feature extractor's plain 2-D convolutions on the tensor cores (tcgen05.mma + TMEM):
//   * FeatureNet.downsample1/2: 3x3 stride-2 pad-1 conv (models/module.py:214,218)
//   * FeatureNet.inner1/inner2: 1x1 conv over cat(nearest-up2(a), b) (models/module.py:253-254,260-261)
// Both consume RAW producer outputs and apply the producer's InstanceNorm + LeakyReLU while staging their operand
// (models/module.py:66-69), and emit raw output + per-(image, channel) sum / sum-of-squares for the next consumer.
//
// One CTA walks kTiles tiles of 128 consecutive output pixels of one image; thread r owns pixel r of the tile:
//   1. gather (issued one tile ahead): for every K slab (tap x 8-channel chunk) the thread loads its pixel's 16 bytes, normalises, activates,
//      rounds to fp16 and stores them at [slab][r] of the tcgen05 K-major no-swizzle A image (zero outside the image:
//      the conv pads the ACTIVATED tensor);
//   2. one elected thread issues the MMAs (M = 128, K = 16 = two slabs, N = 2*Cout) into TMEM and commits;
//   3. epilogue: TMEM -> registers, fp16 store, running statistics; per-CTA statistics go out as fp64 atomics.
// Precision: these layers ran fp32 operands on the CUDA cores and the depth tolerance has no slack,
// so both operands are split: N = 2*Cout, columns [Cout, 2Cout) multiply the fp16 rounding residual of the weights and are
// summed in the epilogue; the normalised activations are staged as an fp16 value slab AND an fp16 residual slab (~22 bits),
// the residual slabs being extra K fed through the same weight image.  The MMA count is tiny here, so this is cheap.
// Weights (host: weights.py pack_conv2d_gtc): [mma][k-chunk 2][N/8][8 n][8 k] fp16, slabs tap-major / chunk-minor,
// zero-padded to an even slab count.
#include "cds_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kTiles = 4;          // tiles per CTA (amortises weights, fp64 norm coefficients, TMEM allocation)
constexpr float kInEps = 1e-5f;
constexpr int kSlab = 128 * 16;

struct G2Params {
    const __half* a;          // MODE 0: input [n,Hi,Wi,CA]; MODE 1: half-resolution input [n,H/2,W/2,CA]
    const __half* a_lo;       // MODE 0, optional: fp16 rounding residual plane of `a` (split-precision storage), same shape
    const double* a_stats;    // [n][CA][2] or null (input used as is)
    const __half* b;          // MODE 1: full-resolution input [n,H,W,CB]
    const double* b_stats;
    const __half* wgt;
    __half* out;              // [n,Ho,Wo,COUT] raw
    __half* out_lo;           // optional: fp16 rounding residual of `out` (split-precision storage), same shape
    double* out_stats;        // [n][COUT][2] or null
    int a_act, b_act;
    int Hi, Wi;               // MODE 0: input extent
    int Ho, Wo;               // output extent
};

__device__ __forceinline__ float act_apply(float v, int act) { return act == 1 ? (v > 0.f ? v : 0.1f * v) : v; }

// 8 channels: normalise, activate, round to fp16 (returned) + the fp16 rounding residual (lo).  On entry `lo` holds the
// residual plane of the stored input (zeros when the producer kept one plane only).
__device__ __forceinline__ uint4 norm8(uint4 raw, const float* nm /* [8][2] mean, rstd */, int act, uint4& lo) {
    __half2* h = reinterpret_cast<__half2*>(&raw);
    __half2* l = reinterpret_cast<__half2*>(&lo);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        float2 f = __half22float2(h[j]);
        const float2 g = __half22float2(l[j]);
        f.x += g.x;
        f.y += g.y;
        f.x = act_apply((f.x - nm[4 * j]) * nm[4 * j + 1], act);
        f.y = act_apply((f.y - nm[4 * j + 2]) * nm[4 * j + 3], act);
        h[j] = __floats2half2_rn(f.x, f.y);
        const float2 back = __half22float2(h[j]);
        l[j] = __floats2half2_rn(f.x - back.x, f.y - back.y);
    }
    return raw;
}

// MODE 0: 3x3 s2 over `a` (CB = 0).  MODE 1: 1x1 over cat(up2(a), b).
template <int MODE, int CA, int CB, int COUT>
struct G2Cfg {
    static constexpr int CIN = CA + CB;
    static constexpr int C8 = CIN / 8;
    static constexpr int NREAL = MODE == 0 ? 9 * C8 : C8;      // real K slabs
    static constexpr int NSLAB = (NREAL + 1) / 2 * 2;
    static constexpr int N = 2 * COUT;
    static constexpr uint32_t A_BYTES = 2 * NSLAB * kSlab;     // value slabs, then residual slabs
    static constexpr uint32_t B_BYTES = (NSLAB / 2) * N * 32;
    static constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    static constexpr size_t SMEM = A_BYTES + B_BYTES + (2 * CIN + 4 * COUT * 2) * 4 + 32;
};

template <int MODE, int CA, int CB, int COUT>
__global__ void __launch_bounds__(128) conv2d_gtc_kernel(const G2Params p) {
    using G = G2Cfg<MODE, CA, CB, COUT>;
    constexpr int CIN = G::CIN, C8 = G::C8, NSLAB = G::NSLAB, NREAL = G::NREAL, N = G::N;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + G::A_BYTES;
    float* s_norm = reinterpret_cast<float*>(sB + G::B_BYTES);   // [CIN][2]
    float* s_red = s_norm + 2 * CIN;                             // [4 warps][COUT][2]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + 4 * COUT * 2);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, r = threadIdx.x;
    const int n = blockIdx.y;
    const int P = p.Ho * p.Wo;

    if (warp == 0) tc::tmem_alloc(tmem_slot, G::TMEM_COLS);
    if (threadIdx.x == 32) { tc::mbar_init(bar, 1); tc::mbar_fence_init(); }
    // weights -> smem (generic proxy; fenced below together with the first tile's A image)
    for (uint32_t o = r * 16; o < G::B_BYTES; o += 128 * 16)
        *reinterpret_cast<uint4*>(sB + o) = __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(p.wgt) + o));
    // producer InstanceNorm coefficients (fp64 statistics -> mean, 1/std)
    for (int c = threadIdx.x; c < CIN; c += 128) {
        const bool is_a = c < CA;
        const double* st = is_a ? p.a_stats : p.b_stats;
        float mean = 0.f, rstd = 1.f;
        if (st) {
            const int cc = is_a ? c : c - CA, C = is_a ? CA : CB;
            const double cnt = MODE == 0 ? (double)p.Hi * p.Wi : (is_a ? (double)(p.Ho / 2) * (p.Wo / 2) : (double)p.Ho * p.Wo);
            const double s = st[((size_t)n * C + cc) * 2], ss = st[((size_t)n * C + cc) * 2 + 1];
            const double m = s / cnt;
            double var = ss / cnt - m * m;
            if (var < 0.0) var = 0.0;
            mean = (float)m;
            rstd = (float)(1.0 / sqrt(var + (double)kInEps));
        }
        s_norm[2 * c] = mean;
        s_norm[2 * c + 1] = rstd;
    }
    // zero-weight pad slab: its A image only has to be finite
    if (NSLAB > NREAL) {
        *reinterpret_cast<uint4*>(sA + (NSLAB - 1) * kSlab + r * 16) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(sA + (2 * NSLAB - 1) * kSlab + r * 16) = make_uint4(0, 0, 0, 0);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_u = tc::uniform(tmem);
    const bool elected = tc::elect_one();

    float st_sum[COUT], st_sq[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { st_sum[c] = 0.f; st_sq[c] = 0.f; }

    const __half* a_img = p.a + (size_t)n * (MODE == 0 ? (size_t)p.Hi * p.Wi : (size_t)(p.Ho / 2) * (p.Wo / 2)) * CA;
    const __half* b_img = MODE == 1 ? p.b + (size_t)n * P * CB : nullptr;

    // gather of one tile: the thread's NREAL 16-byte pieces (zeros where the tap falls outside the image / the tile ends)
    constexpr int NLO = MODE == 0 ? NREAL : 1;   // residual-plane pieces (3x3 s2 layers only)
    const bool has_lo = MODE == 0 && p.a_lo != nullptr;
    const __half* a_lo_img = has_lo ? p.a_lo + (size_t)n * (size_t)p.Hi * p.Wi * CA : nullptr;
    auto gather = [&](int t, uint4 (&raw)[NREAL], uint4 (&rlo)[NLO], bool (&ok)[NREAL]) {
        const int m = (blockIdx.x * kTiles + t) * 128 + r;
        const bool live = t < kTiles && m < P;
        const int ox = live ? m % p.Wo : 0, oy = live ? m / p.Wo : 0;
#pragma unroll
        for (int s = 0; s < NREAL; ++s) {
            const __half* src;
            if (MODE == 0) {
                const int tap = s / C8, c8 = s % C8;
                const int iy = 2 * oy - 1 + tap / 3, ix = 2 * ox - 1 + tap % 3;
                ok[s] = live && (unsigned)iy < (unsigned)p.Hi && (unsigned)ix < (unsigned)p.Wi;
                src = a_img + ((size_t)iy * p.Wi + ix) * CA + c8 * 8;
            } else {
                ok[s] = live;
                src = s < CA / 8 ? a_img + ((size_t)(oy / 2) * (p.Wo / 2) + ox / 2) * CA + s * 8
                                 : b_img + (size_t)m * CB + (s - CA / 8) * 8;
            }
            raw[s] = ok[s] ? __ldg(reinterpret_cast<const uint4*>(src)) : make_uint4(0, 0, 0, 0);
            if (MODE == 0) rlo[s] = (ok[s] && has_lo) ? __ldg(reinterpret_cast<const uint4*>(a_lo_img + (src - a_img))) : make_uint4(0, 0, 0, 0);
        }
    };
    // the NEXT tile's loads are issued before this tile's MMAs / epilogue (when they fit in registers), so that the DRAM
    // round trip is not in every tile's critical path
    constexpr bool PREFETCH = NREAL <= 9;
    uint4 raw[NREAL], rlo[NLO];
    bool ok[NREAL];
    gather(0, raw, rlo, ok);

#pragma unroll 1
    for (int t = 0; t < kTiles; ++t) {
        const int m0 = (blockIdx.x * kTiles + t) * 128;
        if (m0 >= P) break;   // block-uniform
        const int m = m0 + r;
        const bool live = m < P;

        // ---- 1. normalise -> A image --------------------------------------------------------------------------------
        if (!PREFETCH && t > 0) gather(t, raw, rlo, ok);
#pragma unroll
        for (int s = 0; s < NREAL; ++s) {
            const int c8 = MODE == 0 ? s % C8 : s;
            const int act = (MODE == 1 && c8 >= CA / 8) ? p.b_act : p.a_act;
            uint4 lo = MODE == 0 ? rlo[s] : make_uint4(0, 0, 0, 0);
            const uint4 v = ok[s] ? norm8(raw[s], s_norm + c8 * 16, act, lo) : make_uint4(0, 0, 0, 0);
            *reinterpret_cast<uint4*>(sA + s * kSlab + r * 16) = v;
            *reinterpret_cast<uint4*>(sA + (NSLAB + s) * kSlab + r * 16) = lo;
        }
        if (PREFETCH) gather(t + 1, raw, rlo, ok);
        tc::fence_proxy_async();   // generic-proxy smem writes -> visible to the tensor core's operand reads
        tc::tc_fence_before();     // (the previous tile's TMEM loads are ordered before the MMAs below)
        __syncthreads();

        // ---- 2. MMAs ------------------------------------------------------------------------------------------------------
        if (warp == 0) {
            tc::tc_fence_after();
            constexpr uint32_t idesc = tc::instr_desc_f16(128, N);
#pragma unroll
            for (int j = 0; j < NSLAB; ++j) {   // value slab pairs, then the residual slab pairs through the same weights
                const uint64_t da = tc::smem_desc(sA_u + j * 2 * kSlab, kSlab, 128);
                const uint64_t db = tc::smem_desc(sB_u + (j % (NSLAB / 2)) * N * 32, N * 16, 128);
                if (elected) tc::mma_f16(tmem_u, da, db, idesc, j > 0);
            }
            if (elected) tc::mma_commit(bar);
            __syncwarp();
        }
        tc::mbar_wait(bar, t & 1);
        tc::tc_fence_after();

        // ---- 3. epilogue ------------------------------------------------------------------------------------------------
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            uint32_t hi[8], lo[8];
            tc::tmem_ld8_nowait(taddr + c8 * 8, hi);
            tc::tmem_ld8_nowait(taddr + COUT + c8 * 8, lo);
            tc::tmem_ld_wait();
            if (live) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    v[i] = __uint_as_float(hi[i]) + __uint_as_float(lo[i]);
                    st_sum[c8 * 8 + i] += v[i];
                    st_sq[c8 * 8 + i] += v[i] * v[i];
                }
                Vec8<__half>::store(p.out + ((size_t)n * P + m) * COUT + c8 * 8, v);
                if (p.out_lo) {   // what fp16 rounding just dropped
                    float res[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) res[i] = v[i] - __half2float(__float2half_rn(v[i]));
                    Vec8<__half>::store(p.out_lo + ((size_t)n * P + m) * COUT + c8 * 8, res);
                }
            }
        }
    }

    // ---- per-CTA statistics -> fp64 atomics ---------------------------------------------------------------------------------
    if (p.out_stats) {
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            const float s = warp_sum(st_sum[c]), q = warp_sum(st_sq[c]);
            if (lane == 0) { s_red[(warp * COUT + c) * 2] = s; s_red[(warp * COUT + c) * 2 + 1] = q; }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (p.out_stats && threadIdx.x < COUT * 2) {
        double tsum = 0.0;
        for (int w = 0; w < 4; ++w) tsum += (double)s_red[w * COUT * 2 + threadIdx.x];
        atomicAdd(p.out_stats + (size_t)n * COUT * 2 + threadIdx.x, tsum);
    }
    if (warp == 0) tc::tmem_dealloc(tmem, G::TMEM_COLS);
}

template <int MODE, int CA, int CB, int COUT>
int launch_g2(const G2Params& p, int n, cudaStream_t st) {
    using G = G2Cfg<MODE, CA, CB, COUT>;
    auto kern = conv2d_gtc_kernel<MODE, CA, CB, COUT>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) { cds_set_error("cds_conv2d_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(cds_div_up((long long)p.Ho * p.Wo, 128 * kTiles), n);
    kern<<<grid, 128, G::SMEM, st>>>(p);
    return cds_check_launch("cds_conv2d_tc");
}

}  // namespace

extern "C" {

int cds_conv2d_3x3s2_tc_supported(int Cin, int Cout) { return (Cin == 8 && Cout == 16) || (Cin == 16 && Cout == 32); }
int cds_conv2d_3x3s2_tc_weight_halfs(int Cin, int Cout) { return ((9 * (Cin / 8) + 1) / 2) * (2 * Cout) * 16; }

// in [n,H,W,Cin] fp16 raw (+ stats, activation of its producer) -> out [n,ceil(H/2),ceil(W/2),Cout] fp16 raw + out_stats;
// in_lo (optional, shape of in) is the fp16 rounding residual plane of in, out_lo (optional, shape of out) receives that of
// out (split-precision storage: value + residual = ~22-bit activations)
int cds_conv2d_3x3s2_tc(const void* in, const void* in_lo, const double* in_stats, int in_act, const void* wgt_packed, int n, int Cin,
                        int Cout, int H, int W, void* out, void* out_lo, double* out_stats, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt_packed && out, CDS_EARG, "cds_conv2d_3x3s2_tc: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && H > 0 && W > 0 && (long long)H * W < (1ll << 30), CDS_ESHAPE, "cds_conv2d_3x3s2_tc: bad shape");
    CDS_REQUIRE(cds_conv2d_3x3s2_tc_supported(Cin, Cout), CDS_EUNSUPPORTED, "cds_conv2d_3x3s2_tc: unsupported layer Cin=%d Cout=%d", Cin, Cout);
    G2Params p{};
    p.a = (const __half*)in; p.a_lo = (const __half*)in_lo; p.a_stats = in_stats; p.a_act = in_act; p.wgt = (const __half*)wgt_packed;
    p.out = (__half*)out; p.out_lo = (__half*)out_lo; p.out_stats = out_stats; p.Hi = H; p.Wi = W; p.Ho = (H + 1) / 2; p.Wo = (W + 1) / 2;
    if (Cin == 8) return launch_g2<0, 8, 0, 16>(p, n, stream);
    return launch_g2<0, 16, 0, 32>(p, n, stream);
}

int cds_conv2d_1x1_cat_tc_supported(int Ca, int Cb, int Cout) { return (Ca == 32 && Cb == 16 && Cout == 16) || (Ca == 16 && Cb == 8 && Cout == 8); }
int cds_conv2d_1x1_cat_tc_weight_halfs(int Ca, int Cb, int Cout) { return (((Ca + Cb) / 8 + 1) / 2) * (2 * Cout) * 16; }

// a [n,H/2,W/2,Ca], b [n,H,W,Cb] fp16 raw (+ their producers' stats / activations) -> out [n,H,W,Cout] fp16 raw + out_stats
int cds_conv2d_1x1_cat_tc(const void* a, const double* a_stats, int a_act, const void* b, const double* b_stats, int b_act,
                          const void* wgt_packed, int n, int Ca, int Cb, int Cout, int H, int W, void* out, double* out_stats,
                          cudaStream_t stream) {
    CDS_REQUIRE(a && b && wgt_packed && out, CDS_EARG, "cds_conv2d_1x1_cat_tc: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && (long long)H * W < (1ll << 30), CDS_ESHAPE,
                "cds_conv2d_1x1_cat_tc: bad shape (H, W must be even)");
    CDS_REQUIRE(cds_conv2d_1x1_cat_tc_supported(Ca, Cb, Cout), CDS_EUNSUPPORTED,
                "cds_conv2d_1x1_cat_tc: unsupported layer Ca=%d Cb=%d Cout=%d", Ca, Cb, Cout);
    G2Params p{};
    p.a = (const __half*)a; p.a_stats = a_stats; p.a_act = a_act; p.b = (const __half*)b; p.b_stats = b_stats; p.b_act = b_act;
    p.wgt = (const __half*)wgt_packed; p.out = (__half*)out; p.out_stats = out_stats; p.Ho = H; p.Wo = W;
    if (Ca == 32) return launch_g2<1, 32, 16, 16>(p, n, stream);
    return launch_g2<1, 16, 8, 8>(p, n, stream);
}

}  // extern "C"
