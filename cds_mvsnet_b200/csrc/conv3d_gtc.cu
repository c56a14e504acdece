// Gather-form tensor-core 3-D convolutions for the regulariser layers the TMA-window kernel (conv3d_tc.cu)
// cannot express: stride-2 Conv3d (conv1/3/5), the small deep stride-1 layers (conv4/conv6 at 1/4 and 1/8
// resolution) and the deep transposed conv (conv7, 64 -> 32).   tcgen05.mma + TMEM, operands staged by cp.async.
//
// Reference semantics: models/module.py:80-122 (Conv3d block: conv k3 p1 s{1,2} -> BN -> ReLU), :125-166
// (Deconv3d block: ConvTranspose3d k3 s2 p1 op1 -> BN -> ReLU, skip added afterwards at :310-312).
//
// Implicit GEMM, one CTA = 128 output rows (any 128 consecutive voxels of the linearised [B, Do, Ho, Wo] grid;
// thread r owns row r):
//   * per K step the 128 threads GATHER the rows' input voxels of one tap group straight into the tcgen05
//     K-major no-swizzle operand layout ([slab][128 rows][8 x fp16], one 16-byte cp.async per (row, 8-channel
//     slab), zero-filled outside the volume = the conv padding) plus that step's slice of the packed weights,
//     through a 3-stage ring; one elected thread issues the step's MMAs (K = 16 = two slabs each) and commits
//     them to the stage's mbarrier, which is what frees the stage for the gather two steps later;
//   * stride 2 is just a different row -> voxel map; the transposed conv runs as its 8 output-parity classes
//     (blockIdx.y), class (pd,ph,pw) being a dense conv over the 1..8 input neighbours that reach it
//     (out[2i-1+k] += in[i] w[k]: parity 0 <- k=1 from i; parity 1 <- k=2 from i and k=0 from i+1);
//   * N = 2*Cout: columns [Cout, 2Cout) multiply the fp16 rounding RESIDUAL of the folded weights and are summed
//     in the epilogue, so the weights are effectively fp32-accurate (these layers used to run fp32 weights on the
//     CUDA cores and the depth tolerance has no room to spare);
//   * epilogue: TMEM -> registers, + bias, ReLU (+ skip for the transposed conv), fp16, channel-blocked store.
// Activations: channel-blocked channels-last [B][C/8][D][H][W][8] fp16 (see conv3d.cu).
// Weights (host: weights.py pack_conv3d_gtc / pack_deconv3d_gtc): [mma][k-chunk 2][N/8][8 n][8 k] fp16, MMAs in
// tap-major / channel-chunk-minor slab order (Cin = 8: slabs [tap0, zero, tap1, ..., tap26]).
#include "cds_common.cuh"
#include "tc_common.cuh"

namespace {

constexpr int kStages = 3;
constexpr int kSlabBytes = 128 * 16;

struct GtcParams {
    const __half* in;
    const __half* wgt;
    const float* bias;
    const __half* skip;   // transposed conv only (may be null)
    __half* out;
    int B, Di, Hi, Wi;    // input extent
    int Ro, Rh, Rw;       // row grid: conv = output extent; transposed conv = input extent
    int relu;
    int ntaps[8];         // per class: tap-table entries (conv: class 0 only)
    int wofs[8];          // per class: index of its first MMA in the packed weights
    signed char off[8][28][3];   // per class, per tap: input offset (dz, dy, dx) relative to the row's base voxel
};

// MODE 0: conv stride 1, 1: conv stride 2, 2: transposed conv (one output-parity class per blockIdx.y)
template <int CIN, int COUT, int MODE>
struct GtcCfg {
    static constexpr int C8 = CIN / 8;
    static constexpr int N = 2 * COUT;
    static constexpr int SPS = C8 >= 4 ? C8 : (C8 == 2 ? (MODE == 2 ? 2 : 6) : 4);   // slabs per K step
    static constexpr int TPS = C8 == 1 ? SPS : SPS / C8;                               // tap-table entries per step
    static constexpr int MPS = SPS / 2;                                                // MMAs per step
    static constexpr uint32_t A_STAGE = SPS * kSlabBytes;
    static constexpr uint32_t B_MMA = N * 32;                                          // bytes of one MMA's B operand
    static constexpr uint32_t B_STAGE = MPS * B_MMA;
    static constexpr uint32_t STAGE = A_STAGE + B_STAGE;
    static constexpr uint32_t TMEM_COLS = N <= 32 ? 32 : (N <= 64 ? 64 : 128);
    static constexpr size_t SMEM = (size_t)kStages * STAGE + 64;
};

template <int CIN, int COUT, int MODE>
__global__ void __launch_bounds__(128) conv3d_gtc_kernel(const __grid_constant__ GtcParams p) {
    using G = GtcCfg<CIN, COUT, MODE>;
    constexpr int C8 = G::C8, N = G::N, SPS = G::SPS, TPS = G::TPS, MPS = G::MPS;
    extern __shared__ __align__(128) uint8_t smem[];
    uint64_t* bar_mma = reinterpret_cast<uint64_t*>(smem + kStages * G::STAGE);   // [kStages]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_mma + kStages);
    const uint32_t smem_u = tc::smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const int r = threadIdx.x;
    const int cls = MODE == 2 ? blockIdx.y : 0;
    const int nsteps = p.ntaps[cls] / TPS;

    if (warp == 0) tc::tmem_alloc(tmem_slot, G::TMEM_COLS);
    if (threadIdx.x == 32) {
#pragma unroll
        for (int s = 0; s < kStages; ++s) tc::mbar_init(bar_mma + s, 1);
        tc::mbar_fence_init();
    }

    // ---- this thread's row -----------------------------------------------------------------------------------
    const long long M = (long long)p.B * p.Ro * p.Rh * p.Rw;
    const long long m = (long long)blockIdx.x * 128 + r;
    const bool live = m < M;
    const long long mm = live ? m : M - 1;
    const int rx = (int)(mm % p.Rw), ry = (int)((mm / p.Rw) % p.Rh);
    const int rz = (int)((mm / ((long long)p.Rw * p.Rh)) % p.Ro), b = (int)(mm / ((long long)p.Rw * p.Rh * p.Ro));
    const int bz = MODE == 1 ? 2 * rz : rz, by = MODE == 1 ? 2 * ry : ry, bx = MODE == 1 ? 2 * rx : rx;
    const size_t Min = (size_t)p.Di * p.Hi * p.Wi;
    const __half* in_b = p.in + (size_t)b * C8 * Min * 8;
    const uint8_t* wgt_c = reinterpret_cast<const uint8_t*>(p.wgt) + (size_t)p.wofs[cls] * G::B_MMA;

    auto load_step = [&](int step, int stage) {
        const uint32_t sa = smem_u + stage * G::STAGE;
#pragma unroll
        for (int i = 0; i < SPS; ++i) {
            const int t = step * TPS + (C8 == 1 ? i : i / C8);
            const int c8 = C8 == 1 ? 0 : i % C8;
            const int zz = bz + p.off[cls][t][0], yy = by + p.off[cls][t][1], xx = bx + p.off[cls][t][2];
            const bool ok = live && (unsigned)zz < (unsigned)p.Di && (unsigned)yy < (unsigned)p.Hi && (unsigned)xx < (unsigned)p.Wi;
            const __half* src = ok ? in_b + ((size_t)c8 * Min + ((size_t)zz * p.Hi + yy) * p.Wi + xx) * 8 : in_b;
            tc::cp_async16_ca(sa + i * kSlabBytes + r * 16, src, ok);
        }
        const uint8_t* wsrc = wgt_c + (size_t)step * G::B_STAGE;
#pragma unroll
        for (uint32_t o = 0; o < G::B_STAGE; o += 128 * 16)
            if (o + r * 16 < G::B_STAGE) tc::cp_async16(sa + G::A_STAGE + o + r * 16, wsrc + o + r * 16, true);
    };

    // ---- prologue: first kStages-1 steps in flight ---------------------------------------------------------------
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if (s < nsteps) load_step(s, s);
        tc::cp_async_commit();
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tmem_u = tc::uniform(tmem);
    const bool elected = tc::elect_one();   // evaluated by every (converged) warp; only warp 0 issues

    constexpr uint32_t idesc = tc::instr_desc_f16(128, N);
#pragma unroll 1
    for (int step = 0; step < nsteps; ++step) {
        const int stage = step % kStages;
        tc::cp_async_wait<kStages - 2>();   // this thread's copies of `step` have landed
        tc::fence_proxy_async();            // ... and are visible to the tensor core's (async-proxy) reads
        __syncthreads();
        if (warp == 0) {
            tc::tc_fence_after();
            const uint32_t sa = smem_u + stage * G::STAGE;
#pragma unroll
            for (int j = 0; j < MPS; ++j) {
                const uint64_t da = tc::smem_desc(sa + j * 2 * kSlabBytes, kSlabBytes, 128);
                const uint64_t db = tc::smem_desc(sa + G::A_STAGE + j * G::B_MMA, N * 16, 128);
                if (elected) tc::mma_f16(tmem_u, da, db, idesc, step > 0 || j > 0);
            }
            if (elected) tc::mma_commit(bar_mma + stage);
            __syncwarp();
        }
        // refill the stage the PREVIOUS step used, once its MMAs have drained
        const int nxt = step + kStages - 1;
        if (nxt < nsteps) {
            if (step >= 1) tc::mbar_wait(bar_mma + (step - 1) % kStages, ((step - 1) / kStages) & 1);
            load_step(nxt, nxt % kStages);
        }
        tc::cp_async_commit();
    }
    // MMAs complete in order: the last step's commit covers everything
    tc::mbar_wait(bar_mma + (nsteps - 1) % kStages, ((nsteps - 1) / kStages) & 1);
    tc::tc_fence_after();

    // ---- epilogue ------------------------------------------------------------------------------------------------------
    size_t Mout, ovox;
    if (MODE == 2) {
        const int pd = cls >> 2, ph = (cls >> 1) & 1, pw = cls & 1;
        Mout = 8 * (size_t)p.Ro * p.Rh * p.Rw;
        ovox = ((size_t)(2 * rz + pd) * (2 * p.Rh) + 2 * ry + ph) * (2 * p.Rw) + 2 * rx + pw;
    } else {
        Mout = (size_t)p.Ro * p.Rh * p.Rw;
        ovox = ((size_t)rz * p.Rh + ry) * p.Rw + rx;
    }
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16);
#pragma unroll
    for (int c8 = 0; c8 < COUT / 8; ++c8) {
        uint32_t hi[8], lo[8];
        tc::tmem_ld8_nowait(taddr + c8 * 8, hi);            // warp-collective: every lane executes it
        tc::tmem_ld8_nowait(taddr + COUT + c8 * 8, lo);
        tc::tmem_ld_wait();
        if (live) {
            const size_t o = (((size_t)b * (COUT / 8) + c8) * Mout + ovox) * 8;
            float v[8], sk[8];
            if (MODE == 2 && p.skip) Vec8<__half>::load(p.skip + o, sk);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                float t = __uint_as_float(hi[i]) + __uint_as_float(lo[i]) + __ldg(p.bias + c8 * 8 + i);
                if (MODE == 2) {
                    t = fmaxf(t, 0.f);
                    v[i] = p.skip ? sk[i] + t : t;
                } else {
                    v[i] = p.relu ? fmaxf(t, 0.f) : t;
                }
            }
            Vec8<__half>::store(p.out + o, v);
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, G::TMEM_COLS);
}

template <int CIN, int COUT, int MODE>
int launch_gtc(GtcParams& p, int nclasses, cudaStream_t st) {
    using G = GtcCfg<CIN, COUT, MODE>;
    static_assert(G::SMEM <= 227 * 1024, "gather-conv stages do not fit in shared memory");
    auto kern = conv3d_gtc_kernel<CIN, COUT, MODE>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G::SMEM);
    if (e != cudaSuccess) { cds_set_error("cds_conv3d_gtc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    const long long M = (long long)p.B * p.Ro * p.Rh * p.Rw;
    dim3 grid(cds_div_up(M, 128), nclasses);
    kern<<<grid, 128, G::SMEM, st>>>(p);
    return cds_check_launch("cds_conv3d_gtc");
}

bool conv_pair_ok(int Cin, int Cout) {
    return (Cin == 8 && Cout == 16) || (Cin == 16 && Cout == 16) || (Cin == 16 && Cout == 32) || (Cin == 32 && Cout == 32) ||
           (Cin == 32 && Cout == 64) || (Cin == 64 && Cout == 64);
}
bool deconv_pair_ok(int Cin, int Cout) { return (Cin == 64 && Cout == 32) || (Cin == 32 && Cout == 16); }

}  // namespace

extern "C" {

int cds_conv3d_k3_gtc_supported(int Cin, int Cout, int stride) { return (stride == 1 || stride == 2) && conv_pair_ok(Cin, Cout); }

// fp16 elements of the packed weight image: (28 or 27*Cin/8 slabs) / 2 MMAs of 2*Cout columns x 16 k
int cds_conv3d_k3_gtc_weight_halfs(int Cin, int Cout) {
    int nslab = Cin == 8 ? 28 : 27 * (Cin / 8);
    return (nslab / 2) * (2 * Cout) * 16;
}

int cds_conv3d_k3_gtc(const void* in, const void* wgt_packed, const float* bias, int B, int Cin, int Cout, int D, int H, int W,
                      int stride, int relu, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt_packed && bias && out, CDS_EARG, "cds_conv3d_k3_gtc: null pointer");
    CDS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, CDS_ESHAPE, "cds_conv3d_k3_gtc: bad shape B=%d D=%d H=%d W=%d", B, D, H, W);
    CDS_REQUIRE(cds_conv3d_k3_gtc_supported(Cin, Cout, stride), CDS_EUNSUPPORTED,
                "cds_conv3d_k3_gtc: unsupported layer Cin=%d Cout=%d stride=%d", Cin, Cout, stride);
    GtcParams p = {};
    p.in = (const __half*)in; p.wgt = (const __half*)wgt_packed; p.bias = bias; p.skip = nullptr; p.out = (__half*)out;
    p.B = B; p.Di = D; p.Hi = H; p.Wi = W;
    p.Ro = (D + stride - 1) / stride; p.Rh = (H + stride - 1) / stride; p.Rw = (W + stride - 1) / stride;
    p.relu = relu;
    // tap table: Cin = 8 carries a zero-weight pad slab after tap 0 (it re-gathers tap 0: finite data x zero weights)
    int n = 0;
    for (int t = 0; t < 27; ++t) {
        p.off[0][n][0] = (signed char)(t / 9 - 1); p.off[0][n][1] = (signed char)((t / 3) % 3 - 1); p.off[0][n][2] = (signed char)(t % 3 - 1);
        ++n;
        if (Cin == 8 && t == 0) { p.off[0][n][0] = -1; p.off[0][n][1] = -1; p.off[0][n][2] = -1; ++n; }
    }
    p.ntaps[0] = n; p.wofs[0] = 0;
#define CDS_GTC(ci, co)                                                                   \
    if (Cin == ci && Cout == co)                                                          \
        return stride == 1 ? launch_gtc<ci, co, 0>(p, 1, stream) : launch_gtc<ci, co, 1>(p, 1, stream);
    CDS_GTC(8, 16) CDS_GTC(16, 16) CDS_GTC(16, 32) CDS_GTC(32, 32) CDS_GTC(32, 64) CDS_GTC(64, 64)
#undef CDS_GTC
    return CDS_EUNSUPPORTED;
}

int cds_deconv3d_k3s2_gtc_supported(int Cin, int Cout) { return deconv_pair_ok(Cin, Cout); }

int cds_deconv3d_k3s2_gtc_weight_halfs(int Cin, int Cout) { return (27 * (Cin / 8) / 2) * (2 * Cout) * 16; }

// in [B, Cin/8, D, H, W, 8] -> out [B, Cout/8, 2D, 2H, 2W, 8] (+ skip of the output shape)
int cds_deconv3d_k3s2_gtc(const void* in, const void* wgt_packed, const float* bias, const void* skip, int B, int Cin, int Cout,
                          int D, int H, int W, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt_packed && bias && out, CDS_EARG, "cds_deconv3d_k3s2_gtc: null pointer");
    CDS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, CDS_ESHAPE, "cds_deconv3d_k3s2_gtc: bad shape");
    CDS_REQUIRE(deconv_pair_ok(Cin, Cout), CDS_EUNSUPPORTED, "cds_deconv3d_k3s2_gtc: unsupported layer Cin=%d Cout=%d", Cin, Cout);
    GtcParams p = {};
    p.in = (const __half*)in; p.wgt = (const __half*)wgt_packed; p.bias = bias; p.skip = (const __half*)skip; p.out = (__half*)out;
    p.B = B; p.Di = D; p.Hi = H; p.Wi = W; p.Ro = D; p.Rh = H; p.Rw = W; p.relu = 1;
    int mma = 0;
    for (int cls = 0; cls < 8; ++cls) {
        const int pd = cls >> 2, ph = (cls >> 1) & 1, pw = cls & 1;
        int n = 0;
        for (int sd = 0; sd <= pd; ++sd)
            for (int sh = 0; sh <= ph; ++sh)
                for (int sw = 0; sw <= pw; ++sw) {
                    p.off[cls][n][0] = (signed char)sd; p.off[cls][n][1] = (signed char)sh; p.off[cls][n][2] = (signed char)sw;
                    ++n;
                }
        p.ntaps[cls] = n;
        p.wofs[cls] = mma;
        mma += n * (Cin / 8) / 2;
    }
    if (Cin == 64) return launch_gtc<64, 32, 2>(p, 8, stream);
    return launch_gtc<32, 16, 2>(p, 8, stream);
}

}  // extern "C"
