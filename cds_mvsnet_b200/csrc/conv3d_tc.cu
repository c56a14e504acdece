// 3-D regulariser convolutions on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM).
//
// Conv3d k3 p1 stride 1 (+ folded BatchNorm bias, ReLU) as an implicit GEMM with NO im2col:
//   M = 128 consecutive voxels along w (one "row unit", 126 of them valid outputs), N = Cout (padded to 16),
//   K = 27 taps x Cin.
// Activations are channel-blocked [C/8][D][H][W][8 x fp16], so an 8-channel slab of a voxel row is contiguous.
// A CTA owns TY row units of one depth slice.  Its haloed input window [3 d][TY+2 h][128 w] is staged once
// in shared memory by TMA (one cp.async.bulk.tensor per 8-channel slab, 2 KB rows; out-of-volume coordinates
// are zero filled by the hardware = the conv padding) as [slab][d][h][w][8 x fp16], 16 B per voxel, which is the tcgen05
// K-major no-swizzle operand layout, so the A operand of tap (kd,kh,kw) is the same tile read at the start
// address ((kd,kh) row, kw column): 14..54 MMAs (K=16 = two 8-channel slabs each) per row unit, issued by
// one thread, accumulate in TMEM; then 4 warps pull their 32 lanes with tcgen05.ld, add bias, ReLU, convert
// to fp16 and write 16 B per 8 channels (coalesced NDHWC).
//
// Reference semantics: models/module.py:80-122 (Conv3d block), wired at :305-309.
// Weights: fp16, pre-laid-out by the host (weights.py: pack_conv3d_tc) in exactly the smem image order:
//   [mma j][k-chunk 2][n-group NPAD/8][8 rows n][8 halfs k]   (K-major, no swizzle; LBO = NPAD*16, SBO = 128).
// For Cout = 8 the 8 padding columns of N = 16 carry the fp16 rounding residual of the weights (summed in the
// epilogue), so those layers see effectively fp32-accurate weights at no extra tensor-core cost.
#include "cds_common.cuh"
#include "tc_common.cuh"
#include "tma_host.h"

namespace {

constexpr int TX = 128;               // voxels per staged row = MMA M (TMA box limit: 256 x 8-byte elements)
constexpr int TXO = TX - 2;           // valid outputs per row unit (the other two MMA rows run into the halo)
constexpr int ROW_BYTES = TX * 16;

template <int CIN>
struct KOrder {
    static constexpr int C8 = CIN / 8;
    static constexpr int NSLAB = 27 * C8;
    static constexpr int NMMA = (NSLAB + 1) / 2;
};

// bytes of one 8-channel slab of the window (a multiple of 128: every slab is a legal TMA destination)
template <int TY>
__host__ __device__ constexpr uint32_t chunk_bytes() { return (uint32_t)(3 * (TY + 2) * ROW_BYTES); }

// byte offset of slab (tap, c8) inside the A tile, relative to the row unit's origin
template <int TY>
__host__ __device__ constexpr uint32_t slab_off(int tap, int c8) {
    int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
    return (uint32_t)(c8 * chunk_bytes<TY>() + (kd * (TY + 2) + kh) * ROW_BYTES + kw * 16);
}

struct ConvTcParams {
    const __half* wgt;   // packed fp16 image, NMMA * 2 * NPAD * 8 halfs
    const float* bias;   // [COUT]
    __half* out;         // [COUT/8, D, H, W, 8]  (one batch item)
    float* logits;       // prob-head mode (COUT == 1): fp32 [D, H, W], no bias / ReLU
    int D, H, W, relu;
};

// tmap: 4-D view (2W x 8-byte elements, H, D, C/8) of one batch item's channel-blocked input, box (256, TY+2, 3, 1)
template <int CIN, int COUT, int NPAD, int TY>
__global__ void __launch_bounds__(128) conv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap, ConvTcParams p) {
    using KO = KOrder<CIN>;
    constexpr int C8 = KO::C8, NMMA = KO::NMMA;
    constexpr uint32_t CHUNK = chunk_bytes<TY>();
    constexpr uint32_t A_BYTES = C8 * CHUNK;
    constexpr uint32_t A_TX = A_BYTES;                            // bytes the TMA delivers
    constexpr uint32_t B_BYTES = NMMA * 2 * NPAD * 16;
    constexpr uint32_t TMEM_COLS = (TY * NPAD <= 32) ? 32 : (TY * NPAD <= 64 ? 64 : (TY * NPAD <= 128 ? 128 : 256));
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
    uint64_t* bar_mma = bar_load + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = max(0, min((int)blockIdx.x * TXO, p.W - TXO));   // last tile overlaps its neighbour; W < TXO: one partial tile
    const int y0 = blockIdx.y * TY;
    const int d = blockIdx.z;
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);

    if (warp == 0) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 32) {
        tc::mbar_init(bar_load, 1);
        tc::mbar_init(bar_mma, TY < 4 ? TY : 4);   // one commit per issuing warp
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    // ---- one thread launches the TMA of the haloed window + weights -----------------------------------------
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(bar_load, A_TX + B_BYTES);
#pragma unroll
        for (int c8 = 0; c8 < C8; ++c8) tc::tma_load_4d(sA_u + c8 * CHUNK, &tmap, bar_load, 2 * (x0 - 1), y0 - 1, d - 1, c8);
        tc::bulk_copy_g2s(sB_u, p.wgt, B_BYTES, bar_load);
    }
    // ---- MMA issue, spread over the warps: lane 0 of warp w issues the MMAs of row units w, w+4, ... ------------
    // (the issue rate of ONE thread, ~50 cycles per descriptor+MMA, would otherwise bound an 8-cycle N=16 MMA)
    const uint32_t warp_u = tc::uniform((uint32_t)warp);
    const uint32_t tmem_u = tc::uniform(tmem);
    if (warp_u < (uint32_t)TY) {
        tc::mbar_wait(bar_load, 0);
        tc::tc_fence_after();
        const bool elected = tc::elect_one();
        constexpr uint32_t idesc = tc::instr_desc_f16(128, NPAD);
        // Descriptors: hi word is constant (SBO = 128 B, version 1); lo word = (start >> 4) | (LBO >> 4) << 16.
        // The j loop is fully unrolled so every slab offset / LBO is a compile-time constant.
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
        constexpr uint32_t b_lo_const = ((uint32_t)(NPAD * 16) >> 4) << 16;
#pragma unroll 1
        for (uint32_t u = warp_u; u < (uint32_t)TY; u += 4) {
            const uint32_t a_base = (sA_u + u * ROW_BYTES) >> 4;
            const uint32_t b_base = sB_u >> 4;
            const uint32_t acc_col = tmem_u + u * NPAD;
#pragma unroll
            for (int j = 0; j < NMMA; ++j) {
                uint32_t off0, lbo;
                if (C8 == 1) {
                    // 27 slabs: [tap0, pad], [tap1, tap2], ..., [tap25, tap26]; the pad slab has zero weights
                    if (j == 0) { off0 = 0; lbo = 16; }
                    else { off0 = slab_off<TY>(2 * j - 1, 0); lbo = slab_off<TY>(2 * j, 0) - off0; }
                } else {
                    const int s0 = 2 * j;     // slabs ordered tap-major, channel-chunk minor; a pair shares the tap
                    off0 = slab_off<TY>(s0 / C8, s0 % C8);
                    lbo = CHUNK;
                }
                const uint32_t a_lo = a_base + ((off0 >> 4) | ((lbo >> 4) << 16));
                const uint32_t b_lo = b_base + (((uint32_t)j * (2 * NPAD * 16)) >> 4 | b_lo_const);
                const uint64_t da = ((uint64_t)desc_hi << 32) | a_lo;
                const uint64_t db = ((uint64_t)desc_hi << 32) | b_lo;
                if (elected) tc::mma_f16(acc_col, da, db, idesc, j > 0);
            }
        }
        if (elected) tc::mma_commit(bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(bar_mma, 0);
    tc::tc_fence_after();

    // ---- epilogue: TMEM -> registers -> bias/ReLU -> fp16, channel-blocked ------------------------------------
    const int r = warp * 32 + lane;          // MMA row = voxel x0 + r; rows 126, 127 are halo garbage
    const size_t M = (size_t)p.D * p.H * p.W;
#pragma unroll 1
    for (int u = 0; u < TY; ++u) {
        const int y = y0 + u;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)u * NPAD;
        if constexpr (COUT == 1) {
            // prob head (models/module.py:303): column 0 = product with the fp16-rounded weights, column 1 = with
            // their rounding residual; plain sum, fp32 logits in [D,H,W] order
            float v[8];
            tc::tmem_ld8(taddr, v);
            if (y < p.H && r < TXO && x0 + r < p.W) p.logits[((size_t)d * p.H + y) * p.W + x0 + r] = v[0] + v[1];
        } else {
#pragma unroll
            for (int c8 = 0; c8 < COUT / 8; ++c8) {
                float v[8];
                tc::tmem_ld8(taddr + c8 * 8, v);   // warp-collective: every lane executes it
                if (COUT == 8) {
                    // N was padded 8 -> 16: columns 8..15 hold the product with the weights' fp16 rounding residual
                    float lo[8];
                    tc::tmem_ld8(taddr + 8, lo);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[i] += lo[i];
                }
                if (y < p.H && r < TXO && x0 + r < p.W) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        float t = v[i] + __ldg(p.bias + c8 * 8 + i);
                        v[i] = p.relu ? fmaxf(t, 0.f) : t;
                    }
                    Vec8<__half>::store(p.out + ((size_t)c8 * M + ((size_t)d * p.H + y) * p.W + x0 + r) * 8, v);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------------------------
// ConvTranspose3d k3 s2 p1 op1 + BN + ReLU, then + skip (models/module.py:125-166, 310-312) on the tensor cores.
// out[2i - 1 + k] += in[i] * w[k] per axis: output parity 0 takes tap k=1 from input i, parity 1 takes k=2 from i and
// k=0 from i+1.  So with M = 128 consecutive INPUT voxels as the GEMM rows, the A operands are the 8 neighbour
// offsets (sd,sh,sw) in {0,1}^3 of the same haloed window, and the 8 output parity classes sit side by side in N
// (N = 8*Cout, zero weight blocks where a parity does not use an offset): 8 * Cin/16 MMAs per row unit.
// The epilogue scatters each row's 8 parity results to the 2x2x2 output cell (bias, ReLU, + skip).
// Weights: [mma (offset s, chunk pair q)][k-chunk 2][N/8][8 n][8 k], n = parity*Cout + co.
// ---------------------------------------------------------------------------------------------------------------
struct DeconvTcParams {
    const __half* wgt;
    const float* bias;    // [COUT]
    const __half* skip;   // [COUT/8, 2D, 2H, 2W, 8] or null (one batch item)
    __half* out;          // [COUT/8, 2D, 2H, 2W, 8]
    int D, H, W;          // INPUT extent
};

template <int CIN, int COUT, int TY>
__global__ void __launch_bounds__(128) deconv3d_tc_kernel(const __grid_constant__ CUtensorMap tmap, DeconvTcParams p) {
    constexpr int C8 = CIN / 8, NMMA = 8 * C8 / 2, N = 8 * COUT;
    constexpr int TXI = TX - 1;                                       // input voxels per row unit that own outputs
    constexpr uint32_t CHUNK = 2 * (TY + 1) * ROW_BYTES;              // one 8-channel slab: [2 d][TY+1 h][128 w]
    constexpr uint32_t A_BYTES = C8 * CHUNK;
    constexpr uint32_t B_BYTES = NMMA * 2 * N * 16;
    constexpr uint32_t TMEM_COLS = TY * N <= 64 ? 64 : (TY * N <= 128 ? 128 : (TY * N <= 256 ? 256 : 512));
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES;
    uint64_t* bar_load = reinterpret_cast<uint64_t*>(smem + A_BYTES + B_BYTES);
    uint64_t* bar_mma = bar_load + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_load + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int x0 = max(0, min((int)blockIdx.x * TXI, p.W - TXI));
    const int y0 = blockIdx.y * TY;
    const int d = blockIdx.z;
    const uint32_t sA_u = tc::smem_u32(sA), sB_u = tc::smem_u32(sB);

    if (warp == 0) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 32) {
        tc::mbar_init(bar_load, 1);
        tc::mbar_init(bar_mma, TY < 4 ? TY : 4);
        tc::mbar_fence_init();
        tc::tma_prefetch_desc(&tmap);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(bar_load, A_BYTES + B_BYTES);
#pragma unroll
        for (int c8 = 0; c8 < C8; ++c8) tc::tma_load_4d(sA_u + c8 * CHUNK, &tmap, bar_load, 2 * x0, y0, d, c8);
        tc::bulk_copy_g2s(sB_u, p.wgt, B_BYTES, bar_load);
    }
    const uint32_t warp_u = tc::uniform((uint32_t)warp);
    const uint32_t tmem_u = tc::uniform(tmem);
    if (warp_u < (uint32_t)TY) {
        tc::mbar_wait(bar_load, 0);
        tc::tc_fence_after();
        const bool elected = tc::elect_one();
        constexpr uint32_t idesc = tc::instr_desc_f16(128, N);
        constexpr uint32_t desc_hi = (128u >> 4) | (1u << 14);
        constexpr uint32_t b_lo_const = ((uint32_t)(N * 16) >> 4) << 16;
#pragma unroll 1
        for (uint32_t u = warp_u; u < (uint32_t)TY; u += 4) {
            const uint32_t a_base = (sA_u + u * ROW_BYTES) >> 4;
            const uint32_t b_base = sB_u >> 4;
            const uint32_t acc_col = tmem_u + u * N;
#pragma unroll
            for (int j = 0; j < NMMA; ++j) {
                const int s = j / (C8 / 2), q = j % (C8 / 2);   // neighbour offset (sd,sh,sw), channel-chunk pair
                const int sd = s >> 2, sh = (s >> 1) & 1, sw = s & 1;
                const uint32_t off0 = (uint32_t)(2 * q) * CHUNK + (uint32_t)((sd * (TY + 1) + sh) * ROW_BYTES + sw * 16);
                const uint32_t a_lo = a_base + ((off0 >> 4) | ((CHUNK >> 4) << 16));
                const uint32_t b_lo = b_base + (((uint32_t)j * (2 * N * 16)) >> 4 | b_lo_const);
                if (elected) tc::mma_f16(acc_col, ((uint64_t)desc_hi << 32) | a_lo, ((uint64_t)desc_hi << 32) | b_lo, idesc, j > 0);
            }
        }
        if (elected) tc::mma_commit(bar_mma);
    }
    __syncwarp();
    tc::mbar_wait(bar_mma, 0);
    tc::tc_fence_after();

    // ---- epilogue: row r = input voxel x0 + r owns the 2x2x2 output cell at (2d, 2y, 2x) -------------------------
    const int r = warp * 32 + lane;
    const int xi = x0 + r;
    const int Do = 2 * p.D, Ho = 2 * p.H, Wo = 2 * p.W;
    const size_t Mo = (size_t)Do * Ho * Wo;
    const bool own = r < TXI && xi < p.W && xi >= (int)blockIdx.x * TXI;   // tiles overlap at the right edge: one owner per voxel
    float bias[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) bias[c] = __ldg(p.bias + c);
#pragma unroll 1
    for (int u = 0; u < TY; ++u) {
        const int yi = y0 + u;
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)u * N;
        const bool live = own && yi < p.H;
        const size_t o000 = ((size_t)(2 * d) * Ho + 2 * yi) * Wo + 2 * xi;   // output voxel of parity (0,0,0)
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            // the whole 2x2x2 cell of skip values is fetched up front: the loads do not depend on the accumulators, and a
            // DRAM round trip per parity class in the store loop's critical path is what used to bound this kernel
            uint4 sk[8];
#pragma unroll
            for (int par = 0; par < 8; ++par) {
                const size_t o = ((size_t)c8 * Mo + o000 + ((size_t)(par >> 2) * Ho + ((par >> 1) & 1)) * Wo + (par & 1)) * 8;
                sk[par] = (live && p.skip) ? __ldcs(reinterpret_cast<const uint4*>(p.skip + o)) : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
            for (int par = 0; par < 8; ++par) {
                float v[8];
                tc::tmem_ld8(taddr + par * COUT + c8 * 8, v);
                if (live) {
                    const size_t o = ((size_t)c8 * Mo + o000 + ((size_t)(par >> 2) * Ho + ((par >> 1) & 1)) * Wo + (par & 1)) * 8;
                    const __half2* sh = reinterpret_cast<const __half2*>(&sk[par]);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 s2 = __half22float2(sh[i]);
                        v[2 * i] = s2.x + fmaxf(v[2 * i] + bias[c8 * 8 + 2 * i], 0.f);
                        v[2 * i + 1] = s2.y + fmaxf(v[2 * i + 1] + bias[c8 * 8 + 2 * i + 1], 0.f);
                    }
                    Vec8<__half>::store(p.out + o, v);
                }
            }
        }
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
}

template <int CIN, int COUT, int TY>
int launch_deconv_tc(const void* in, const void* wgt, const float* bias, const void* skip, int B, int D, int H, int W, void* out,
                     cudaStream_t st) {
    constexpr int C8 = CIN / 8;
    constexpr size_t smem = (size_t)C8 * 2 * (TY + 1) * ROW_BYTES + (size_t)(8 * C8 / 2) * 2 * (8 * COUT) * 16 + 32;
    static_assert(smem <= 227 * 1024, "deconv tile does not fit in shared memory");
    auto kern = deconv3d_tc_kernel<CIN, COUT, TY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cds_set_error("cds_deconv3d_k3s2_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(cds_div_up(W, TX - 1), cds_div_up(H, TY), D);
    for (int b = 0; b < B; ++b) {
        const __half* base = (const __half*)in + (size_t)b * D * H * W * CIN;
        CUtensorMap tmap;
        const uint64_t dims[4] = {2 * (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)C8};
        const uint64_t strides[4] = {0, (uint64_t)W * 16, (uint64_t)H * W * 16, (uint64_t)D * H * W * 16};
        const uint32_t box[4] = {2 * TX, TY + 1, 2, 1};
        if (!tma::make_u64(&tmap, base, 4, dims, strides, box)) return CDS_EUNSUPPORTED;
        DeconvTcParams p;
        p.wgt = (const __half*)wgt;
        p.bias = bias;
        p.skip = skip ? (const __half*)skip + (size_t)b * 8 * D * H * W * COUT : nullptr;
        p.out = (__half*)out + (size_t)b * 8 * D * H * W * COUT;
        p.D = D; p.H = H; p.W = W;
        kern<<<grid, 128, smem, st>>>(tmap, p);
    }
    return cds_check_launch("cds_deconv3d_k3s2_tc");
}

template <int CIN, int COUT, int NPAD, int TY>
int launch_tc(const void* in, const void* wgt, const float* bias, int B, int D, int H, int W, int relu, void* out,
              cudaStream_t st) {
    using KO = KOrder<CIN>;
    constexpr size_t smem = (size_t)KO::C8 * chunk_bytes<TY>() + (size_t)KO::NMMA * 2 * NPAD * 16 + 32;
    auto kern = conv3d_tc_kernel<CIN, COUT, NPAD, TY>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cds_set_error("cds_conv3d_k3_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(cds_div_up(W, TXO), cds_div_up(H, TY), D);
    for (int b = 0; b < B; ++b) {
        const __half* base = (const __half*)in + (size_t)b * D * H * W * CIN;
        CUtensorMap tmap;
        const uint64_t dims[4] = {2 * (uint64_t)W, (uint64_t)H, (uint64_t)D, (uint64_t)KO::C8};
        const uint64_t strides[4] = {0, (uint64_t)W * 16, (uint64_t)H * W * 16, (uint64_t)D * H * W * 16};
        const uint32_t box[4] = {2 * TX, TY + 2, 3, 1};
        if (!tma::make_u64(&tmap, base, 4, dims, strides, box)) return CDS_EUNSUPPORTED;
        ConvTcParams p;
        p.out = COUT == 1 ? nullptr : (__half*)out + (size_t)b * D * H * W * COUT;
        p.logits = COUT == 1 ? (float*)out + (size_t)b * D * H * W : nullptr;
        p.wgt = (const __half*)wgt;
        p.bias = bias;
        p.D = D; p.H = H; p.W = W; p.relu = relu;
        kern<<<grid, 128, smem, st>>>(tmap, p);
    }
    return cds_check_launch("cds_conv3d_k3_tc");
}

}  // namespace

extern "C" {

// 1 when the tensor-core kernel covers this layer shape (otherwise use cds_conv3d_k3)
int cds_conv3d_k3_tc_supported(int Cin, int Cout, int D, int H, int W, int stride) {
    if (stride != 1 || W < 8 || D < 1 || H < 1 || D > 65535) return 0;
    return (Cin == 8 && Cout == 1) || (Cin == 8 && Cout == 8) || (Cin == 16 && Cout == 8) || (Cin == 32 && Cout == 8) || (Cin == 16 && Cout == 16) ||
           (Cin == 32 && Cout == 32);
}

// number of fp16 elements of the packed weight image for (Cin, Cout)
int cds_conv3d_k3_tc_weight_halfs(int Cin, int Cout) {
    int npad = Cout < 16 ? 16 : Cout;
    int nmma = (27 * (Cin / 8) + 1) / 2;
    return nmma * 2 * npad * 8;
}

int cds_conv3d_k3_tc(const void* in, const void* wgt_packed, const float* bias, int B, int Cin, int Cout, int D, int H,
                     int W, int relu, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt_packed && out && (bias || Cout == 1), CDS_EARG, "cds_conv3d_k3_tc: null pointer");
    if (Cin == 8 && Cout == 1) {   // prob head: fp32 logits [B,D,H,W]
        CDS_REQUIRE(cds_conv3d_k3_tc_supported(8, 1, D, H, W, 1), CDS_EUNSUPPORTED, "cds_conv3d_k3_tc: prob head needs W >= 8");
        return launch_tc<8, 1, 16, 8>(in, wgt_packed, bias, B, D, H, W, 0, out, stream);
    }
    CDS_REQUIRE(cds_conv3d_k3_tc_supported(Cin, Cout, D, H, W, 1), CDS_EUNSUPPORTED,
                "cds_conv3d_k3_tc: unsupported shape Cin=%d Cout=%d D=%d H=%d W=%d (needs W >= 8)", Cin, Cout, D, H, W);
    if (Cin == 8 && Cout == 8) return launch_tc<8, 8, 16, 8>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
    if (Cin == 16 && Cout == 8) return launch_tc<16, 8, 16, 4>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
    if (Cin == 32 && Cout == 8) return launch_tc<32, 8, 16, 2>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
    if (Cin == 16 && Cout == 16) return launch_tc<16, 16, 16, 4>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
    return launch_tc<32, 32, 32, 2>(in, wgt_packed, bias, B, D, H, W, relu, out, stream);
}

// ---- transposed convolution on the tensor cores (input extent D,H,W; output 2D,2H,2W) ---------------------------------
int cds_deconv3d_k3s2_tc_supported(int Cin, int Cout, int D, int H, int W) {
    if (W < 8 || D < 1 || H < 1 || D > 65535) return 0;
    return (Cin == 16 && Cout == 8) || (Cin == 32 && Cout == 16);   // 64 -> 32: the weight image (256 KB) exceeds smem
}

int cds_deconv3d_k3s2_tc_weight_halfs(int Cin, int Cout) { return (8 * (Cin / 8) / 2) * 2 * (8 * Cout) * 8; }

int cds_deconv3d_k3s2_tc(const void* in, const void* wgt_packed, const float* bias, const void* skip, int B, int Cin, int Cout,
                         int D, int H, int W, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt_packed && bias && out, CDS_EARG, "cds_deconv3d_k3s2_tc: null pointer");
    CDS_REQUIRE(cds_deconv3d_k3s2_tc_supported(Cin, Cout, D, H, W), CDS_EUNSUPPORTED,
                "cds_deconv3d_k3s2_tc: unsupported shape Cin=%d Cout=%d D=%d H=%d W=%d (needs input W >= 8)", Cin, Cout, D, H, W);
    if (Cin == 16) return launch_deconv_tc<16, 8, 4>(in, wgt_packed, bias, skip, B, D, H, W, out, stream);
    return launch_deconv_tc<32, 16, 2>(in, wgt_packed, bias, skip, B, D, H, W, out, stream);
}

}  // extern "C"
