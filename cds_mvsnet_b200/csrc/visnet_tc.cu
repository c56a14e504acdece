// Visibility net (A3) on the tensor cores: three 3x3 conv layers chained through shared memory as tcgen05 tap GEMMs.
//
// Reference: models/model.py:14,51 ; ConvBnReLU models/module.py:169-198 (BatchNorm folded by the host).
//   cat(entropy, |curv|) -> conv3x3 2->16 +ReLU -> conv3x3 16->16 +ReLU -> conv3x3 16->16 +ReLU -> conv1x1 16->1 -> sigmoid
// A CTA owns TY output rows x 122 columns.  Every buffer is a set of 128-pixel rows of 8-channel fp16 slabs (the
// tcgen05 K-major no-swizzle operand layout, see tc_common.cuh) sharing one x origin xs = x0 - 3:
//   in8  [TY+6 rows]      (entropy_hi, curv_hi, entropy_lo, curv_lo, 0,0,0,0): fp32 inputs split into two fp16 terms,
//                         the layer-1 weights are duplicated for the "lo" channels, so the inputs are exact
//   a1   [2][TY+4 rows]   layer-1 output (16 ch = 2 slabs), written by the epilogue at pixel j+1
//   a2   [2][TY+2 rows]   layer-2 output, re-using the in8 region
// Each layer: M = 128 pixels of a row (output pixel = j+1), the kernel ROWS folded into N (as csrc/dynconv_kh.cu): ONE MMA on
// input row i and horizontal tap step s produces, in N = 48 columns, its contributions to the three output rows i-2, i-1, i
// (column group g <-> output row i-2+g, kernel row 2-g; ranges clipped at the buffer's first / last rows).  The accumulators of
// a layer's output rows sit side by side in TMEM (slot = output row, 16 columns each), every MMA accumulates and the slots are
// handed over ZEROED (tcgen05.st in the epilogue that drained them).  An N = 16 MMA costs 36 cycles of operand fetch for 8 of
// math; N = 48 costs 44 for three times the work: 62 MMAs per tile instead of 130.  One warp issues a layer's MMAs in program
// order (fixed accumulation order).  TMEM accumulators are read back by the 8 warps, bias+ReLU, zeroed outside the image
// (each layer zero-pads its own input), packed to fp16 and stored as the next layer's operand; layer 3 finishes with the 1x1
// conv + sigmoid and writes fp32.
#include <algorithm>

#include "cds_common.cuh"
#include "tc_common.cuh"

namespace {

#ifndef CDS_VIS_TY
#define CDS_VIS_TY 4   // 8-row tiles (2 resident CTAs, fewer halo MMAs) measured the same (0.381 vs 0.376 ms at stage 3): the tile is not MMA-bound
#endif
constexpr int TX = 128, TXO = TX - 6, TY = CDS_VIS_TY;
constexpr int ROW_BYTES = TX * 16;
constexpr int R_IN = TY + 6, R_A1 = TY + 4, R_A2 = TY + 2;
constexpr uint32_t X_BYTES = 2 * R_A2 * ROW_BYTES;       // in8 (R_IN rows, 1 slab) then a2 (2 slabs x R_A2 rows)
constexpr uint32_t Y_BYTES = 2 * R_A1 * ROW_BYTES;       // a1
constexpr int MMA_L1 = 2, MMA_L23 = 3;                    // horizontal tap steps per input row: (kw0, kw1), (-, kw2); kw0, kw1, kw2
constexpr uint32_t W_STEP = 2 * 48 * 16;                  // one step's operand image: [k-chunk 2][48 columns][8 k] fp16
constexpr uint32_t W_BYTES = (MMA_L1 + 2 * MMA_L23) * W_STEP;
static_assert(X_BYTES >= R_IN * ROW_BYTES, "in8 must fit in the region a2 re-uses");
constexpr uint32_t TMEM_COLS = R_A1 * 16 <= 128 ? 128 : 256;   // R_A1 row units x 16 accumulator columns
static_assert(R_A1 * 16 <= 256, "row units of layer 1 exceed the TMEM budget of two resident CTAs");

struct VisTcParams {
    const float* entropy;   // [n][H][W]
    const float* curv;      // [n][H][W]
    const __half* wgt;      // packed fp16: L1 [2], L2 [3], L3 [3] steps x [2 k-chunk][6][8 n][8 k]
    const float* fparams;   // b1[16] b2[16] b3[16] w4[16] b4[1]
    float* vis;             // [n][H][W]
    int H, W;
};

constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);

// One MMA of the row fold: input row I of a buffer of RI rows (outputs RI - 2 rows), tap step S.
//   layer 1 (L1): one 8-channel slab per pixel; a step pairs two neighbouring pixels (LBO 16 B): step 0 = taps (0, 1), step 1 =
//                 taps (1, 2) with zero weights for tap 1 (a pad chunk BEHIND tap 2 would read past the staged rows, and
//                 0 x garbage may be NaN)
//   layers 2/3:   two slabs (channel chunks, CHUNK bytes apart) per pixel; step S = tap S
template <int RI, int NSTEP, bool L1, uint32_t CHUNK, int WOFF, int K>
__device__ __forceinline__ void mma_fold(uint32_t a_base, uint32_t b_base, uint32_t acc, bool elected) {
    constexpr int I = K / NSTEP, S = K % NSTEP, RO = RI - 2;
    constexpr int olo = I - 2 > 0 ? I - 2 : 0, ohi = I < RO - 1 ? I : RO - 1;
    constexpr int g0 = olo - (I - 2), ng = ohi - olo + 1;
    static_assert(ng >= 1 && ng <= 3 && g0 >= 0 && g0 + ng <= 3, "row range of the fold");
    constexpr uint32_t a_off = (uint32_t)(I * ROW_BYTES + S * 16);
    constexpr uint32_t lbo = L1 ? 16u : CHUNK;
    constexpr uint32_t a_const = (a_off >> 4) | ((lbo >> 4) << 16);
    constexpr uint32_t b_const = (((uint32_t)(WOFF + S) * W_STEP + (uint32_t)g0 * 256) >> 4) | ((768u >> 4) << 16);
    if (elected)
        tc::mma_f16(acc + (uint32_t)olo * 16, ((uint64_t)kDescHi << 32) | (a_base + a_const), ((uint64_t)kDescHi << 32) | (b_base + b_const),
                    tc::instr_desc_f16(128, ng * 16), true);
}
template <int RI, int NSTEP, bool L1, uint32_t CHUNK, int WOFF, int... K>
__device__ __forceinline__ void issue_fold(uint32_t a, uint32_t b, uint32_t acc, bool e, std::integer_sequence<int, K...>) {
    (mma_fold<RI, NSTEP, L1, CHUNK, WOFF, K>(a, b, acc, e), ...);
}
// all MMAs of a layer, input rows in order
template <int RI, int NSTEP, bool L1, uint32_t CHUNK, int WOFF>
__device__ __forceinline__ void issue_layer(uint32_t a, uint32_t b, uint32_t acc, bool e) {
    issue_fold<RI, NSTEP, L1, CHUNK, WOFF>(a, b, acc, e, std::make_integer_sequence<int, RI * NSTEP>{});
}

__device__ __forceinline__ void tmem_zero16(uint32_t taddr) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1,%1};\n" ::"r"(taddr), "r"(0u) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }

// bias + ReLU of one row unit's 16 accumulator columns; result as two packed fp16 slabs (or zero outside the image)
__device__ __forceinline__ void act16(uint32_t taddr, const float* __restrict__ bias, bool inside, bool zero_after, uint4& lo, uint4& hi) {
    uint32_t r0[8], r1[8];
    tc::tmem_ld8_nowait(taddr, r0);
    tc::tmem_ld8_nowait(taddr + 8, r1);
    tc::tmem_ld_wait();
    if (zero_after) tmem_zero16(taddr);   // the next layer accumulates into this slot
    __half2* l = reinterpret_cast<__half2*>(&lo);
    __half2* h = reinterpret_cast<__half2*>(&hi);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float a = inside ? fmaxf(__uint_as_float(r0[2 * i]) + bias[2 * i], 0.f) : 0.f;
        float b = inside ? fmaxf(__uint_as_float(r0[2 * i + 1]) + bias[2 * i + 1], 0.f) : 0.f;
        float c = inside ? fmaxf(__uint_as_float(r1[2 * i]) + bias[8 + 2 * i], 0.f) : 0.f;
        float d = inside ? fmaxf(__uint_as_float(r1[2 * i + 1]) + bias[8 + 2 * i + 1], 0.f) : 0.f;
        l[i] = __floats2half2_rn(a, b);
        h[i] = __floats2half2_rn(c, d);
    }
}

constexpr int NT = 256;   // 8 warps: two per TMEM lane quadrant share each layer's epilogue rows; warp 0 issues the MMAs

__global__ void __launch_bounds__(NT) visnet_tc_kernel(VisTcParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* sX = smem;
    uint8_t* sY = smem + X_BYTES;
    uint8_t* sW = sY + Y_BYTES;
    uint64_t* bar_w = reinterpret_cast<uint64_t*>(sW + W_BYTES);
    uint64_t* bar_mma = bar_w + 1;                     // [3], one per layer
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_w + 4);
    float* s_f = reinterpret_cast<float*>(bar_w + 5);  // 65 floats

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.z;
    const int x0 = max(0, min((int)blockIdx.x * TXO, p.W - TXO));
    const int y0 = blockIdx.y * TY;
    const int xs = x0 - 3;
    const size_t plane = (size_t)p.H * p.W;
    const uint32_t sX_u = tc::smem_u32(sX), sY_u = tc::smem_u32(sY), sW_u = tc::smem_u32(sW);

    if (warp == 0) tc::tmem_alloc(tmem_slot, TMEM_COLS);
    if (threadIdx.x == 32) {
        tc::mbar_init(bar_w, 1);
        for (int i = 0; i < 3; ++i) tc::mbar_init(bar_mma + i, 1);
        tc::mbar_fence_init();
    }
    if (threadIdx.x < 65) s_f[threadIdx.x] = __ldg(p.fparams + threadIdx.x);
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    if (threadIdx.x == 0) {
        tc::mbar_expect_tx(bar_w, W_BYTES);
        tc::bulk_copy_g2s(sW_u, p.wgt, W_BYTES, bar_w);
    }
    if (warp < 4) {   // every MMA accumulates: the accumulator slots start zeroed
#pragma unroll
        for (int c = 0; c < R_A1 * 16; c += 16) tmem_zero16(tmem + ((uint32_t)(warp * 32) << 16) + c);
        tmem_st_wait();
    }
    // ---- stage the two fp32 input maps as hi/lo fp16 slabs (zero outside the image) -------------------------------
    static_assert((R_IN * TX) % NT == 0, "input staging: whole rounds");
    float in_e[R_IN * TX / NT], in_c[R_IN * TX / NT];
#pragma unroll
    for (int k = 0; k < R_IN * TX / NT; ++k) {   // every load in flight before the first is used
        const int i = threadIdx.x + k * NT;
        const int gx = xs + i % TX, gy = y0 - 3 + i / TX;
        const bool ok = gx >= 0 && gx < p.W && gy >= 0 && gy < p.H;
        in_e[k] = ok ? __ldg(p.entropy + n * plane + (size_t)gy * p.W + gx) : 0.f;
        in_c[k] = ok ? __ldg(p.curv + n * plane + (size_t)gy * p.W + gx) : 0.f;
    }
#pragma unroll
    for (int k = 0; k < R_IN * TX / NT; ++k) {
        const int i = threadIdx.x + k * NT;
        const float e = in_e[k], c = in_c[k];
        __half eh = __float2half_rn(e), ch = __float2half_rn(c);
        __half el = __float2half_rn(e - __half2float(eh)), cl = __float2half_rn(c - __half2float(ch));
        uint4 v;
        __half2* h = reinterpret_cast<__half2*>(&v);
        h[0] = __halves2half2(eh, ch);
        h[1] = __halves2half2(el, cl);
        h[2] = __floats2half2_rn(0.f, 0.f);
        h[3] = h[2];
        *reinterpret_cast<uint4*>(sX + (size_t)i * 16) = v;
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::mbar_wait(bar_w, 0);

    const uint32_t warp_u = tc::uniform((uint32_t)warp), tmem_u = tc::uniform(tmem);
    const bool elected = tc::elect_one();
    const int quad = warp & 3, half = warp >> 2;       // TMEM lane quadrant of this warp; which half of the row units it drains
    const int j = quad * 32 + lane;                    // MMA row; its result is pixel j + 1 of the output buffer
    const int gx1 = xs + j + 1;
    const uint32_t lane_addr = tmem + ((uint32_t)(quad * 32) << 16);

    // ---- layer 1: in8 -> a1 ------------------------------------------------------------------------------------------
    tc::tc_fence_after();
    if (warp_u == 0) {
        issue_layer<R_IN, MMA_L1, true, 0, 0>(sX_u >> 4, sW_u >> 4, tmem_u, elected);
        if (elected) tc::mma_commit(bar_mma);
        __syncwarp();
    }
    tc::mbar_wait(bar_mma, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int u = half; u < R_A1; u += 2) {
        const int gy = y0 - 2 + u;
        uint4 lo, hi;
        act16(lane_addr + u * 16, s_f, gx1 >= 0 && gx1 < p.W && gy >= 0 && gy < p.H, true, lo, hi);
        if (j < TX - 2) {
            *reinterpret_cast<uint4*>(sY + ((size_t)u * TX + j + 1) * 16) = lo;
            *reinterpret_cast<uint4*>(sY + ((size_t)(R_A1 + u) * TX + j + 1) * 16) = hi;
        }
    }
    tmem_st_wait();
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    // ---- layer 2: a1 -> a2 (re-using the in8 region) ------------------------------------------------------------------
    if (warp_u == 0) {
        issue_layer<R_A1, MMA_L23, false, R_A1 * ROW_BYTES, MMA_L1>(sY_u >> 4, sW_u >> 4, tmem_u, elected);
        if (elected) tc::mma_commit(bar_mma + 1);
        __syncwarp();
    }
    tc::mbar_wait(bar_mma + 1, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int u = half; u < R_A2; u += 2) {
        const int gy = y0 - 1 + u;
        uint4 lo, hi;
        act16(lane_addr + u * 16, s_f + 16, gx1 >= 0 && gx1 < p.W && gy >= 0 && gy < p.H, true, lo, hi);
        if (j < TX - 2) {
            *reinterpret_cast<uint4*>(sX + ((size_t)u * TX + j + 1) * 16) = lo;
            *reinterpret_cast<uint4*>(sX + ((size_t)(R_A2 + u) * TX + j + 1) * 16) = hi;
        }
    }
    tmem_st_wait();
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    // ---- layer 3 + 1x1 + sigmoid ---------------------------------------------------------------------------------------
    if (warp_u == 0) {
        issue_layer<R_A2, MMA_L23, false, R_A2 * ROW_BYTES, MMA_L1 + MMA_L23>(sX_u >> 4, sW_u >> 4, tmem_u, elected);
        if (elected) tc::mma_commit(bar_mma + 2);
        __syncwarp();
    }
    tc::mbar_wait(bar_mma + 2, 0);
    tc::tc_fence_after();
    const bool col_ok = j >= 2 && j < TX - 4 && gx1 < p.W && gx1 >= (int)blockIdx.x * TXO;   // one owner tile per pixel
#pragma unroll 1
    for (int u = half; u < TY; u += 2) {
        const int gy = y0 + u;
        uint32_t r0[8], r1[8];
        tc::tmem_ld8_nowait(lane_addr + u * 16, r0);
        tc::tmem_ld8_nowait(lane_addr + u * 16 + 8, r1);
        tc::tmem_ld_wait();
        float s = s_f[64];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            s += fmaxf(__uint_as_float(r0[i]) + s_f[32 + i], 0.f) * s_f[48 + i];
            s += fmaxf(__uint_as_float(r1[i]) + s_f[40 + i], 0.f) * s_f[56 + i];
        }
        if (col_ok && gy < p.H) p.vis[n * plane + (size_t)gy * p.W + gx1] = 1.f / (1.f + __expf(-s));
    }
    tc::tc_fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

extern "C" {

int cds_visnet_tc_supported(int h, int w) { return h >= 1 && w >= 8; }
int cds_visnet_tc_weight_halfs(void) { return (int)(W_BYTES / 2); }

// wgt_packed: fp16 operand image (cds_visnet_tc_weight_halfs() halfs); fparams: b1[16] b2[16] b3[16] w4[16] b4[1] fp32
int cds_visnet_tc(const float* entropy, const float* curv, const void* wgt_packed, const float* fparams, int n, int h, int w,
                  float* vis, cudaStream_t stream) {
    CDS_REQUIRE(entropy && curv && wgt_packed && fparams && vis, CDS_EARG, "cds_visnet_tc: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && cds_visnet_tc_supported(h, w), CDS_EUNSUPPORTED, "cds_visnet_tc: needs w >= 8 (got %dx%d)", h, w);
    constexpr size_t smem = (size_t)X_BYTES + Y_BYTES + W_BYTES + 8 * 5 + 65 * 4 + 32;
    cudaError_t e = cudaFuncSetAttribute(visnet_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cds_set_error("cds_visnet_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    VisTcParams p{entropy, curv, (const __half*)wgt_packed, fparams, vis, h, w};
    dim3 grid(cds_div_up(w, TXO), cds_div_up(h, TY), n);
    visnet_tc_kernel<<<grid, NT, smem, stream>>>(p);
    return cds_check_launch("cds_visnet_tc");
}

}  // extern "C"
