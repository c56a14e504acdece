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
// Each layer: M = 128 pixels of a row (output pixel = j+1), N = 16, K = 9 taps x channels; TMEM accumulators are
// read back by the 4 warps, bias+ReLU, ZEROED outside the image (each layer zero-pads its own input), packed to
// fp16 and stored as the next layer's operand; layer 3 finishes with the 1x1 conv + sigmoid and writes fp32.
#include <algorithm>

#include "cds_common.cuh"
#include "tc_common.cuh"

namespace {

#ifndef CDS_VIS_TY
#define CDS_VIS_TY 4   // 8-row tiles (2 resident CTAs, 17 % fewer MMAs) measured 0.476 vs 0.466 ms at stage 3: not adopted
#endif
constexpr int TX = 128, TXO = TX - 6, TY = CDS_VIS_TY;
constexpr int ROW_BYTES = TX * 16;
constexpr int R_IN = TY + 6, R_A1 = TY + 4, R_A2 = TY + 2;
constexpr uint32_t X_BYTES = 2 * R_A2 * ROW_BYTES;       // in8 (R_IN rows, 1 slab) then a2 (2 slabs x R_A2 rows)
constexpr uint32_t Y_BYTES = 2 * R_A1 * ROW_BYTES;       // a1
constexpr int MMA_L1 = 5, MMA_L23 = 9;
constexpr uint32_t W_BYTES = (MMA_L1 + 2 * MMA_L23) * 2 * 16 * 16;
static_assert(X_BYTES >= R_IN * ROW_BYTES, "in8 must fit in the region a2 re-uses");
constexpr uint32_t TMEM_COLS = R_A1 * 16 <= 128 ? 128 : 256;   // R_A1 row units x 16 accumulator columns
static_assert(R_A1 * 16 <= 256, "row units of layer 1 exceed the TMEM budget of two resident CTAs");

struct VisTcParams {
    const float* entropy;   // [n][H][W]
    const float* curv;      // [n][H][W]
    const __half* wgt;      // packed fp16: L1 [5], L2 [9], L3 [9] MMAs x [2 k-chunk][2][8 n][8 k]
    const float* fparams;   // b1[16] b2[16] b3[16] w4[16] b4[1]
    float* vis;             // [n][H][W]
    int H, W;
};

constexpr uint32_t kDescHi = (128u >> 4) | (1u << 14);

// layer 1: one slab per tap -> [tap0, pad], [tap1, tap2], ...
template <int J>
__device__ __forceinline__ void mma_l1(uint32_t a_base, uint32_t b_base, uint32_t acc, bool elected) {
    constexpr int t0 = J == 0 ? 0 : 2 * J - 1, t1 = 2 * J;
    constexpr uint32_t off0 = (uint32_t)((t0 / 3) * ROW_BYTES + (t0 % 3) * 16);
    constexpr uint32_t lbo = J == 0 ? 16u : (uint32_t)((t1 / 3) * ROW_BYTES + (t1 % 3) * 16) - off0;
    constexpr uint32_t a_const = (off0 >> 4) | ((lbo >> 4) << 16);
    constexpr uint32_t b_const = (((uint32_t)J * 512) >> 4) | ((256u >> 4) << 16);
    if (elected)
        tc::mma_f16(acc, ((uint64_t)kDescHi << 32) | (a_base + a_const), ((uint64_t)kDescHi << 32) | (b_base + b_const),
                    tc::instr_desc_f16(128, 16), J > 0);
}
// layers 2/3: two slabs (channel chunks) per tap; CHUNK = distance between the chunks of the operand buffer
template <int J, uint32_t CHUNK, int WOFF>
__device__ __forceinline__ void mma_l23(uint32_t a_base, uint32_t b_base, uint32_t acc, bool elected) {
    constexpr uint32_t off0 = (uint32_t)((J / 3) * ROW_BYTES + (J % 3) * 16);
    constexpr uint32_t a_const = (off0 >> 4) | ((CHUNK >> 4) << 16);
    constexpr uint32_t b_const = (((uint32_t)(WOFF + J) * 512) >> 4) | ((256u >> 4) << 16);
    if (elected)
        tc::mma_f16(acc, ((uint64_t)kDescHi << 32) | (a_base + a_const), ((uint64_t)kDescHi << 32) | (b_base + b_const),
                    tc::instr_desc_f16(128, 16), J > 0);
}
template <int... J>
__device__ __forceinline__ void issue_l1(uint32_t a, uint32_t b, uint32_t acc, bool e, std::integer_sequence<int, J...>) {
    (mma_l1<J>(a, b, acc, e), ...);
}
template <uint32_t CHUNK, int WOFF, int... J>
__device__ __forceinline__ void issue_l23(uint32_t a, uint32_t b, uint32_t acc, bool e, std::integer_sequence<int, J...>) {
    (mma_l23<J, CHUNK, WOFF>(a, b, acc, e), ...);
}

// bias + ReLU of one row unit's 16 accumulator columns; result as two packed fp16 slabs (or zero outside the image)
__device__ __forceinline__ void act16(uint32_t taddr, const float* __restrict__ bias, bool inside, uint4& lo, uint4& hi) {
    uint32_t r0[8], r1[8];
    tc::tmem_ld8_nowait(taddr, r0);
    tc::tmem_ld8_nowait(taddr + 8, r1);
    tc::tmem_ld_wait();
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

constexpr int NT = 256;   // 8 warps: two per TMEM lane quadrant share each layer's epilogue rows; all eight issue MMAs

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
        for (int i = 0; i < 3; ++i) tc::mbar_init(bar_mma + i, NT / 32);
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
    // ---- stage the two fp32 input maps as hi/lo fp16 slabs (zero outside the image) -------------------------------
    for (int i = threadIdx.x; i < R_IN * TX; i += NT) {
        const int px = i % TX, ry = i / TX;
        const int gx = xs + px, gy = y0 - 3 + ry;
        float e = 0.f, c = 0.f;
        if (gx >= 0 && gx < p.W && gy >= 0 && gy < p.H) {
            e = __ldg(p.entropy + n * plane + (size_t)gy * p.W + gx);
            c = __ldg(p.curv + n * plane + (size_t)gy * p.W + gx);
        }
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
#pragma unroll 1
    for (uint32_t u = warp_u; u < (uint32_t)R_A1; u += NT / 32)
        issue_l1((sX_u + u * ROW_BYTES) >> 4, sW_u >> 4, tmem_u + u * 16, elected, std::make_integer_sequence<int, MMA_L1>{});
    if (elected) tc::mma_commit(bar_mma);
    __syncwarp();
    tc::mbar_wait(bar_mma, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int u = half; u < R_A1; u += 2) {
        const int gy = y0 - 2 + u;
        uint4 lo, hi;
        act16(lane_addr + u * 16, s_f, gx1 >= 0 && gx1 < p.W && gy >= 0 && gy < p.H, lo, hi);
        if (j < TX - 2) {
            *reinterpret_cast<uint4*>(sY + ((size_t)u * TX + j + 1) * 16) = lo;
            *reinterpret_cast<uint4*>(sY + ((size_t)(R_A1 + u) * TX + j + 1) * 16) = hi;
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    // ---- layer 2: a1 -> a2 (re-using the in8 region) ------------------------------------------------------------------
#pragma unroll 1
    for (uint32_t u = warp_u; u < (uint32_t)R_A2; u += NT / 32)
        issue_l23<R_A1 * ROW_BYTES, MMA_L1>((sY_u + u * ROW_BYTES) >> 4, sW_u >> 4, tmem_u + u * 16, elected,
                                           std::make_integer_sequence<int, MMA_L23>{});
    if (elected) tc::mma_commit(bar_mma + 1);
    __syncwarp();
    tc::mbar_wait(bar_mma + 1, 0);
    tc::tc_fence_after();
#pragma unroll 1
    for (int u = half; u < R_A2; u += 2) {
        const int gy = y0 - 1 + u;
        uint4 lo, hi;
        act16(lane_addr + u * 16, s_f + 16, gx1 >= 0 && gx1 < p.W && gy >= 0 && gy < p.H, lo, hi);
        if (j < TX - 2) {
            *reinterpret_cast<uint4*>(sX + ((size_t)u * TX + j + 1) * 16) = lo;
            *reinterpret_cast<uint4*>(sX + ((size_t)(R_A2 + u) * TX + j + 1) * 16) = hi;
        }
    }
    tc::fence_proxy_async();
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();

    // ---- layer 3 + 1x1 + sigmoid ---------------------------------------------------------------------------------------
#pragma unroll 1
    for (uint32_t u = warp_u; u < (uint32_t)TY; u += NT / 32)
        issue_l23<R_A2 * ROW_BYTES, MMA_L1 + MMA_L23>((sX_u + u * ROW_BYTES) >> 4, sW_u >> 4, tmem_u + u * 16, elected,
                                                     std::make_integer_sequence<int, MMA_L23>{});
    if (elected) tc::mma_commit(bar_mma + 2);
    __syncwarp();
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
