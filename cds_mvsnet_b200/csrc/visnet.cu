// Visibility net (A3): cat(entropy, |curvature|) -> 3x (conv3x3 + BN + ReLU, 16 ch) -> conv1x1 + bias
// -> sigmoid, evaluated as ONE kernel: the 16-channel intermediates of a 32x16 output tile live in
// shared memory (receptive field 7x7), so only the two input maps are read and one map written.
// Reference: models/model.py:14,51 ; models/module.py:169-198 (ConvBnReLU). BatchNorm uses
// running statistics (eval) and is folded into the conv weights/bias by the host.
//
// Packed weights (fp32): L1 [9][2][16], b1[16], L2 [9][16][16], b2[16], L3 [9][16][16], b3[16],
// w4[16], b4[1]   (tap-major, input channel, output channel fastest).
#include "cds_common.cuh"

namespace {

constexpr int TW = 32, TH = 16, CH = 16;
constexpr int W_L1 = 0, B_L1 = W_L1 + 9 * 2 * CH, W_L2 = B_L1 + CH, B_L2 = W_L2 + 9 * CH * CH, W_L3 = B_L2 + CH,
              B_L3 = W_L3 + 9 * CH * CH, W_L4 = B_L3 + CH, B_L4 = W_L4 + CH, W_TOTAL = B_L4 + 1;

// in:  [CIN][rh+2][rw+2] region (smem), out: [CH][rh][rw] region (smem); two vertically adjacent
// pixels per thread so every weight fetched from shared memory feeds two FMAs.
// (oy0, ox0) = image coordinates of out region element (0,0); positions outside the image are
// written as 0 (each conv layer zero-pads ITS input, module.py:191).
template <int CIN, bool LAST>
__device__ __forceinline__ void conv_layer(const float* __restrict__ in, float* __restrict__ out, const float* __restrict__ wgt,
                                           const float* __restrict__ bias, int rh, int rw, int oy0, int ox0, int H, int W,
                                           const float* __restrict__ w4, float b4, float* __restrict__ gout) {
    const int iw = rw + 2, ih = rh + 2;
    const int pairs = (rh / 2) * rw;
    for (int p = threadIdx.x; p < pairs; p += blockDim.x) {
        int xx = p % rw, yy = (p / rw) * 2;
        float acc0[CH], acc1[CH];
#pragma unroll
        for (int c = 0; c < CH; ++c) { acc0[c] = bias[c]; acc1[c] = bias[c]; }
        for (int ci = 0; ci < CIN; ++ci) {
            const float* ip = in + (size_t)ci * ih * iw + (size_t)yy * iw + xx;
            float v[4][3];
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int k = 0; k < 3; ++k) v[r][k] = ip[r * iw + k];
#pragma unroll
            for (int ky = 0; ky < 3; ++ky)
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const float4* wp = reinterpret_cast<const float4*>(wgt + ((ky * 3 + kx) * CIN + ci) * CH);
                    float a = v[ky][kx], b = v[ky + 1][kx];
#pragma unroll
                    for (int q = 0; q < CH / 4; ++q) {
                        float4 ww = wp[q];
                        acc0[4 * q + 0] += a * ww.x; acc0[4 * q + 1] += a * ww.y;
                        acc0[4 * q + 2] += a * ww.z; acc0[4 * q + 3] += a * ww.w;
                        acc1[4 * q + 0] += b * ww.x; acc1[4 * q + 1] += b * ww.y;
                        acc1[4 * q + 2] += b * ww.z; acc1[4 * q + 3] += b * ww.w;
                    }
                }
        }
        int gx = ox0 + xx, gy = oy0 + yy;
        bool in0 = gx >= 0 && gx < W && gy >= 0 && gy < H;
        bool in1 = gx >= 0 && gx < W && gy + 1 >= 0 && gy + 1 < H;
        if (!LAST) {
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                out[(size_t)c * rh * rw + (size_t)yy * rw + xx] = in0 ? fmaxf(acc0[c], 0.f) : 0.f;
                out[(size_t)c * rh * rw + (size_t)(yy + 1) * rw + xx] = in1 ? fmaxf(acc1[c], 0.f) : 0.f;
            }
        } else {
            float s0 = b4, s1 = b4;
#pragma unroll
            for (int c = 0; c < CH; ++c) {
                s0 += fmaxf(acc0[c], 0.f) * w4[c];
                s1 += fmaxf(acc1[c], 0.f) * w4[c];
            }
            if (in0) gout[(size_t)gy * W + gx] = 1.f / (1.f + __expf(-s0));
            if (in1) gout[(size_t)(gy + 1) * W + gx] = 1.f / (1.f + __expf(-s1));
        }
    }
}

__global__ void __launch_bounds__(256) visnet_kernel(const float* __restrict__ entropy, const float* __restrict__ curv,
                                                     const float* __restrict__ wpack, int H, int W, float* __restrict__ vis) {
    extern __shared__ float sm[];
    float* s_w = sm;                                   // W_TOTAL (padded to a multiple of 4)
    float* s_in = s_w + ((W_TOTAL + 3) & ~3);          // [2][TH+6][TW+6]
    float* s_a1 = s_in + 2 * (TH + 6) * (TW + 6);      // [16][TH+4][TW+4]
    float* s_a2 = s_a1 + CH * (TH + 4) * (TW + 4);     // [16][TH+2][TW+2]
    const int n = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;
    const size_t plane = (size_t)H * W;
    for (int i = threadIdx.x; i < W_TOTAL; i += blockDim.x) s_w[i] = __ldg(wpack + i);
    for (int i = threadIdx.x; i < 2 * (TH + 6) * (TW + 6); i += blockDim.x) {
        int c = i / ((TH + 6) * (TW + 6));
        int r = i % ((TH + 6) * (TW + 6));
        int gy = y0 - 3 + r / (TW + 6), gx = x0 - 3 + r % (TW + 6);
        float val = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) val = __ldg((c == 0 ? entropy : curv) + n * plane + (size_t)gy * W + gx);
        s_in[i] = val;
    }
    __syncthreads();
    conv_layer<2, false>(s_in, s_a1, s_w + W_L1, s_w + B_L1, TH + 4, TW + 4, y0 - 2, x0 - 2, H, W, nullptr, 0.f, nullptr);
    __syncthreads();
    conv_layer<CH, false>(s_a1, s_a2, s_w + W_L2, s_w + B_L2, TH + 2, TW + 2, y0 - 1, x0 - 1, H, W, nullptr, 0.f, nullptr);
    __syncthreads();
    conv_layer<CH, true>(s_a2, nullptr, s_w + W_L3, s_w + B_L3, TH, TW, y0, x0, H, W, s_w + W_L4, s_w[B_L4], vis + n * plane);
}

}  // namespace

extern "C" {

int cds_visnet_weight_floats(void) { return W_TOTAL; }

int cds_visnet(const float* entropy, const float* curv, const float* wpack, int n, int h, int w, float* vis,
               cudaStream_t stream) {
    CDS_REQUIRE(entropy && curv && wpack && vis, CDS_EARG, "cds_visnet: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && h > 0 && w > 0, CDS_ESHAPE, "cds_visnet: bad shape n=%d h=%d w=%d", n, h, w);
    size_t smem = sizeof(float) * (((W_TOTAL + 3) & ~3) + 2 * (TH + 6) * (TW + 6) + CH * (TH + 4) * (TW + 4) + CH * (TH + 2) * (TW + 2));
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(visnet_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { cds_set_error("cds_visnet: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
        configured = true;
    }
    dim3 grid(cds_div_up(w, TW), cds_div_up(h, TH), n);
    visnet_kernel<<<grid, 256, smem, stream>>>(entropy, curv, wpack, h, w, vis);
    return cds_check_launch("cds_visnet");
}

}  // extern "C"
