// Training support of DynamicConv: the k x k convolutions of its branches (feature conv + curvature conv of a
// branch run as ONE convolution with Cout + 3 output channels) in fp32 on planar NCHW tensors, forward and backward.
//
// Reference semantics: models/dynamic_conv.py:84-88,112-116 (nn.Conv2d(in_c, 3 | out_c, k, padding=(k-1)//2), stride 1);
// gradients as torch.autograd computes them.  The light per-pixel part of the layer (curvature form, gate MLP with its
// BatchNorm2d, softmax, blend; models/dynamic_conv.py:100-121) stays in the host module (train2d.py).
//   * conv2d_fwd: direct convolution, odd k <= 11, stride 1, pad (k-1)/2.  Also the input gradient (flipped, transposed weights).
//   * conv2d_wgrad: dw[ci][tap][co] = sum_{b,p} g[b,co,p] x[b,ci,p + tap - h].
// Weights tap-major [Cin][k*k][Cout] (permuted on the device by the host side).
#include <algorithm>

#include "cds_common.cuh"

namespace {

constexpr int kCoT = 8;     // output channels per thread
constexpr int kCiT = 8;     // input channels per weight stage (k = 11: 8 x 121 x 8 floats = 31 KB)
constexpr int kMaxK = 11;

// grid (ceil(W / 128) * ceil(Cout / 8), H, B), 128 threads: one output pixel x 8 output channels per thread
__global__ void __launch_bounds__(128) conv2d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int Cin, int Cout, int H,
                                                        int W, int k, float* __restrict__ out) {
    __shared__ float sw[kCiT * kMaxK * kMaxK * kCoT];
    const int xt = (W + 127) / 128, h = (k - 1) / 2, kk = k * k;
    const int co0 = (blockIdx.x / xt) * kCoT, ox = (blockIdx.x % xt) * 128 + threadIdx.x;
    const int oy = blockIdx.y, b = blockIdx.z;
    float acc[kCoT];
#pragma unroll
    for (int j = 0; j < kCoT; ++j) acc[j] = 0.f;
    const size_t plane = (size_t)H * W;
    for (int c0 = 0; c0 < Cin; c0 += kCiT) {
        const int nci = min(kCiT, Cin - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nci * kk * kCoT; i += 128) {
            const int j = i % kCoT, t = (i / kCoT) % kk, c = i / (kCoT * kk);
            sw[i] = co0 + j < Cout ? __ldg(w + ((size_t)(c0 + c) * kk + t) * Cout + co0 + j) : 0.f;
        }
        __syncthreads();
        if (ox < W) {
            for (int c = 0; c < nci; ++c) {
                const float* xc = x + ((size_t)b * Cin + c0 + c) * plane;
                for (int dy = 0; dy < k; ++dy) {
                    const int iy = oy + dy - h;
                    if (iy < 0 || iy >= H) continue;
                    const float* xr = xc + (size_t)iy * W;
                    const float* wr = sw + (c * kk + dy * k) * kCoT;
                    for (int dx = 0; dx < k; ++dx) {
                        const int ix = ox + dx - h;
                        if (ix < 0 || ix >= W) continue;
                        const float v = __ldg(xr + ix);
#pragma unroll
                        for (int j = 0; j < kCoT; ++j) acc[j] = fmaf(v, wr[dx * kCoT + j], acc[j]);
                    }
                }
            }
        }
    }
    if (ox < W) {
#pragma unroll
        for (int j = 0; j < kCoT; ++j)
            if (co0 + j < Cout) out[((size_t)b * Cout + co0 + j) * plane + (size_t)oy * W + ox] = acc[j];
    }
}

// grid (pixel chunks, Cin, ceil(Cout / 4) * k), 256 threads: one kernel ROW (k taps) x 4 output channels per thread
constexpr int kWgCo = 4, kWgPix = 16;
__global__ void __launch_bounds__(256) conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, int B, int Cin, int Cout,
                                                          int H, int W, int k, float* __restrict__ dw) {
    const int ci = blockIdx.y, dy = blockIdx.z % k, co0 = (blockIdx.z / k) * kWgCo, h = (k - 1) / 2;
    const long long plane = (long long)H * W, total = plane * B;
    float acc[kMaxK][kWgCo];
#pragma unroll
    for (int t = 0; t < kMaxK; ++t)
#pragma unroll
        for (int j = 0; j < kWgCo; ++j) acc[t][j] = 0.f;
    const long long base = (long long)blockIdx.x * 256 * kWgPix;
    for (int it = 0; it < kWgPix; ++it) {
        const long long i = base + (long long)it * 256 + threadIdx.x;
        if (i >= total) break;
        const int b = (int)(i / plane);
        const long long p = i % plane;
        const int ox = (int)(p % W), oy = (int)(p / W);
        const int iy = oy + dy - h;
        if (iy < 0 || iy >= H) continue;
        float gv[kWgCo];
#pragma unroll
        for (int j = 0; j < kWgCo; ++j) gv[j] = co0 + j < Cout ? __ldg(g + ((size_t)b * Cout + co0 + j) * plane + p) : 0.f;
        const float* xr = x + ((size_t)b * Cin + ci) * plane + (size_t)iy * W;
#pragma unroll
        for (int dx = 0; dx < kMaxK; ++dx) {
            const int ix = ox + dx - h;
            const float v = (dx < k && ix >= 0 && ix < W) ? __ldg(xr + ix) : 0.f;
#pragma unroll
            for (int j = 0; j < kWgCo; ++j) acc[dx][j] = fmaf(v, gv[j], acc[dx][j]);
        }
    }
    __shared__ float red[8][kMaxK * kWgCo];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < kMaxK; ++t)
#pragma unroll
        for (int j = 0; j < kWgCo; ++j) {
            const float s = warp_sum(acc[t][j]);
            if (lane == 0) red[warp][t * kWgCo + j] = s;
        }
    __syncthreads();
    if (threadIdx.x < k * kWgCo) {
        float s = 0.f;
        for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
        const int dx = threadIdx.x / kWgCo, j = threadIdx.x % kWgCo;
        if (co0 + j < Cout) atomicAdd(dw + ((size_t)ci * k * k + dy * k + dx) * Cout + co0 + j, s);
    }
}

}  // namespace

extern "C" {

int cds_train_conv2d(const float* x, const float* wgt, int B, int Cin, int Cout, int H, int W, int k, float* out, cudaStream_t stream) {
    CDS_REQUIRE(x && wgt && out, CDS_EARG, "cds_train_conv2d: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && H > 0 && H <= 65535 && W > 0, CDS_ESHAPE, "cds_train_conv2d: bad shape");
    CDS_REQUIRE(k >= 1 && k <= kMaxK && (k & 1), CDS_EUNSUPPORTED, "cds_train_conv2d: odd kernel sizes up to 11 (got %d)", k);
    dim3 grid(cds_div_up(W, 128) * cds_div_up(Cout, kCoT), H, B);
    conv2d_fwd_kernel<<<grid, 128, 0, stream>>>(x, wgt, Cin, Cout, H, W, k, out);
    return cds_check_launch("cds_train_conv2d");
}

int cds_train_conv2d_wgrad(const float* x, const float* g, int B, int Cin, int Cout, int H, int W, int k, float* dw, cudaStream_t stream) {
    CDS_REQUIRE(x && g && dw, CDS_EARG, "cds_train_conv2d_wgrad: null pointer");
    CDS_REQUIRE(B > 0 && Cin > 0 && Cin <= 65535 && Cout > 0 && H > 0 && W > 0, CDS_ESHAPE, "cds_train_conv2d_wgrad: bad shape");
    CDS_REQUIRE(k >= 1 && k <= kMaxK && (k & 1), CDS_EUNSUPPORTED, "cds_train_conv2d_wgrad: odd kernel sizes up to 11 (got %d)", k);
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)Cin * k * k * Cout * sizeof(float), stream);
    if (e != cudaSuccess) { cds_set_error("cds_train_conv2d_wgrad: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    const long long total = (long long)B * H * W;
    dim3 grid(cds_div_up(total, 256 * kWgPix), Cin, cds_div_up(Cout, kWgCo) * k);
    CDS_REQUIRE(grid.z <= 65535, CDS_ESHAPE, "cds_train_conv2d_wgrad: too many output channels");
    conv2d_wgrad_kernel<<<grid, 256, 0, stream>>>(x, g, B, Cin, Cout, H, W, k, dw);
    return cds_check_launch("cds_train_conv2d_wgrad");
}

}  // extern "C"
