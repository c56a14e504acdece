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

// grid (ceil(W / 512) * ceil(Cout / 8), H, B), 128 threads: a RUN of four output pixels x 8 output channels per thread -- the k
// taps of the run share k + 3 input columns and every weight read from shared memory feeds four pixels
constexpr int kRun = 4;
template <int K>
__global__ void __launch_bounds__(128) conv2d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int Cin, int Cout, int H,
                                                        int W, float* __restrict__ out) {
    __shared__ float sw[kCiT * K * K * kCoT];
    constexpr int h = (K - 1) / 2, kk = K * K, NCOL = K + kRun - 1;
    const int xt = (W + 128 * kRun - 1) / (128 * kRun);
    const int co0 = (blockIdx.x / xt) * kCoT, ox0 = ((blockIdx.x % xt) * 128 + threadIdx.x) * kRun;
    const int oy = blockIdx.y, b = blockIdx.z;
    float acc[kRun][kCoT];
#pragma unroll
    for (int v = 0; v < kRun; ++v)
#pragma unroll
        for (int j = 0; j < kCoT; ++j) acc[v][j] = 0.f;
    const size_t plane = (size_t)H * W;
    for (int c0 = 0; c0 < Cin; c0 += kCiT) {
        const int nci = min(kCiT, Cin - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nci * kk * kCoT; i += 128) {
            const int j = i % kCoT, t = (i / kCoT) % kk, c = i / (kCoT * kk);
            sw[i] = co0 + j < Cout ? __ldg(w + ((size_t)(c0 + c) * kk + t) * Cout + co0 + j) : 0.f;
        }
        __syncthreads();
        if (ox0 < W) {
            for (int c = 0; c < nci; ++c) {
                const float* xc = x + ((size_t)b * Cin + c0 + c) * plane;
#pragma unroll 1
                for (int dy = 0; dy < K; ++dy) {
                    const int iy = oy + dy - h;
                    if (iy < 0 || iy >= H) continue;
                    const float* xr = xc + (size_t)iy * W;
                    float col[NCOL];
#pragma unroll
                    for (int q = 0; q < NCOL; ++q) {
                        const int ix = ox0 - h + q;
                        col[q] = (ix >= 0 && ix < W) ? __ldg(xr + ix) : 0.f;
                    }
                    const float* wr = sw + (c * kk + dy * K) * kCoT;
#pragma unroll
                    for (int dx = 0; dx < K; ++dx) {
                        float wv[kCoT];
#pragma unroll
                        for (int j = 0; j < kCoT; ++j) wv[j] = wr[dx * kCoT + j];
#pragma unroll
                        for (int v = 0; v < kRun; ++v)
#pragma unroll
                            for (int j = 0; j < kCoT; ++j) acc[v][j] = fmaf(col[v + dx], wv[j], acc[v][j]);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < kRun; ++v)
        if (ox0 + v < W) {
#pragma unroll
            for (int j = 0; j < kCoT; ++j)
                if (co0 + j < Cout) out[((size_t)b * Cout + co0 + j) * plane + (size_t)oy * W + ox0 + v] = acc[v][j];
        }
}

template <int K>
void launch_fwd2d(const float* x, const float* w, int B, int Cin, int Cout, int H, int W, float* out, cudaStream_t st) {
    dim3 grid(cds_div_up(W, 128 * kRun) * cds_div_up(Cout, kCoT), H, B);
    conv2d_fwd_kernel<K><<<grid, 128, 0, st>>>(x, w, Cin, Cout, H, W, out);
}

// grid (pixel-run chunks, Cin, ceil(Cout / 4) * k), 256 threads: one kernel ROW (k taps) x 4 output channels per thread; a thread
// walks runs of four consecutive pixels of a row, whose k taps share k + 3 input columns
constexpr int kWgCo = 4, kWgRun = 4, kWgRuns = 4;
template <int K>
__global__ void __launch_bounds__(256) conv2d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, int B, int Cin, int Cout,
                                                          int H, int W, float* __restrict__ dw) {
    constexpr int h = (K - 1) / 2, NCOL = K + kWgRun - 1;
    const int ci = blockIdx.y, dy = blockIdx.z % K, co0 = (blockIdx.z / K) * kWgCo;
    const int wq = (W + kWgRun - 1) / kWgRun;
    const long long plane = (long long)H * W, total = (long long)B * H * wq;
    float acc[K][kWgCo];
#pragma unroll
    for (int t = 0; t < K; ++t)
#pragma unroll
        for (int j = 0; j < kWgCo; ++j) acc[t][j] = 0.f;
    const long long base = (long long)blockIdx.x * 256 * kWgRuns;
    for (int it = 0; it < kWgRuns; ++it) {
        const long long i = base + (long long)it * 256 + threadIdx.x;
        if (i >= total) break;
        const int q = (int)(i % wq);
        const long long row = i / wq;
        const int oy = (int)(row % H), b = (int)(row / H);
        const int iy = oy + dy - h;
        if (iy < 0 || iy >= H) continue;
        const int ox0 = q * kWgRun;
        float gv[kWgRun][kWgCo];
#pragma unroll
        for (int v = 0; v < kWgRun; ++v)
#pragma unroll
            for (int j = 0; j < kWgCo; ++j)
                gv[v][j] = (ox0 + v < W && co0 + j < Cout) ? __ldg(g + ((size_t)b * Cout + co0 + j) * plane + (size_t)oy * W + ox0 + v) : 0.f;
        const float* xr = x + ((size_t)b * Cin + ci) * plane + (size_t)iy * W;
        float col[NCOL];
#pragma unroll
        for (int c = 0; c < NCOL; ++c) {
            const int ix = ox0 - h + c;
            col[c] = (ix >= 0 && ix < W) ? __ldg(xr + ix) : 0.f;
        }
#pragma unroll
        for (int dx = 0; dx < K; ++dx)
#pragma unroll
            for (int v = 0; v < kWgRun; ++v)
#pragma unroll
                for (int j = 0; j < kWgCo; ++j) acc[dx][j] = fmaf(col[v + dx], gv[v][j], acc[dx][j]);
    }
    __shared__ float red[8][K * kWgCo];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < K; ++t)
#pragma unroll
        for (int j = 0; j < kWgCo; ++j) {
            const float s = warp_sum(acc[t][j]);
            if (lane == 0) red[warp][t * kWgCo + j] = s;
        }
    __syncthreads();
    if (threadIdx.x < K * kWgCo) {
        float s = 0.f;
        for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
        const int dx = threadIdx.x / kWgCo, j = threadIdx.x % kWgCo;
        if (co0 + j < Cout) atomicAdd(dw + ((size_t)ci * K * K + dy * K + dx) * Cout + co0 + j, s);
    }
}

template <int K>
void launch_wgrad2d(const float* x, const float* g, int B, int Cin, int Cout, int H, int W, float* dw, cudaStream_t st) {
    const long long total = (long long)B * H * ((W + kWgRun - 1) / kWgRun);
    dim3 grid(cds_div_up(total, 256 * kWgRuns), Cin, cds_div_up(Cout, kWgCo) * K);
    conv2d_wgrad_kernel<K><<<grid, 256, 0, st>>>(x, g, B, Cin, Cout, H, W, dw);
}

}  // namespace

extern "C" {

int cds_train_conv2d(const float* x, const float* wgt, int B, int Cin, int Cout, int H, int W, int k, float* out, cudaStream_t stream) {
    CDS_REQUIRE(x && wgt && out, CDS_EARG, "cds_train_conv2d: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && Cin > 0 && Cout > 0 && H > 0 && H <= 65535 && W > 0, CDS_ESHAPE, "cds_train_conv2d: bad shape");
    CDS_REQUIRE(k >= 1 && k <= kMaxK && (k & 1), CDS_EUNSUPPORTED, "cds_train_conv2d: odd kernel sizes up to 11 (got %d)", k);
    switch (k) {
        case 1: launch_fwd2d<1>(x, wgt, B, Cin, Cout, H, W, out, stream); break;
        case 3: launch_fwd2d<3>(x, wgt, B, Cin, Cout, H, W, out, stream); break;
        case 5: launch_fwd2d<5>(x, wgt, B, Cin, Cout, H, W, out, stream); break;
        case 7: launch_fwd2d<7>(x, wgt, B, Cin, Cout, H, W, out, stream); break;
        case 9: launch_fwd2d<9>(x, wgt, B, Cin, Cout, H, W, out, stream); break;
        default: launch_fwd2d<11>(x, wgt, B, Cin, Cout, H, W, out, stream); break;
    }
    return cds_check_launch("cds_train_conv2d");
}

int cds_train_conv2d_wgrad(const float* x, const float* g, int B, int Cin, int Cout, int H, int W, int k, float* dw, cudaStream_t stream) {
    CDS_REQUIRE(x && g && dw, CDS_EARG, "cds_train_conv2d_wgrad: null pointer");
    CDS_REQUIRE(B > 0 && Cin > 0 && Cin <= 65535 && Cout > 0 && H > 0 && W > 0, CDS_ESHAPE, "cds_train_conv2d_wgrad: bad shape");
    CDS_REQUIRE(k >= 1 && k <= kMaxK && (k & 1), CDS_EUNSUPPORTED, "cds_train_conv2d_wgrad: odd kernel sizes up to 11 (got %d)", k);
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)Cin * k * k * Cout * sizeof(float), stream);
    if (e != cudaSuccess) { cds_set_error("cds_train_conv2d_wgrad: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    CDS_REQUIRE(cds_div_up(Cout, kWgCo) * k <= 65535, CDS_ESHAPE, "cds_train_conv2d_wgrad: too many output channels");
    switch (k) {
        case 1: launch_wgrad2d<1>(x, g, B, Cin, Cout, H, W, dw, stream); break;
        case 3: launch_wgrad2d<3>(x, g, B, Cin, Cout, H, W, dw, stream); break;
        case 5: launch_wgrad2d<5>(x, g, B, Cin, Cout, H, W, dw, stream); break;
        case 7: launch_wgrad2d<7>(x, g, B, Cin, Cout, H, W, dw, stream); break;
        case 9: launch_wgrad2d<9>(x, g, B, Cin, Cout, H, W, dw, stream); break;
        default: launch_wgrad2d<11>(x, g, B, Cin, Cout, H, W, dw, stream); break;
    }
    return cds_check_launch("cds_train_conv2d_wgrad");
}

}  // extern "C"
