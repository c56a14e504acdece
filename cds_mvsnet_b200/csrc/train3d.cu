// Training support of the 3-D regulariser: the blocks of CostRegNet in TRAINING mode and their backward.
//
// Reference semantics: models/module.py:80-122 (Conv3d k3 p1 stride 1|2, no bias -> BatchNorm3d with BATCH statistics ->
// ReLU), :125-166 (ConvTranspose3d k3 s2 p1 op1 -> BatchNorm3d -> ReLU), :303-315 (wiring, skip adds after the ReLU, prob head).
// Gradients are what torch.autograd computes for that graph.
//
// Everything here is fp32 on planar NCDHW tensors, the layout the reference's modules exchange (the inference path keeps its
// channel-blocked fp16 layout and tensor-core kernels; training runs at small crops where the convolutions are a minor cost
// next to keeping the gradients in fp32):
//   * conv3d_fwd: direct k3 p1 convolution, stride 1|2.  Also the input gradient of a stride-1 block (flipped, transposed
//     weights) and of a transposed block (its weight read as a stride-2 conv weight).
//   * deconv3d_fwd: direct transposed convolution k3 s2 p1 op1 (gather form).  Also the input gradient of a stride-2 block.
//   * conv3d_wgrad: dW[ci][tap][co] = sum_{b,o} g[b,co,o] x[b,ci,o*s-1+tap]; the transposed block's weight gradient is the same
//     sum with the roles of input and gradient swapped.
//   * bn_stats / bn_apply / bn_backward_reduce / bn_backward_apply: BatchNorm3d with batch statistics + ReLU (+ skip add after
//     the ReLU), forward and backward; per-channel sums in fp64.
// Weights are passed tap-major, [Cin][27][Cout] (conv: w[co][ci][k] permuted; transposed conv: w[ci][co][k] permuted): the host
// side permutes the parameter on the device.
#include <algorithm>

#include "cds_common.cuh"

namespace {

constexpr int kCoT = 8;    // output channels per thread
constexpr int kCiT = 16;   // input channels per weight stage

// ---- direct convolution k3 p1, stride s -------------------------------------------------------------------------------
// grid (ceil(Wo / 512) * ceil(Cout / 8), Ho, B * Do), 128 threads: a RUN of four output voxels of a row x 8 output channels per
// thread -- the 3 kw taps of the run share its 6 (stride 1) or 9 (stride 2) input columns, every weight feeds four voxels
constexpr int kRun = 4;
template <int STRIDE>
__global__ void __launch_bounds__(128) conv3d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int B, int Cin, int Cout,
                                                        int Di, int Hi, int Wi, int Do, int Ho, int Wo, float* __restrict__ out) {
    __shared__ float sw[kCiT * 27 * kCoT];
    constexpr int NCOL = (kRun - 1) * STRIDE + 3;
    const int xt = (Wo + 128 * kRun - 1) / (128 * kRun);
    const int co0 = (blockIdx.x / xt) * kCoT, ox0 = ((blockIdx.x % xt) * 128 + threadIdx.x) * kRun;
    const int oy = blockIdx.y, od = blockIdx.z % Do, b = blockIdx.z / Do;
    float acc[kRun][kCoT];
#pragma unroll
    for (int v = 0; v < kRun; ++v)
#pragma unroll
        for (int j = 0; j < kCoT; ++j) acc[v][j] = 0.f;
    const size_t plane = (size_t)Hi * Wi, vol = plane * Di;
    const int ix0 = ox0 * STRIDE - 1;
    for (int c0 = 0; c0 < Cin; c0 += kCiT) {
        const int nci = min(kCiT, Cin - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nci * 27 * kCoT; i += 128) {
            const int j = i % kCoT, t = (i / kCoT) % 27, c = i / (kCoT * 27);
            sw[i] = co0 + j < Cout ? __ldg(w + ((size_t)(c0 + c) * 27 + t) * Cout + co0 + j) : 0.f;
        }
        __syncthreads();
        if (ox0 < Wo) {
            for (int c = 0; c < nci; ++c) {
                const float* xc = x + ((size_t)b * Cin + c0 + c) * vol;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int id = od * STRIDE - 1 + kd;
                    if (id < 0 || id >= Di) continue;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int iy = oy * STRIDE - 1 + kh;
                        if (iy < 0 || iy >= Hi) continue;
                        const float* xr = xc + (size_t)id * plane + (size_t)iy * Wi;
                        float col[NCOL];
#pragma unroll
                        for (int q = 0; q < NCOL; ++q) col[q] = (ix0 + q >= 0 && ix0 + q < Wi) ? __ldg(xr + ix0 + q) : 0.f;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const float* wt = sw + (c * 27 + (kd * 3 + kh) * 3 + kw) * kCoT;
                            float wv[kCoT];
#pragma unroll
                            for (int j = 0; j < kCoT; ++j) wv[j] = wt[j];
#pragma unroll
                            for (int v = 0; v < kRun; ++v)
#pragma unroll
                                for (int j = 0; j < kCoT; ++j) acc[v][j] = fmaf(col[v * STRIDE + kw], wv[j], acc[v][j]);
                        }
                    }
                }
            }
        }
    }
    const size_t ovol = (size_t)Do * Ho * Wo;
#pragma unroll
    for (int v = 0; v < kRun; ++v)
        if (ox0 + v < Wo) {
#pragma unroll
            for (int j = 0; j < kCoT; ++j)
                if (co0 + j < Cout) out[((size_t)b * Cout + co0 + j) * ovol + ((size_t)od * Ho + oy) * Wo + ox0 + v] = acc[v][j];
        }
}

// ---- direct transposed convolution k3 s2 p1 op1: out [B,Cout,2D,2H,2W], out[o] = sum_{k, i: o = 2i - 1 + k} x[i] w[k] -------
__global__ void __launch_bounds__(128) deconv3d_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, int B, int Cin, int Cout,
                                                          int Di, int Hi, int Wi, float* __restrict__ out) {
    __shared__ float sw[kCiT * 27 * kCoT];
    const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
    const int xt = (Wo + 127) / 128;
    const int co0 = (blockIdx.x / xt) * kCoT, ox = (blockIdx.x % xt) * 128 + threadIdx.x;
    const int oy = blockIdx.y, od = blockIdx.z % Do, b = blockIdx.z / Do;
    float acc[kCoT];
#pragma unroll
    for (int j = 0; j < kCoT; ++j) acc[j] = 0.f;
    const size_t plane = (size_t)Hi * Wi, vol = plane * Di;
    for (int c0 = 0; c0 < Cin; c0 += kCiT) {
        const int nci = min(kCiT, Cin - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nci * 27 * kCoT; i += 128) {
            const int j = i % kCoT, t = (i / kCoT) % 27, c = i / (kCoT * 27);
            sw[i] = co0 + j < Cout ? __ldg(w + ((size_t)(c0 + c) * 27 + t) * Cout + co0 + j) : 0.f;
        }
        __syncthreads();
        if (ox < Wo) {
            for (int c = 0; c < nci; ++c) {
                const float* xc = x + ((size_t)b * Cin + c0 + c) * vol;
#pragma unroll
                for (int kd = 0; kd < 3; ++kd) {
                    const int td = od + 1 - kd;
                    if (td < 0 || (td & 1) || (td >> 1) >= Di) continue;
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const int ty = oy + 1 - kh;
                        if (ty < 0 || (ty & 1) || (ty >> 1) >= Hi) continue;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw) {
                            const int tx = ox + 1 - kw;
                            if (tx < 0 || (tx & 1) || (tx >> 1) >= Wi) continue;
                            const float v = __ldg(xc + (size_t)(td >> 1) * plane + (size_t)(ty >> 1) * Wi + (tx >> 1));
                            const float* wt = sw + (c * 27 + (kd * 3 + kh) * 3 + kw) * kCoT;
#pragma unroll
                            for (int j = 0; j < kCoT; ++j) acc[j] = fmaf(v, wt[j], acc[j]);
                        }
                    }
                }
            }
        }
    }
    if (ox < Wo) {
        const size_t ovol = (size_t)Do * Ho * Wo;
#pragma unroll
        for (int j = 0; j < kCoT; ++j)
            if (co0 + j < Cout) out[((size_t)b * Cout + co0 + j) * ovol + ((size_t)od * Ho + oy) * Wo + ox] = acc[j];
    }
}

// ---- weight gradient: dw[ci][tap][co] += sum over the block's output voxels of g[co][o] * x[ci][o*s - 1 + tap] --------------
// grid (voxel-run chunks, Cin, ceil(Cout / 4)), 256 threads; fp32 atomics onto a zeroed dw.  A thread walks RUNS of four
// consecutive output voxels of a row: the 3 kw taps of the run share its 6 (stride 1) or 9 (stride 2) input columns, so a
// (kd, kh) row costs 6-9 loads for 48 FMAs instead of 12.
constexpr int kWgCo = 4, kWgRun = 4, kWgRuns = 4;   // output channels per thread; voxels per run; runs per thread
template <int STRIDE>
__global__ void __launch_bounds__(256) conv3d_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ g, int B, int Cin, int Cout,
                                                          int Di, int Hi, int Wi, int Do, int Ho, int Wo, float* __restrict__ dw) {
    constexpr int NCOL = (kWgRun - 1) * STRIDE + 3;
    const int ci = blockIdx.y, co0 = blockIdx.z * kWgCo;
    const int wq = (Wo + kWgRun - 1) / kWgRun;
    const long long ovol = (long long)Do * Ho * Wo, rows = (long long)B * Do * Ho, total = rows * wq;
    const size_t plane = (size_t)Hi * Wi, vol = plane * Di;
    float acc[27][kWgCo];
#pragma unroll
    for (int t = 0; t < 27; ++t)
#pragma unroll
        for (int j = 0; j < kWgCo; ++j) acc[t][j] = 0.f;
    const long long base = (long long)blockIdx.x * 256 * kWgRuns;
    for (int it = 0; it < kWgRuns; ++it) {
        const long long i = base + (long long)it * 256 + threadIdx.x;
        if (i >= total) break;
        const int q = (int)(i % wq);
        const long long row = i / wq;
        const int oy = (int)(row % Ho), od = (int)((row / Ho) % Do), b = (int)(row / ((long long)Ho * Do));
        const int ox0 = q * kWgRun;
        float gv[kWgRun][kWgCo];
#pragma unroll
        for (int v = 0; v < kWgRun; ++v)
#pragma unroll
            for (int j = 0; j < kWgCo; ++j)
                gv[v][j] = (ox0 + v < Wo && co0 + j < Cout)
                               ? __ldg(g + ((size_t)b * Cout + co0 + j) * ovol + ((size_t)od * Ho + oy) * Wo + ox0 + v) : 0.f;
        const float* xc = x + ((size_t)b * Cin + ci) * vol;
        const int ix0 = ox0 * STRIDE - 1;
#pragma unroll
        for (int kd = 0; kd < 3; ++kd) {
            const int id = od * STRIDE - 1 + kd;
            if (id < 0 || id >= Di) continue;
#pragma unroll
            for (int kh = 0; kh < 3; ++kh) {
                const int iy = oy * STRIDE - 1 + kh;
                if (iy < 0 || iy >= Hi) continue;
                const float* xr = xc + (size_t)id * plane + (size_t)iy * Wi;
                float col[NCOL];
#pragma unroll
                for (int c = 0; c < NCOL; ++c) col[c] = (ix0 + c >= 0 && ix0 + c < Wi) ? __ldg(xr + ix0 + c) : 0.f;
#pragma unroll
                for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                    for (int v = 0; v < kWgRun; ++v)
#pragma unroll
                        for (int j = 0; j < kWgCo; ++j)
                            acc[(kd * 3 + kh) * 3 + kw][j] = fmaf(col[v * STRIDE + kw], gv[v][j], acc[(kd * 3 + kh) * 3 + kw][j]);
            }
        }
    }
    __shared__ float red[8][27 * kWgCo];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
    for (int t = 0; t < 27; ++t)
#pragma unroll
        for (int j = 0; j < kWgCo; ++j) {
            const float s = warp_sum(acc[t][j]);
            if (lane == 0) red[warp][t * kWgCo + j] = s;
        }
    __syncthreads();
    if (threadIdx.x < 27 * kWgCo) {
        float s = 0.f;
        for (int wv = 0; wv < 8; ++wv) s += red[wv][threadIdx.x];
        const int t = threadIdx.x / kWgCo, j = threadIdx.x % kWgCo;
        if (co0 + j < Cout) atomicAdd(dw + ((size_t)ci * 27 + t) * Cout + co0 + j, s);
    }
}

// ---- BatchNorm3d with batch statistics --------------------------------------------------------------------------------
// sums[c] = (sum x, sum x^2) over (b, voxels); grid (chunks, C, B)
__global__ void __launch_bounds__(256) bn_stats_kernel(const float* __restrict__ x, int C, long long V, double* __restrict__ sums) {
    const int c = blockIdx.y, b = blockIdx.z;
    const float* p = x + ((size_t)b * C + c) * V;
    float s = 0.f, q = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
        const float v = __ldg(p + i);
        s += v;
        q = fmaf(v, v, q);
    }
    __shared__ double rs[8], rq[8];
    const double ds = (double)warp_sum(s), dq = (double)warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = ds; rq[threadIdx.x >> 5] = dq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, bq = 0.0;
        for (int i = 0; i < 8; ++i) { a += rs[i]; bq += rq[i]; }
        atomicAdd(sums + 2 * c, a);
        atomicAdd(sums + 2 * c + 1, bq);
    }
}

// y = relu(gamma (x - mean) rstd + beta) (+ skip, added AFTER the ReLU: models/module.py:310-312)
__global__ void __launch_bounds__(256) bn_apply_kernel(const float* __restrict__ x, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                      const float* __restrict__ gamma, const float* __restrict__ beta,
                                                      const float* __restrict__ skip, int relu, int C, long long V,
                                                      float* __restrict__ y) {
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t off = ((size_t)b * C + c) * V;
    const float sc = __ldg(gamma + c) * __ldg(rstd + c), sh = __ldg(beta + c) - __ldg(mean + c) * sc;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
        float v = fmaf(__ldg(x + off + i), sc, sh);
        if (relu) v = fmaxf(v, 0.f);
        if (skip) v += __ldg(skip + off + i);
        y[off + i] = v;
    }
}

// backward, pass 1: dz = dy * [z > 0] (z recomputed from x), sums[c] = (sum dz, sum dz * xhat)
__global__ void __launch_bounds__(256) bn_backward_reduce_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                const float* __restrict__ mean, const float* __restrict__ rstd,
                                                                const float* __restrict__ gamma, const float* __restrict__ beta, int relu,
                                                                int C, long long V, double* __restrict__ sums) {
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t off = ((size_t)b * C + c) * V;
    const float m = __ldg(mean + c), r = __ldg(rstd + c), ga = __ldg(gamma + c), be = __ldg(beta + c);
    float s = 0.f, q = 0.f;
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
        const float xh = (__ldg(x + off + i) - m) * r;
        float d = __ldg(dy + off + i);
        if (relu && fmaf(ga, xh, be) <= 0.f) d = 0.f;
        s += d;
        q = fmaf(d, xh, q);
    }
    __shared__ double rs[8], rq[8];
    const double ds = (double)warp_sum(s), dq = (double)warp_sum(q);
    if ((threadIdx.x & 31) == 0) { rs[threadIdx.x >> 5] = ds; rq[threadIdx.x >> 5] = dq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, bq = 0.0;
        for (int i = 0; i < 8; ++i) { a += rs[i]; bq += rq[i]; }
        atomicAdd(sums + 2 * c, a);
        atomicAdd(sums + 2 * c + 1, bq);
    }
}

// backward, pass 2: dx = gamma rstd (dz - mean(dz) - xhat mean(dz xhat)); dgamma = sum dz xhat, dbeta = sum dz are sums[]
__global__ void __launch_bounds__(256) bn_backward_apply_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                               const float* __restrict__ mean, const float* __restrict__ rstd,
                                                               const float* __restrict__ gamma, const float* __restrict__ beta,
                                                               const double* __restrict__ sums, int relu, int C, long long V, double count,
                                                               float* __restrict__ dx) {
    const int c = blockIdx.y, b = blockIdx.z;
    const size_t off = ((size_t)b * C + c) * V;
    const float m = __ldg(mean + c), r = __ldg(rstd + c), ga = __ldg(gamma + c), be = __ldg(beta + c);
    const float mdz = (float)(sums[2 * c] / count), mdzx = (float)(sums[2 * c + 1] / count);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < V; i += (long long)gridDim.x * 256) {
        const float xh = (__ldg(x + off + i) - m) * r;
        float d = __ldg(dy + off + i);
        if (relu && fmaf(ga, xh, be) <= 0.f) d = 0.f;
        dx[off + i] = ga * r * (d - mdz - xh * mdzx);
    }
}

int chunks_for(long long V) { return (int)std::min<long long>(std::max<long long>(1, (V + 256 * 8 - 1) / (256 * 8)), 1024); }

}  // namespace

extern "C" {

int cds_train_conv3d(const float* x, const float* wgt, int B, int Cin, int Cout, int D, int H, int W, int stride, float* out,
                     cudaStream_t stream) {
    CDS_REQUIRE(x && wgt && out, CDS_EARG, "cds_train_conv3d: null pointer");
    CDS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0 && (stride == 1 || stride == 2), CDS_ESHAPE,
                "cds_train_conv3d: bad shape / stride");
    const int Do = (D + stride - 1) / stride, Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    CDS_REQUIRE(Ho <= 65535 && (long long)B * Do <= 65535, CDS_ESHAPE, "cds_train_conv3d: volume too large for the launch grid");
    dim3 grid(cds_div_up(Wo, 128 * kRun) * cds_div_up(Cout, kCoT), Ho, B * Do);
    if (stride == 1) conv3d_fwd_kernel<1><<<grid, 128, 0, stream>>>(x, wgt, B, Cin, Cout, D, H, W, Do, Ho, Wo, out);
    else conv3d_fwd_kernel<2><<<grid, 128, 0, stream>>>(x, wgt, B, Cin, Cout, D, H, W, Do, Ho, Wo, out);
    return cds_check_launch("cds_train_conv3d");
}

int cds_train_deconv3d(const float* x, const float* wgt, int B, int Cin, int Cout, int D, int H, int W, float* out, cudaStream_t stream) {
    CDS_REQUIRE(x && wgt && out, CDS_EARG, "cds_train_deconv3d: null pointer");
    CDS_REQUIRE(B > 0 && Cin > 0 && Cout > 0 && D > 0 && H > 0 && W > 0, CDS_ESHAPE, "cds_train_deconv3d: bad shape");
    CDS_REQUIRE(2 * H <= 65535 && (long long)B * 2 * D <= 65535, CDS_ESHAPE, "cds_train_deconv3d: volume too large for the launch grid");
    dim3 grid(cds_div_up(2 * W, 128) * cds_div_up(Cout, kCoT), 2 * H, B * 2 * D);
    deconv3d_fwd_kernel<<<grid, 128, 0, stream>>>(x, wgt, B, Cin, Cout, D, H, W, out);
    return cds_check_launch("cds_train_deconv3d");
}

int cds_train_conv3d_wgrad(const float* x, const float* g, int B, int Cin, int Cout, int D, int H, int W, int stride, float* dw,
                           cudaStream_t stream) {
    CDS_REQUIRE(x && g && dw, CDS_EARG, "cds_train_conv3d_wgrad: null pointer");
    CDS_REQUIRE(B > 0 && Cin > 0 && Cin <= 65535 && Cout > 0 && D > 0 && H > 0 && W > 0 && (stride == 1 || stride == 2), CDS_ESHAPE,
                "cds_train_conv3d_wgrad: bad shape / stride");
    const int Do = (D + stride - 1) / stride, Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    cudaError_t e = cudaMemsetAsync(dw, 0, (size_t)Cin * 27 * Cout * sizeof(float), stream);
    if (e != cudaSuccess) { cds_set_error("cds_train_conv3d_wgrad: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    const long long total = (long long)B * Do * Ho * ((Wo + kWgRun - 1) / kWgRun);
    dim3 grid(cds_div_up(total, 256 * kWgRuns), Cin, cds_div_up(Cout, kWgCo));
    if (stride == 1) conv3d_wgrad_kernel<1><<<grid, 256, 0, stream>>>(x, g, B, Cin, Cout, D, H, W, Do, Ho, Wo, dw);
    else conv3d_wgrad_kernel<2><<<grid, 256, 0, stream>>>(x, g, B, Cin, Cout, D, H, W, Do, Ho, Wo, dw);
    return cds_check_launch("cds_train_conv3d_wgrad");
}

int cds_train_bn_stats(const float* x, int B, int C, long long V, double* sums, cudaStream_t stream) {
    CDS_REQUIRE(x && sums, CDS_EARG, "cds_train_bn_stats: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && C > 0 && C <= 65535 && V > 0, CDS_ESHAPE, "cds_train_bn_stats: bad shape");
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)C * 2 * sizeof(double), stream);
    if (e != cudaSuccess) { cds_set_error("cds_train_bn_stats: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    bn_stats_kernel<<<dim3(chunks_for(V), C, B), 256, 0, stream>>>(x, C, V, sums);
    return cds_check_launch("cds_train_bn_stats");
}

int cds_train_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* skip,
                       int relu, int B, int C, long long V, float* y, cudaStream_t stream) {
    CDS_REQUIRE(x && mean && rstd && gamma && beta && y, CDS_EARG, "cds_train_bn_apply: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && C > 0 && C <= 65535 && V > 0, CDS_ESHAPE, "cds_train_bn_apply: bad shape");
    bn_apply_kernel<<<dim3(chunks_for(V), C, B), 256, 0, stream>>>(x, mean, rstd, gamma, beta, skip, relu, C, V, y);
    return cds_check_launch("cds_train_bn_apply");
}

int cds_train_bn_backward(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                          int relu, int B, int C, long long V, double* sums, float* dx, cudaStream_t stream) {
    CDS_REQUIRE(dy && x && mean && rstd && gamma && beta && sums && dx, CDS_EARG, "cds_train_bn_backward: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && C > 0 && C <= 65535 && V > 0, CDS_ESHAPE, "cds_train_bn_backward: bad shape");
    cudaError_t e = cudaMemsetAsync(sums, 0, (size_t)C * 2 * sizeof(double), stream);
    if (e != cudaSuccess) { cds_set_error("cds_train_bn_backward: cudaMemsetAsync: %s", cudaGetErrorString(e)); return (int)e; }
    const dim3 grid(chunks_for(V), C, B);
    bn_backward_reduce_kernel<<<grid, 256, 0, stream>>>(dy, x, mean, rstd, gamma, beta, relu, C, V, sums);
    int rc = cds_check_launch("cds_train_bn_backward (reduce)");
    if (rc) return rc;
    bn_backward_apply_kernel<<<grid, 256, 0, stream>>>(dy, x, mean, rstd, gamma, beta, sums, relu, C, V, (double)B * (double)V, dx);
    return cds_check_launch("cds_train_bn_backward (apply)");
}

}  // extern "C"
