// Curvature-guided dynamic-scale convolution (A6) and the 2-D plumbing of the feature extractor (A7).
//
// Reference: models/dynamic_conv.py:97-122 (DynamicConv.forward), models/module.py:28-71 (Conv2d =
// conv -> InstanceNorm2d(affine=False) -> LeakyReLU(0.1)), :236-267 (FeatureNet wiring).
//
// Design: activations are channels-last [n, H, W, C]; a conv kernel writes its RAW (pre-norm) output
// plus per-(image, channel) sum / sum-of-squares (fp64 atomics), and every CONSUMER applies
// (x - mean) * rstd and the activation while loading its input tile, so InstanceNorm never costs a
// pass over memory.  One DynamicConv launch evaluates all K kernel sizes from one haloed input tile
// in shared memory:
//   phase A  the 3 curvature channels (a,b,c) of every branch -> curv_k = a u^2 + 2b uv + c v^2
//   gate     w = softmax(W2 relu(BN(W1 curv)) / T) per pixel (BN folded)
//   phase B  sum_k w_k * conv_k(x), accumulated in ONE register tile by scaling the input sample
//            with the pixel's gate weight (conv is linear), + sum_k w_k * bias_k
// A thread owns two vertically adjacent pixels so each weight fetched from smem feeds two FMAs.
#include <algorithm>

#include "cds_common.cuh"

namespace {

constexpr int TW = 32, TH = 16;   // output tile per 256-thread block
constexpr float kInEps = 1e-5f;
constexpr int kMaxK = 3;          // kernel sizes per DynamicConv

#define ACT_NONE 0
#define ACT_LRELU 1
#define ACT_TANH 2

__device__ __forceinline__ float apply_act(float x, int act) {
    if (act == ACT_LRELU) return x > 0.f ? x : 0.1f * x;
    if (act == ACT_TANH) return tanhf(x);
    return x;
}

// mean / rstd of channel c of image n from the fp64 (sum, sumsq) accumulators
__device__ __forceinline__ void norm_coeffs(const double* __restrict__ stats, int n, int C, int c, double count, float& mean,
                                            float& rstd) {
    double s = stats[((size_t)n * C + c) * 2], ss = stats[((size_t)n * C + c) * 2 + 1];
    double m = s / count;
    double var = ss / count - m * m;
    if (var < 0.0) var = 0.0;
    mean = (float)m;
    rstd = (float)(1.0 / sqrt(var + (double)kInEps));
}

// block-wide per-channel (sum, sumsq) -> fp64 atomics.  vals: this thread's NP pixels x C channels.
template <int C, int NP>
__device__ __forceinline__ void accumulate_stats(const float (&vals)[NP][C], const bool (&valid)[NP], double* __restrict__ stats,
                                                 float* s_red /* [warps][C][2] */) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int p = 0; p < NP; ++p)
            if (valid[p]) { s += vals[p][c]; ss += vals[p][c] * vals[p][c]; }
        s = warp_sum(s);
        ss = warp_sum(ss);
        if (lane == 0) { s_red[(warp * C + c) * 2] = s; s_red[(warp * C + c) * 2 + 1] = ss; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += (double)s_red[w * C * 2 + i];
        atomicAdd(stats + i, t);
    }
}

constexpr int kChunks = 8;   // pixel chunks (of blockDim pixels) per block in the small 2-D conv kernels

// same reduction for per-thread running (sum, sumsq) accumulators
template <int C>
__device__ __forceinline__ void accumulate_stats2(const float (&st)[2][C], double* __restrict__ stats, float* s_red) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        float s = warp_sum(st[0][c]), ss = warp_sum(st[1][c]);
        if (lane == 0) { s_red[(warp * C + c) * 2] = s; s_red[(warp * C + c) * 2 + 1] = ss; }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < C * 2; i += blockDim.x) {
        double t = 0.0;
        for (int w = 0; w < nwarps; ++w) t += (double)s_red[w * C * 2 + i];
        atomicAdd(stats + i, t);
    }
}

struct DynParams {
    const void* x;            // input activations
    const int* img_index;     // in_mode 1: image used by batch item n (ref image shared by several items)
    const double* in_stats;   // [n][CIN][2] or null (no normalisation on load)
    const float* epipole;     // [n][2] (x, y) at full resolution
    const float* w_att;       // per branch: [k*k][CIN][4]  (a, b, c, 0)
    const float* w_conv;      // per branch: [k*k][CIN][COUT]
    const float* bias;        // [NK][COUT] or null
    const float* gate;        // W1f [4][NK], b1 [4], W2 [NK][4]   (BN folded into W1f/b1)
    void* out_raw;            // [n][H][W][COUT]
    double* out_stats;        // [n][COUT][2]
    float* norm_curv;         // [n][H][W] or null
    float* nc_sq;             // [n][H][W] running (nc_a^2 + nc_b^2 + nc_c^2)/3 or null
    float* nc_abs;            // [n][H][W] or null
    int in_mode, in_act, nc_mode;
    int H, W, k[kMaxK];
    float epi_scale, inv_temperature;
};

template <typename T, int CIN, int COUT, int NK>
__global__ void __launch_bounds__(256) dynconv_kernel(DynParams p) {
    extern __shared__ __align__(16) float sm[];
    const int H = p.H, W = p.W;
    int kmax = 0, ntaps = 0;
#pragma unroll
    for (int i = 0; i < NK; ++i) { kmax = max(kmax, p.k[i]); ntaps += p.k[i] * p.k[i]; }
    const int halo = (kmax - 1) / 2;
    const int tw = TW + 2 * halo, th = TH + 2 * halo;
    float* s_att = sm;                               // [ntaps][CIN][4]
    float* s_conv = s_att + ntaps * CIN * 4;         // [ntaps][CIN][COUT]
    float* s_in = s_conv + ntaps * CIN * COUT;       // [CIN][th][tw]
    float* s_red = s_in + CIN * th * tw;             // [8 warps][COUT][2]
    float* s_norm = s_red + 8 * COUT * 2;            // [CIN][2] mean, rstd

    const int n = blockIdx.z;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH;

    for (int i = threadIdx.x; i < ntaps * CIN; i += blockDim.x)
        reinterpret_cast<float4*>(s_att)[i] = __ldg(reinterpret_cast<const float4*>(p.w_att) + i);
    for (int i = threadIdx.x; i < ntaps * CIN * COUT / 4; i += blockDim.x)
        reinterpret_cast<float4*>(s_conv)[i] = __ldg(reinterpret_cast<const float4*>(p.w_conv) + i);
    if (p.in_stats) {
        for (int c = threadIdx.x; c < CIN; c += blockDim.x) norm_coeffs(p.in_stats, n, CIN, c, (double)H * W, s_norm[2 * c], s_norm[2 * c + 1]);
        __syncthreads();
    }
    // ---- haloed input tile -> smem (planar fp32), normalised + activated, zero outside the image
    if (p.in_mode == 1) {
        const float* img = (const float*)p.x + (size_t)(p.img_index ? p.img_index[n] : n) * CIN * H * W;
        for (int i = threadIdx.x; i < CIN * th * tw; i += blockDim.x) {
            int c = i / (th * tw), r = i % (th * tw);
            int gy = y0 - halo + r / tw, gx = x0 - halo + r % tw;
            s_in[i] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(img + ((size_t)c * H + gy) * W + gx) : 0.f;
        }
    } else {
        if constexpr (CIN % 8 == 0) {
            const T* src = (const T*)p.x + (size_t)n * H * W * CIN;
            constexpr int C8 = CIN / 8;
            for (int i = threadIdx.x; i < th * tw * C8; i += blockDim.x) {
                int c8 = i % C8, r = i / C8;
                int ty = r / tw, tx = r % tw;
                int gy = y0 - halo + ty, gx = x0 - halo + tx;
                float v[8];
                bool inb = gy >= 0 && gy < H && gx >= 0 && gx < W;
                if (inb) Vec8<T>::load(src + ((size_t)gy * W + gx) * CIN + c8 * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    int c = c8 * 8 + j;
                    float t = 0.f;
                    if (inb) {
                        t = v[j];
                        if (p.in_stats) t = (t - s_norm[2 * c]) * s_norm[2 * c + 1];
                        t = apply_act(t, p.in_act);
                    }
                    s_in[(c * th + ty) * tw + tx] = t;
                }
            }
        }
    }
    __syncthreads();

    const int tx = threadIdx.x % TW, tp = threadIdx.x / TW;  // pixel pair (y0 + 2 tp, +1)
    const int gx = x0 + tx, gy = y0 + 2 * tp;
    // ---- phase A: curvature channels of every branch
    float curv[NK][2];
    {
        float ex = __ldg(p.epipole + 2 * n) * p.epi_scale, ey = __ldg(p.epipole + 2 * n + 1) * p.epi_scale;
        float uu[2], uv2[2], vv[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float u = (float)gx - ex, v = (float)(gy + q) - ey;
            float r = sqrtf(u * u + v * v) + 1e-6f;
            u /= r;
            v /= r;
            uu[q] = u * u;
            uv2[q] = 2.f * u * v;
            vv[q] = v * v;
        }
        int tap0 = 0;
#pragma unroll
        for (int b = 0; b < NK; ++b) {
            const int k = p.k[b], off = halo - (k - 1) / 2;
            float a0 = 0.f, b0 = 0.f, c0 = 0.f, a1 = 0.f, b1 = 0.f, c1 = 0.f;
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) {
                    const float4* wp = reinterpret_cast<const float4*>(s_att) + (size_t)(tap0 + ky * k + kx) * CIN;
                    const float* ip = s_in + (2 * tp + ky + off) * tw + tx + kx + off;
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) {
                        float4 w = wp[ci];
                        float xa = ip[ci * th * tw], xb = ip[ci * th * tw + tw];
                        a0 += xa * w.x; b0 += xa * w.y; c0 += xa * w.z;
                        a1 += xb * w.x; b1 += xb * w.y; c1 += xb * w.z;
                    }
                }
            curv[b][0] = (a0 * uu[0] + b0 * uv2[0]) + c0 * vv[0];
            curv[b][1] = (a1 * uu[1] + b1 * uv2[1]) + c1 * vv[1];
            tap0 += k * k;
        }
    }
    // ---- gate: softmax(W2 relu(W1f curv + b1) / T)
    float gw[NK][2], ncurv[2];
    {
        const float* g = p.gate;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            float hdn[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float s = __ldg(g + 4 * NK + j);
#pragma unroll
                for (int b = 0; b < NK; ++b) s += __ldg(g + j * NK + b) * curv[b][q];
                hdn[j] = fmaxf(s, 0.f);
            }
            float logit[NK], mx = -INFINITY;
#pragma unroll
            for (int b = 0; b < NK; ++b) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 4; ++j) s += __ldg(g + 4 * NK + 4 + b * 4 + j) * hdn[j];
                logit[b] = s * p.inv_temperature;
                mx = fmaxf(mx, logit[b]);
            }
            float den = 0.f;
#pragma unroll
            for (int b = 0; b < NK; ++b) { logit[b] = expf(logit[b] - mx); den += logit[b]; }
            float nc = 0.f;
#pragma unroll
            for (int b = 0; b < NK; ++b) { gw[b][q] = logit[b] / den; nc += curv[b][q] * gw[b][q]; }
            ncurv[q] = nc;
        }
    }
    // ---- phase B: gate-weighted sum of the branch convolutions in one accumulator tile
    float acc[2][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) {
        float b0 = 0.f, b1 = 0.f;
        if (p.bias) {
#pragma unroll
            for (int b = 0; b < NK; ++b) {
                float bb = __ldg(p.bias + b * COUT + c);
                b0 += gw[b][0] * bb;
                b1 += gw[b][1] * bb;
            }
        }
        acc[0][c] = b0;
        acc[1][c] = b1;
    }
    {
        int tap0 = 0;
#pragma unroll
        for (int b = 0; b < NK; ++b) {
            const int k = p.k[b], off = halo - (k - 1) / 2;
            const float g0 = gw[b][0], g1 = gw[b][1];
            for (int ky = 0; ky < k; ++ky)
                for (int kx = 0; kx < k; ++kx) {
                    const float* wp = s_conv + (size_t)(tap0 + ky * k + kx) * CIN * COUT;
                    const float* ip = s_in + (2 * tp + ky + off) * tw + tx + kx + off;
#pragma unroll
                    for (int ci = 0; ci < CIN; ++ci) {
                        float xa = ip[ci * th * tw] * g0, xb = ip[ci * th * tw + tw] * g1;
                        const float4* w4 = reinterpret_cast<const float4*>(wp + ci * COUT);
#pragma unroll
                        for (int q = 0; q < COUT / 4; ++q) {
                            float4 w = w4[q];
                            acc[0][4 * q + 0] += xa * w.x; acc[0][4 * q + 1] += xa * w.y;
                            acc[0][4 * q + 2] += xa * w.z; acc[0][4 * q + 3] += xa * w.w;
                            acc[1][4 * q + 0] += xb * w.x; acc[1][4 * q + 1] += xb * w.y;
                            acc[1][4 * q + 2] += xb * w.z; acc[1][4 * q + 3] += xb * w.w;
                        }
                    }
                }
            tap0 += k * k;
        }
    }
    // ---- epilogue
    bool valid[2] = {gx < W && gy < H, gx < W && gy + 1 < H};
    T* outp = (T*)p.out_raw + (size_t)n * H * W * COUT;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        if (!valid[q]) continue;
        size_t pix = (size_t)(gy + q) * W + gx;
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = acc[q][c8 * 8 + j];
            Vec8<T>::store(outp + pix * COUT + c8 * 8, y);
        }
        size_t m = (size_t)n * H * W + pix;
        float nc = ncurv[q];
        if (p.norm_curv) p.norm_curv[m] = nc;
        if (p.nc_sq) {
            if (p.nc_mode == 0) p.nc_sq[m] = nc * nc;
            else if (p.nc_mode == 1) p.nc_sq[m] = p.nc_sq[m] + nc * nc;
            else p.nc_sq[m] = (p.nc_sq[m] + nc * nc) / 3.f;
        }
        if (p.nc_abs) p.nc_abs[m] = fabsf(nc);
    }
    if (p.out_stats) accumulate_stats<COUT, 2>(acc, valid, p.out_stats + (size_t)n * COUT * 2, s_red);
}

// ---------------------------------------------------------------------------------------------
// 3x3 stride-2 pad-1 conv (FeatureNet.downsample1/2, module.py:214,218): normalise+LeakyReLU on load,
// raw output + statistics.  One thread per output pixel.
// ---------------------------------------------------------------------------------------------
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(128) conv3x3s2_kernel(const T* __restrict__ in, const double* __restrict__ in_stats, int in_act,
                                                        const float* __restrict__ wgt /*[9][CIN][COUT]*/, int Hi, int Wi,
                                                        T* __restrict__ out, T* __restrict__ out_lo, double* __restrict__ out_stats) {
    extern __shared__ __align__(16) float sm[];
    float* s_w = sm;                       // [9][CIN][COUT]
    float* s_norm = s_w + 9 * CIN * COUT;  // [CIN][2]
    float* s_red = s_norm + CIN * 2;       // [4 warps][COUT][2]
    const int n = blockIdx.y;
    const int Ho = (Hi + 1) / 2, Wo = (Wi + 1) / 2;
    for (int i = threadIdx.x; i < 9 * CIN * COUT / 4; i += blockDim.x)
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wgt) + i);
    for (int c = threadIdx.x; c < CIN; c += blockDim.x) {
        if (in_stats) norm_coeffs(in_stats, n, CIN, c, (double)Hi * Wi, s_norm[2 * c], s_norm[2 * c + 1]);
        else { s_norm[2 * c] = 0.f; s_norm[2 * c + 1] = 1.f; }
    }
    __syncthreads();
    const T* src = in + (size_t)n * Hi * Wi * CIN;
    float st[2][COUT];   // running (sum, sumsq) of this thread's outputs
#pragma unroll
    for (int c = 0; c < COUT; ++c) { st[0][c] = 0.f; st[1][c] = 0.f; }
    // each block walks kChunks pixel chunks so the prologue (weights, fp64 norm coefficients) is amortised
    for (int chunk = 0; chunk < kChunks; ++chunk) {
    long long m = ((long long)blockIdx.x * kChunks + chunk) * blockDim.x + threadIdx.x;
    bool live = m < (long long)Ho * Wo;
    int ox = live ? (int)(m % Wo) : 0, oy = live ? (int)(m / Wo) : 0;
    float acc[1][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[0][c] = 0.f;
    if (live) {
        for (int ky = 0; ky < 3; ++ky) {
            int iy = 2 * oy - 1 + ky;
            if (iy < 0 || iy >= Hi) continue;
            for (int kx = 0; kx < 3; ++kx) {
                int ix = 2 * ox - 1 + kx;
                if (ix < 0 || ix >= Wi) continue;
                const float* wp = s_w + (ky * 3 + kx) * CIN * COUT;
#pragma unroll
                for (int c8 = 0; c8 < CIN / 8; ++c8) {
                    float v[8];
                    Vec8<T>::load(src + ((size_t)iy * Wi + ix) * CIN + c8 * 8, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        int c = c8 * 8 + j;
                        float t = apply_act((v[j] - s_norm[2 * c]) * s_norm[2 * c + 1], in_act);
                        const float4* w4 = reinterpret_cast<const float4*>(wp + c * COUT);
#pragma unroll
                        for (int q = 0; q < COUT / 4; ++q) {
                            float4 w = w4[q];
                            acc[0][4 * q + 0] += t * w.x; acc[0][4 * q + 1] += t * w.y;
                            acc[0][4 * q + 2] += t * w.z; acc[0][4 * q + 3] += t * w.w;
                        }
                    }
                }
            }
        }
        T* op = out + ((size_t)n * Ho * Wo + m) * COUT;
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) y[j] = acc[0][c8 * 8 + j];
            Vec8<T>::store(op + c8 * 8, y);
            if (out_lo) {   // split-precision storage: the residual the storage type just dropped
                float b[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) b[j] = y[j] - to_f32<T>(from_f32<T>(y[j]));
                Vec8<T>::store(out_lo + ((size_t)n * Ho * Wo + m) * COUT + c8 * 8, b);
            }
        }
#pragma unroll
        for (int c = 0; c < COUT; ++c) { st[0][c] += acc[0][c]; st[1][c] += acc[0][c] * acc[0][c]; }
    }
    }
    if (out_stats) accumulate_stats2<COUT>(st, out_stats + (size_t)n * COUT * 2, s_red);
}

// ---------------------------------------------------------------------------------------------
// 1x1 conv over cat(nearest-up2(A), B) (FeatureNet.inner1/inner2, module.py:253-254,260-261).
// A is at half resolution; both inputs are normalised/activated on load (A optionally not: inner2
// takes the already-activated stage-2 feature).
// ---------------------------------------------------------------------------------------------
template <typename T, int CA, int CB, int COUT>
__global__ void __launch_bounds__(128) conv1x1_cat_kernel(const T* __restrict__ A, const double* __restrict__ a_stats, int a_act,
                                                          const T* __restrict__ Bp, const double* __restrict__ b_stats, int b_act,
                                                          const float* __restrict__ wgt /*[CA+CB][COUT]*/, int H, int W,
                                                          T* __restrict__ out, double* __restrict__ out_stats) {
    extern __shared__ __align__(16) float sm[];
    constexpr int CIN = CA + CB;
    float* s_w = sm;                   // [CIN][COUT]
    float* s_norm = s_w + CIN * COUT;  // [CIN][2]
    float* s_red = s_norm + CIN * 2;   // [4][COUT][2]
    const int n = blockIdx.y;
    const int Ha = H / 2, Wa = W / 2;
    for (int i = threadIdx.x; i < CIN * COUT / 4; i += blockDim.x)
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wgt) + i);
    for (int c = threadIdx.x; c < CIN; c += blockDim.x) {
        s_norm[2 * c] = 0.f;
        s_norm[2 * c + 1] = 1.f;
        if (c < CA && a_stats) norm_coeffs(a_stats, n, CA, c, (double)Ha * Wa, s_norm[2 * c], s_norm[2 * c + 1]);
        if (c >= CA && b_stats) norm_coeffs(b_stats, n, CB, c - CA, (double)H * W, s_norm[2 * c], s_norm[2 * c + 1]);
    }
    __syncthreads();
    float st[2][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { st[0][c] = 0.f; st[1][c] = 0.f; }
    for (int chunk = 0; chunk < kChunks; ++chunk) {
    long long m = ((long long)blockIdx.x * kChunks + chunk) * blockDim.x + threadIdx.x;
    bool live = m < (long long)H * W;
    int x = live ? (int)(m % W) : 0, y = live ? (int)(m / W) : 0;
    float acc[1][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[0][c] = 0.f;
    if (live) {
        const T* pa = A + (((size_t)n * Ha + y / 2) * Wa + x / 2) * CA;
        const T* pb = Bp + (((size_t)n * H + y) * W + x) * CB;
#pragma unroll
        for (int c8 = 0; c8 < CIN / 8; ++c8) {
            float v[8];
            if (c8 < CA / 8) Vec8<T>::load(pa + c8 * 8, v);
            else Vec8<T>::load(pb + (c8 - CA / 8) * 8, v);
            const int act = c8 < CA / 8 ? a_act : b_act;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                int c = c8 * 8 + j;
                float t = apply_act((v[j] - s_norm[2 * c]) * s_norm[2 * c + 1], act);
                const float4* w4 = reinterpret_cast<const float4*>(s_w + c * COUT);
#pragma unroll
                for (int q = 0; q < COUT / 4; ++q) {
                    float4 w = w4[q];
                    acc[0][4 * q + 0] += t * w.x; acc[0][4 * q + 1] += t * w.y;
                    acc[0][4 * q + 2] += t * w.z; acc[0][4 * q + 3] += t * w.w;
                }
            }
        }
        T* op = out + ((size_t)n * H * W + m) * COUT;
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            float yv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) yv[j] = acc[0][c8 * 8 + j];
            Vec8<T>::store(op + c8 * 8, yv);
        }
#pragma unroll
        for (int c = 0; c < COUT; ++c) { st[0][c] += acc[0][c]; st[1][c] += acc[0][c] * acc[0][c]; }
    }
    }
    if (out_stats) accumulate_stats2<COUT>(st, out_stats + (size_t)n * COUT * 2, s_red);
}

// Same op, one thread per 2x2 block of output pixels: the four pixels share their half-resolution A pixel, so its
// CA x COUT product is formed once and seeds their accumulators; every weight row fetched from shared memory feeds the
// four pixels (a quarter of the LDS traffic that bounded the one-pixel form); the multiply-adds are packed fp32x2 (FFMA2).
// fp32 arithmetic on the stored activations: no operand rounding beyond the storage type's own.
__device__ __forceinline__ float2 ffma2x(float2 a, float2 b, float2 c) {
    float2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(reinterpret_cast<uint64_t&>(d))
        : "l"(reinterpret_cast<uint64_t&>(a)), "l"(reinterpret_cast<uint64_t&>(b)), "l"(reinterpret_cast<uint64_t&>(c)));
    return d;
}

constexpr int kQuadChunks = 4;   // 2x2 blocks per thread

template <typename T, int CA, int CB, int COUT>
__global__ void __launch_bounds__(128) conv1x1_cat_quad_kernel(const T* __restrict__ A, const double* __restrict__ a_stats, int a_act,
                                                               const T* __restrict__ Bp, const double* __restrict__ b_stats, int b_act,
                                                               const float* __restrict__ wgt /*[CA+CB][COUT]*/, int H, int W,
                                                               T* __restrict__ out, double* __restrict__ out_stats) {
    extern __shared__ __align__(16) float sm[];
    constexpr int CIN = CA + CB, N2 = COUT / 2;
    float* s_w = sm;                   // [CIN][COUT]
    float* s_norm = s_w + CIN * COUT;  // [CIN][2]
    float* s_red = s_norm + CIN * 2;   // [4][COUT][2]
    const int n = blockIdx.y;
    const int Ha = H / 2, Wa = W / 2;
    for (int i = threadIdx.x; i < CIN * COUT / 4; i += blockDim.x)
        reinterpret_cast<float4*>(s_w)[i] = __ldg(reinterpret_cast<const float4*>(wgt) + i);
    for (int c = threadIdx.x; c < CIN; c += blockDim.x) {
        s_norm[2 * c] = 0.f;
        s_norm[2 * c + 1] = 1.f;
        if (c < CA && a_stats) norm_coeffs(a_stats, n, CA, c, (double)Ha * Wa, s_norm[2 * c], s_norm[2 * c + 1]);
        if (c >= CA && b_stats) norm_coeffs(b_stats, n, CB, c - CA, (double)H * W, s_norm[2 * c], s_norm[2 * c + 1]);
    }
    __syncthreads();
    float st[2][COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) { st[0][c] = 0.f; st[1][c] = 0.f; }
#pragma unroll 1
    for (int chunk = 0; chunk < kQuadChunks; ++chunk) {
        const long long q = ((long long)blockIdx.x * kQuadChunks + chunk) * blockDim.x + threadIdx.x;
        if (q >= (long long)Ha * Wa) continue;
        const int bx = (int)(q % Wa), by = (int)(q / Wa);
        float2 acc[4][N2];
        {   // the shared half-resolution pixel
            float2 accA[N2];
#pragma unroll
            for (int k = 0; k < N2; ++k) accA[k] = make_float2(0.f, 0.f);
            const T* pa = A + (((size_t)n * Ha + by) * Wa + bx) * CA;
#pragma unroll
            for (int c8 = 0; c8 < CA / 8; ++c8) {
                float v[8];
                Vec8<T>::load(pa + c8 * 8, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int c = c8 * 8 + j;
                    const float t = apply_act((v[j] - s_norm[2 * c]) * s_norm[2 * c + 1], a_act);
                    const float2* w2 = reinterpret_cast<const float2*>(s_w + c * COUT);
                    const float2 tt = make_float2(t, t);
#pragma unroll
                    for (int k = 0; k < N2; ++k) accA[k] = ffma2x(tt, w2[k], accA[k]);
                }
            }
#pragma unroll
            for (int p = 0; p < 4; ++p)
#pragma unroll
                for (int k = 0; k < N2; ++k) acc[p][k] = accA[k];
        }
        const T* pb = Bp + (((size_t)n * H + 2 * by) * W + 2 * bx) * CB;
#pragma unroll
        for (int c8 = 0; c8 < CB / 8; ++c8) {
            float v[4][8];
#pragma unroll
            for (int p = 0; p < 4; ++p) Vec8<T>::load(pb + ((size_t)(p >> 1) * W + (p & 1)) * CB + c8 * 8, v[p]);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = CA + c8 * 8 + j;
                const float mean = s_norm[2 * c], rstd = s_norm[2 * c + 1];
                const float2* w2 = reinterpret_cast<const float2*>(s_w + c * COUT);
                float2 tt[4];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    const float t = apply_act((v[p][j] - mean) * rstd, b_act);
                    tt[p] = make_float2(t, t);
                }
#pragma unroll
                for (int k = 0; k < N2; ++k) {
                    const float2 w = w2[k];
#pragma unroll
                    for (int p = 0; p < 4; ++p) acc[p][k] = ffma2x(tt[p], w, acc[p][k]);
                }
            }
        }
#pragma unroll
        for (int p = 0; p < 4; ++p) {
            T* op = out + (((size_t)n * H + 2 * by + (p >> 1)) * W + 2 * bx + (p & 1)) * COUT;
#pragma unroll
            for (int c8 = 0; c8 < COUT / 8; ++c8) {
                float yv[8];
#pragma unroll
                for (int j = 0; j < 4; ++j) { yv[2 * j] = acc[p][c8 * 4 + j].x; yv[2 * j + 1] = acc[p][c8 * 4 + j].y; }
                Vec8<T>::store(op + c8 * 8, yv);
            }
#pragma unroll
            for (int k = 0; k < N2; ++k) {
                st[0][2 * k] += acc[p][k].x; st[1][2 * k] += acc[p][k].x * acc[p][k].x;
                st[0][2 * k + 1] += acc[p][k].y; st[1][2 * k + 1] += acc[p][k].y * acc[p][k].y;
            }
        }
    }
    if (out_stats) accumulate_stats2<COUT>(st, out_stats + (size_t)n * COUT * 2, s_red);
}

// InstanceNorm + activation materialised (the three stage features: InstanceNorm2d -> Tanh, module.py:223,230,232)
template <typename T>
__global__ void __launch_bounds__(256) instnorm_act_kernel(const T* __restrict__ raw, const double* __restrict__ stats, int act,
                                                           int C, long long HW, T* __restrict__ out) {
    extern __shared__ float s_norm[];  // [C][2]
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) norm_coeffs(stats, n, C, c, (double)HW, s_norm[2 * c], s_norm[2 * c + 1]);
    __syncthreads();
    const int C8 = C / 8;
    long long total = HW * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c8 = (int)(i % C8);
        float v[8];
        Vec8<T>::load(raw + (size_t)n * HW * C + i * 8, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = c8 * 8 + j;
            v[j] = apply_act((v[j] - s_norm[2 * c]) * s_norm[2 * c + 1], act);
        }
        Vec8<T>::store(out + (size_t)n * HW * C + i * 8, v);
    }
}

// the same from fp16 value + residual planes to fp32 (nothing is rounded on the way)
__global__ void __launch_bounds__(256) instnorm_act_split_kernel(const __half* __restrict__ raw, const __half* __restrict__ raw_lo,
                                                                 const double* __restrict__ stats, int act, int C, long long HW,
                                                                 float* __restrict__ out, __half* __restrict__ out16) {
    extern __shared__ float s_norm[];  // [C][2]
    const int n = blockIdx.y;
    for (int c = threadIdx.x; c < C; c += blockDim.x) norm_coeffs(stats, n, C, c, (double)HW, s_norm[2 * c], s_norm[2 * c + 1]);
    __syncthreads();
    const int C8 = C / 8;
    long long total = HW * C8;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c8 = (int)(i % C8);
        float v[8], l[8];
        Vec8<__half>::load(raw + (size_t)n * HW * C + i * 8, v);
        if (raw_lo) {
            Vec8<__half>::load(raw_lo + (size_t)n * HW * C + i * 8, l);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += l[j];
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int c = c8 * 8 + j;
            v[j] = apply_act((v[j] - s_norm[2 * c]) * s_norm[2 * c + 1], act);
        }
        Vec8<float>::store(out + (size_t)n * HW * C + i * 8, v);
        if (out16) Vec8<__half>::store(out16 + (size_t)n * HW * C + i * 8, v);
    }
}

// ---- layout converters at the drop-in boundary (fp32 NCHW <-> channels-last storage type) ----
__global__ void u8_to_f32_kernel(const uchar4* __restrict__ in, long long n4, float4* __restrict__ out) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const uchar4 v = __ldg(in + i);
        out[i] = make_float4(__fdiv_rn((float)v.x, 255.f), __fdiv_rn((float)v.y, 255.f), __fdiv_rn((float)v.z, 255.f),
                             __fdiv_rn((float)v.w, 255.f));
    }
}
__global__ void u8_to_f32_tail_kernel(const unsigned char* __restrict__ in, long long from, long long n, float* __restrict__ out) {
    const long long i = from + threadIdx.x;
    if (i < n) out[i] = __fdiv_rn((float)in[i], 255.f);
}

template <typename T>
__global__ void nchw_to_nhwc_kernel(const float* __restrict__ in, int C, long long HW, T* __restrict__ out) {
    const int n = blockIdx.y;
    long long total = HW * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int c = (int)(i % C);
        long long p = i / C;
        out[(size_t)n * total + i] = from_f32<T>(__ldg(in + (size_t)n * total + (size_t)c * HW + p));
    }
}
template <typename T>
__global__ void nhwc_to_nchw_kernel(const T* __restrict__ in, int C, long long HW, float* __restrict__ out) {
    const int n = blockIdx.y;
    long long total = HW * C;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        long long p = i % HW;
        int c = (int)(i / HW);
        out[(size_t)n * total + i] = to_f32<T>(in[(size_t)n * total + p * C + c]);
    }
}

template <typename T, int CIN, int COUT, int NK>
int launch_dyn(const DynParams& p, int n, cudaStream_t st) {
    int kmax = 0, ntaps = 0;
    for (int i = 0; i < NK; ++i) { kmax = std::max(kmax, p.k[i]); ntaps += p.k[i] * p.k[i]; }
    int halo = (kmax - 1) / 2;
    size_t smem = sizeof(float) * ((size_t)ntaps * CIN * 4 + (size_t)ntaps * CIN * COUT + (size_t)CIN * (TH + 2 * halo) * (TW + 2 * halo) +
                                   8 * COUT * 2 + CIN * 2);
    if (smem > 227 * 1024) { cds_set_error("cds_dynamic_conv: tile needs %zu bytes of shared memory", smem); return CDS_EUNSUPPORTED; }
    cudaError_t e = cudaFuncSetAttribute(dynconv_kernel<T, CIN, COUT, NK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { cds_set_error("cds_dynamic_conv: cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return (int)e; }
    dim3 grid(cds_div_up(p.W, TW), cds_div_up(p.H, TH), n);
    dynconv_kernel<T, CIN, COUT, NK><<<grid, 256, smem, st>>>(p);
    return cds_check_launch("cds_dynamic_conv");
}

template <typename T>
int dispatch_dyn(const DynParams& p, int n, int Cin, int Cout, int nk, cudaStream_t st) {
#define CDS_CASE(ci, co, kk) \
    if (Cin == ci && Cout == co && nk == kk) return launch_dyn<T, ci, co, kk>(p, n, st);
    CDS_CASE(3, 8, 3) CDS_CASE(8, 8, 3) CDS_CASE(16, 16, 2) CDS_CASE(32, 32, 2) CDS_CASE(8, 8, 2)
#undef CDS_CASE
    cds_set_error("cds_dynamic_conv: unsupported (Cin=%d, Cout=%d, kernels=%d)", Cin, Cout, nk);
    return CDS_EUNSUPPORTED;
}

}  // namespace

extern "C" {

int cds_dynamic_conv(const void* x, int in_mode, const int* img_index, const double* in_stats, int in_act,
                     const float* epipole, float epi_scale, const float* w_att, const float* w_conv, const float* bias,
                     const float* gate, int n, int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes,
                     float temperature, int dtype, void* out_raw, double* out_stats, float* norm_curv, float* nc_sq,
                     int nc_mode, float* nc_abs, cudaStream_t stream) {
    CDS_REQUIRE(x && epipole && w_att && w_conv && gate && out_raw && kernel_sizes, CDS_EARG, "cds_dynamic_conv: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && H > 0 && W > 0, CDS_ESHAPE, "cds_dynamic_conv: bad shape n=%d H=%d W=%d", n, H, W);
    CDS_REQUIRE(num_kernels >= 2 && num_kernels <= kMaxK, CDS_EUNSUPPORTED, "cds_dynamic_conv: 2 or 3 kernel sizes supported");
    CDS_REQUIRE(temperature > 0.f, CDS_EARG, "cds_dynamic_conv: temperature must be positive");
    CDS_REQUIRE((in_mode == 1) == (Cin == 3), CDS_EUNSUPPORTED, "cds_dynamic_conv: planar fp32 input is for the 3-channel image only");
    DynParams p{};
    p.x = x; p.img_index = img_index; p.in_stats = in_stats; p.epipole = epipole; p.w_att = w_att; p.w_conv = w_conv;
    p.bias = bias; p.gate = gate; p.out_raw = out_raw; p.out_stats = out_stats; p.norm_curv = norm_curv; p.nc_sq = nc_sq;
    p.nc_abs = nc_abs; p.in_mode = in_mode; p.in_act = in_act; p.nc_mode = nc_mode; p.H = H; p.W = W;
    for (int i = 0; i < kMaxK; ++i) p.k[i] = 0;
    for (int i = 0; i < num_kernels; ++i) {
        CDS_REQUIRE(kernel_sizes[i] >= 1 && kernel_sizes[i] <= 11 && (kernel_sizes[i] & 1), CDS_EUNSUPPORTED,
                    "cds_dynamic_conv: kernel sizes must be odd and <= 11");
        p.k[i] = kernel_sizes[i];
    }
    p.epi_scale = epi_scale;
    p.inv_temperature = 1.f / temperature;
    if (dtype == CDS_F16) return dispatch_dyn<__half>(p, n, Cin, Cout, num_kernels, stream);
    if (dtype == CDS_F32) return dispatch_dyn<float>(p, n, Cin, Cout, num_kernels, stream);
    cds_set_error("cds_dynamic_conv: unknown dtype %d", dtype);
    return CDS_EARG;
}

int cds_conv2d_3x3s2(const void* in, const double* in_stats, int in_act, const float* wgt, int n, int Cin, int Cout, int H,
                     int W, int dtype, void* out, void* out_lo, double* out_stats, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt && out, CDS_EARG, "cds_conv2d_3x3s2: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && H > 0 && W > 0, CDS_ESHAPE, "cds_conv2d_3x3s2: bad shape");
    int Ho = (H + 1) / 2, Wo = (W + 1) / 2;
    dim3 grid(cds_div_up((long long)Ho * Wo, 128 * kChunks), n);
#define CDS_GO(T, ci, co)                                                                                                   \
    {                                                                                                                        \
        size_t smem = sizeof(float) * (9 * ci * co + ci * 2 + 4 * co * 2);                                                   \
        cudaFuncSetAttribute(conv3x3s2_kernel<T, ci, co>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);          \
        conv3x3s2_kernel<T, ci, co><<<grid, 128, smem, stream>>>((const T*)in, in_stats, in_act, wgt, H, W, (T*)out, (T*)out_lo, out_stats); \
        return cds_check_launch("cds_conv2d_3x3s2");                                                                        \
    }
    if (dtype == CDS_F16) {
        if (Cin == 8 && Cout == 16) CDS_GO(__half, 8, 16)
        if (Cin == 16 && Cout == 32) CDS_GO(__half, 16, 32)
    } else if (dtype == CDS_F32) {
        if (Cin == 8 && Cout == 16) CDS_GO(float, 8, 16)
        if (Cin == 16 && Cout == 32) CDS_GO(float, 16, 32)
    }
#undef CDS_GO
    cds_set_error("cds_conv2d_3x3s2: unsupported (Cin=%d, Cout=%d, dtype=%d)", Cin, Cout, dtype);
    return CDS_EUNSUPPORTED;
}

int cds_conv2d_1x1_cat(const void* a, const double* a_stats, int a_act, const void* b, const double* b_stats, int b_act,
                       const float* wgt, int n, int Ca, int Cb, int Cout, int H, int W, int dtype, void* out,
                       double* out_stats, cudaStream_t stream) {
    CDS_REQUIRE(a && b && wgt && out, CDS_EARG, "cds_conv2d_1x1_cat: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, CDS_ESHAPE, "cds_conv2d_1x1_cat: bad shape");
    dim3 grid(cds_div_up((long long)(H / 2) * (W / 2), 128 * kQuadChunks), n);
#define CDS_GO(T, ca, cb, co)                                                                                                  \
    {                                                                                                                           \
        size_t smem = sizeof(float) * ((ca + cb) * co + (ca + cb) * 2 + 4 * co * 2);                                            \
        conv1x1_cat_quad_kernel<T, ca, cb, co><<<grid, 128, smem, stream>>>((const T*)a, a_stats, a_act, (const T*)b, b_stats, \
                                                                            b_act, wgt, H, W, (T*)out, out_stats);             \
        return cds_check_launch("cds_conv2d_1x1_cat");                                                                         \
    }
    if (dtype == CDS_F16) {
        if (Ca == 32 && Cb == 16 && Cout == 16) CDS_GO(__half, 32, 16, 16)
        if (Ca == 16 && Cb == 8 && Cout == 8) CDS_GO(__half, 16, 8, 8)
    } else if (dtype == CDS_F32) {
        if (Ca == 32 && Cb == 16 && Cout == 16) CDS_GO(float, 32, 16, 16)
        if (Ca == 16 && Cb == 8 && Cout == 8) CDS_GO(float, 16, 8, 8)
    }
#undef CDS_GO
    cds_set_error("cds_conv2d_1x1_cat: unsupported (Ca=%d, Cb=%d, Cout=%d, dtype=%d)", Ca, Cb, Cout, dtype);
    return CDS_EUNSUPPORTED;
}

int cds_instnorm_act(const void* raw, const double* stats, int act, int n, int C, int H, int W, int dtype, void* out,
                     cudaStream_t stream) {
    CDS_REQUIRE(raw && stats && out, CDS_EARG, "cds_instnorm_act: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && C % 8 == 0 && C > 0, CDS_ESHAPE, "cds_instnorm_act: bad shape");
    long long HW = (long long)H * W;
    dim3 grid((unsigned)std::min<long long>(148 * 8, (HW * (C / 8) + 255) / 256), n);
    if (dtype == CDS_F16)
        instnorm_act_kernel<__half><<<grid, 256, C * 2 * sizeof(float), stream>>>((const __half*)raw, stats, act, C, HW, (__half*)out);
    else if (dtype == CDS_F32)
        instnorm_act_kernel<float><<<grid, 256, C * 2 * sizeof(float), stream>>>((const float*)raw, stats, act, C, HW, (float*)out);
    else { cds_set_error("cds_instnorm_act: unknown dtype %d", dtype); return CDS_EARG; }
    return cds_check_launch("cds_instnorm_act");
}

int cds_instnorm_act_split_f32(const void* raw, const void* raw_lo, const double* stats, int act, int n, int C, int H, int W,
                               float* out, void* out_f16, cudaStream_t stream) {
    CDS_REQUIRE(raw && stats && out, CDS_EARG, "cds_instnorm_act_split_f32: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && C % 8 == 0 && C > 0, CDS_ESHAPE, "cds_instnorm_act_split_f32: bad shape");
    long long HW = (long long)H * W;
    dim3 grid((unsigned)std::min<long long>(148 * 8, (HW * (C / 8) + 255) / 256), n);
    instnorm_act_split_kernel<<<grid, 256, C * 2 * sizeof(float), stream>>>((const __half*)raw, (const __half*)raw_lo, stats, act, C, HW, out,
                                                                            (__half*)out_f16);
    return cds_check_launch("cds_instnorm_act_split_f32");
}

int cds_image_u8_to_f32(const unsigned char* in, long long count, float* out, cudaStream_t stream) {
    CDS_REQUIRE(in && out && count > 0, CDS_EARG, "cds_image_u8_to_f32: bad arguments");
    CDS_REQUIRE(((uintptr_t)in & 3) == 0 && ((uintptr_t)out & 15) == 0, CDS_EARG, "cds_image_u8_to_f32: in must be 4-byte, out 16-byte aligned");
    const long long n4 = count / 4;
    if (n4 > 0) {
        const unsigned grid = (unsigned)std::min<long long>(148 * 16, (n4 + 255) / 256);
        u8_to_f32_kernel<<<grid, 256, 0, stream>>>((const uchar4*)in, n4, (float4*)out);
    }
    if (count % 4) u8_to_f32_tail_kernel<<<1, 4, 0, stream>>>(in, n4 * 4, count, out);
    return cds_check_launch("cds_image_u8_to_f32");
}

int cds_nchw_to_nhwc(const float* in, int n, int C, int H, int W, int dtype, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && out && n > 0 && n <= 65535, CDS_EARG, "cds_nchw_to_nhwc: bad arguments");
    long long HW = (long long)H * W;
    dim3 grid((unsigned)std::min<long long>(148 * 8, (HW * C + 255) / 256), n);
    if (dtype == CDS_F16) nchw_to_nhwc_kernel<__half><<<grid, 256, 0, stream>>>(in, C, HW, (__half*)out);
    else if (dtype == CDS_F32) nchw_to_nhwc_kernel<float><<<grid, 256, 0, stream>>>(in, C, HW, (float*)out);
    else { cds_set_error("cds_nchw_to_nhwc: unknown dtype %d", dtype); return CDS_EARG; }
    return cds_check_launch("cds_nchw_to_nhwc");
}

int cds_nhwc_to_nchw(const void* in, int n, int C, int H, int W, int dtype, float* out, cudaStream_t stream) {
    CDS_REQUIRE(in && out && n > 0 && n <= 65535, CDS_EARG, "cds_nhwc_to_nchw: bad arguments");
    long long HW = (long long)H * W;
    dim3 grid((unsigned)std::min<long long>(148 * 8, (HW * C + 255) / 256), n);
    if (dtype == CDS_F16) nhwc_to_nchw_kernel<__half><<<grid, 256, 0, stream>>>((const __half*)in, C, HW, out);
    else if (dtype == CDS_F32) nhwc_to_nchw_kernel<float><<<grid, 256, 0, stream>>>((const float*)in, C, HW, out);
    else { cds_set_error("cds_nhwc_to_nchw: unknown dtype %d", dtype); return CDS_EARG; }
    return cds_check_launch("cds_nhwc_to_nchw");
}

}  // extern "C"
