// 3-D regulariser building blocks (A4) and the soft-argmin tail (A5), CUDA-core direct form.
// These are the straightforward fp32-accumulate kernels: the numerical anchor on the GPU for the
// tensor-core (tcgen05) implicit-GEMM kernels in conv3d_tc.cu, and the fallback for layer shapes
// the tensor-core path does not cover.
//
// Reference: models/module.py:80-166 (Conv3d / Deconv3d blocks: conv -> BatchNorm3d -> ReLU),
// :270-315 (CostRegNet wiring, skip added AFTER the deconv's ReLU), :373-391 (regression).
// BatchNorm (eval, running stats) is folded by the host: weights arrive as [27][Cin][Cout] fp32
// already scaled, plus a per-channel bias.  Activations are channel-BLOCKED channels-last:
// [B][C/8][D][H][W][8] (for C = 8 this is plain NDHWC), so that one 8-channel slab of a row of voxels is
// contiguous -- the unit the TMA / tcgen05 kernels in conv3d_tc.cu consume.
#include "cds_common.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// conv3d k=3, pad=1, stride 1|2, + bias (+ReLU)
// ---------------------------------------------------------------------------------------------
template <typename T, int CIN, int COUT, int VPT>
__global__ void __launch_bounds__(128) conv3d_kernel(const T* __restrict__ in, const float* __restrict__ wgt,
                                                     const float* __restrict__ bias, int B, int Di, int Hi, int Wi, int Do,
                                                     int Ho, int Wo, int stride, int relu, T* __restrict__ out) {
    __shared__ __align__(16) float w_s[CIN * COUT];
    const long long M = (long long)B * Do * Ho * Wo;
    long long base = (long long)blockIdx.x * (128 * VPT) + threadIdx.x;
    int od[VPT], oh[VPT], ow[VPT], ob[VPT];
    bool live[VPT];
    float acc[VPT][COUT];
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        long long m = base + (long long)v * 128;
        live[v] = m < M;
        long long mm = live[v] ? m : 0;
        ow[v] = (int)(mm % Wo);
        oh[v] = (int)((mm / Wo) % Ho);
        od[v] = (int)((mm / ((long long)Wo * Ho)) % Do);
        ob[v] = (int)(mm / ((long long)Wo * Ho * Do));
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[v][c] = 0.f;
    }
    for (int tap = 0; tap < 27; ++tap) {
        __syncthreads();
        for (int i = threadIdx.x; i < CIN * COUT / 4; i += 128)
            reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(wgt + (size_t)tap * CIN * COUT) + i);
        __syncthreads();
        int kd = tap / 9, kh = (tap / 3) % 3, kw = tap % 3;
#pragma unroll
        for (int v = 0; v < VPT; ++v) {
            int id = od[v] * stride - 1 + kd, ih = oh[v] * stride - 1 + kh, iw = ow[v] * stride - 1 + kw;
            if (!live[v] || id < 0 || id >= Di || ih < 0 || ih >= Hi || iw < 0 || iw >= Wi) continue;
            const size_t Min = (size_t)Di * Hi * Wi;
            const T* ip = in + ((size_t)ob[v] * (CIN / 8) * Min + ((size_t)id * Hi + ih) * Wi + iw) * 8;
#pragma unroll
            for (int c8 = 0; c8 < CIN / 8; ++c8) {
                float x[8];
                Vec8<T>::load(ip + (size_t)c8 * Min * 8, x);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4* wp = reinterpret_cast<const float4*>(w_s + (c8 * 8 + j) * COUT);
#pragma unroll
                    for (int q = 0; q < COUT / 4; ++q) {
                        float4 ww = wp[q];
                        acc[v][4 * q + 0] += x[j] * ww.x;
                        acc[v][4 * q + 1] += x[j] * ww.y;
                        acc[v][4 * q + 2] += x[j] * ww.z;
                        acc[v][4 * q + 3] += x[j] * ww.w;
                    }
                }
            }
        }
    }
#pragma unroll
    for (int v = 0; v < VPT; ++v) {
        if (!live[v]) continue;
        const size_t Mout = (size_t)Do * Ho * Wo;
        const size_t ov = ((size_t)od[v] * Ho + oh[v]) * Wo + ow[v];
        T* op = out + ((size_t)ob[v] * (COUT / 8) * Mout + ov) * 8;
#pragma unroll
        for (int c8 = 0; c8 < COUT / 8; ++c8) {
            float y[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float t = acc[v][c8 * 8 + j] + __ldg(bias + c8 * 8 + j);
                y[j] = relu ? fmaxf(t, 0.f) : t;
            }
            Vec8<T>::store(op + (size_t)c8 * Mout * 8, y);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// ConvTranspose3d k=3, stride 2, pad 1, output_padding 1 (+bias, ReLU, then + skip)
// out[2i - 1 + k] += in[i] * w[k] per axis: even outputs take tap k=1 from i=o/2, odd outputs take
// k=2 from i=(o-1)/2 and k=0 from i=(o+1)/2.  A thread owns one input-aligned 2x2x2 output cell and
// walks its 8 parity classes (1,2,2,2,4,4,4,8 taps) so the whole warp runs the same tap list.
// ---------------------------------------------------------------------------------------------
template <typename T, int CIN, int COUT>
__global__ void __launch_bounds__(128) deconv3d_kernel(const T* __restrict__ in, const float* __restrict__ wgt,
                                                       const float* __restrict__ bias, const T* __restrict__ skip, int B,
                                                       int Di, int Hi, int Wi, T* __restrict__ out) {
    __shared__ __align__(16) float w_s[CIN * COUT];
    const long long M = (long long)B * Di * Hi * Wi;
    long long m = (long long)blockIdx.x * 128 + threadIdx.x;
    bool live = m < M;
    long long mm = live ? m : 0;
    int iw = (int)(mm % Wi), ih = (int)((mm / Wi) % Hi), id = (int)((mm / ((long long)Wi * Hi)) % Di);
    int b = (int)(mm / ((long long)Wi * Hi * Di));
    const int Do = 2 * Di, Ho = 2 * Hi, Wo = 2 * Wi;
    for (int par = 0; par < 8; ++par) {
        int pd = par >> 2, ph = (par >> 1) & 1, pw = par & 1;
        float acc[COUT];
#pragma unroll
        for (int c = 0; c < COUT; ++c) acc[c] = 0.f;
        int nd = pd ? 2 : 1, nh = ph ? 2 : 1, nw = pw ? 2 : 1;
        for (int t = 0; t < nd * nh * nw; ++t) {
            int sd = t / (nh * nw), sh = (t / nw) % nh, sw = t % nw;  // 0: same input voxel, 1: +1 neighbour
            int kd = pd ? (sd ? 0 : 2) : 1, kh = ph ? (sh ? 0 : 2) : 1, kw = pw ? (sw ? 0 : 2) : 1;
            int tap = (kd * 3 + kh) * 3 + kw;
            __syncthreads();
            for (int i = threadIdx.x; i < CIN * COUT / 4; i += 128)
                reinterpret_cast<float4*>(w_s)[i] = __ldg(reinterpret_cast<const float4*>(wgt + (size_t)tap * CIN * COUT) + i);
            __syncthreads();
            int jd = id + sd, jh = ih + sh, jw = iw + sw;
            if (!live || jd >= Di || jh >= Hi || jw >= Wi) continue;
            const size_t Min = (size_t)Di * Hi * Wi;
            const T* ip = in + ((size_t)b * (CIN / 8) * Min + ((size_t)jd * Hi + jh) * Wi + jw) * 8;
#pragma unroll
            for (int c8 = 0; c8 < CIN / 8; ++c8) {
                float x[8];
                Vec8<T>::load(ip + (size_t)c8 * Min * 8, x);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4* wp = reinterpret_cast<const float4*>(w_s + (c8 * 8 + j) * COUT);
#pragma unroll
                    for (int q = 0; q < COUT / 4; ++q) {
                        float4 ww = wp[q];
                        acc[4 * q + 0] += x[j] * ww.x;
                        acc[4 * q + 1] += x[j] * ww.y;
                        acc[4 * q + 2] += x[j] * ww.z;
                        acc[4 * q + 3] += x[j] * ww.w;
                    }
                }
            }
        }
        if (live) {
            const size_t Mout = (size_t)Do * Ho * Wo;
            size_t o = ((size_t)b * (COUT / 8) * Mout + ((size_t)(2 * id + pd) * Ho + 2 * ih + ph) * Wo + 2 * iw + pw) * 8;
#pragma unroll
            for (int c8 = 0; c8 < COUT / 8; ++c8) {
                float y[8], s[8];
                if (skip) Vec8<T>::load(skip + o + (size_t)c8 * Mout * 8, s);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float t = fmaxf(acc[c8 * 8 + j] + __ldg(bias + c8 * 8 + j), 0.f);
                    y[j] = skip ? s[j] + t : t;
                }
                Vec8<T>::store(out + o + (size_t)c8 * Mout * 8, y);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// prob head: plain Conv3d C -> 1, k3 p1, no bias / BN / ReLU (module.py:303); fp32 logits [B,D,h,w]
// ---------------------------------------------------------------------------------------------
template <typename T, int CIN>
__global__ void __launch_bounds__(256) prob_conv_kernel(const T* __restrict__ in, const float* __restrict__ wgt, int B,
                                                        int D, int H, int W, float* __restrict__ logits) {
    __shared__ float w_s[27 * CIN];
    for (int i = threadIdx.x; i < 27 * CIN; i += blockDim.x) w_s[i] = __ldg(wgt + i);
    __syncthreads();
    const long long M = (long long)B * D * H * W;
    long long m = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    int x = (int)(m % W), y = (int)((m / W) % H), d = (int)((m / ((long long)W * H)) % D);
    int b = (int)(m / ((long long)W * H * D));
    float acc = 0.f;
    for (int kd = 0; kd < 3; ++kd) {
        int zd = d - 1 + kd;
        if (zd < 0 || zd >= D) continue;
        for (int kh = 0; kh < 3; ++kh) {
            int zy = y - 1 + kh;
            if (zy < 0 || zy >= H) continue;
#pragma unroll
            for (int kw = 0; kw < 3; ++kw) {
                int zx = x - 1 + kw;
                if (zx < 0 || zx >= W) continue;
                const T* ip = in + ((((size_t)b * D + zd) * H + zy) * W + zx) * CIN;
                const float* wp = w_s + ((kd * 3 + kh) * 3 + kw) * CIN;
#pragma unroll
                for (int c8 = 0; c8 < CIN / 8; ++c8) {
                    float v[8];
                    Vec8<T>::load(ip + c8 * 8, v);
#pragma unroll
                    for (int j = 0; j < 8; ++j) acc += v[j] * wp[c8 * 8 + j];
                }
            }
        }
    }
    logits[m] = acc;
}

// ---------------------------------------------------------------------------------------------
// softmax over D + expectation depth + 4-plane confidence window (model.py:90-92, module.py:373-391)
// one thread per pixel; logits / probabilities are [B, D, h, w] fp32 (plane-major => coalesced)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_regress_kernel(const float* __restrict__ logits, const float* __restrict__ depth,
                                                              int per_pixel, int is_prob, int B, int D, long long P,
                                                              float* __restrict__ depth_out, float* __restrict__ conf_out,
                                                              float* __restrict__ prob_out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * P) return;
    int b = (int)(i / P);
    long long p = i % P;
    const float* lp = logits + (size_t)b * D * P + p;
    const float* dpp = per_pixel ? depth + (size_t)b * D * P + p : depth + (size_t)b * D;
    const size_t dstride = per_pixel ? (size_t)P : 1;
    float m = 0.f, inv = 1.f, ed = 0.f, ei = 0.f;
    if (!is_prob && !prob_out) {
        // one pass, online softmax: running max m, S = sum e^(l-m), and the two expectations rescaled together
        float S = 0.f;
        m = -INFINITY;
        for (int d = 0; d < D; ++d) {
            const float l = __ldg(lp + (size_t)d * P);
            const float dep = __ldg(dpp + (size_t)d * dstride);
            if (l > m) {
                const float r = __expf(m - l);   // 0 on the first plane (m = -inf)
                S *= r; ed *= r; ei *= r;
                m = l;
            }
            const float e = __expf(l - m);
            S += e;
            ed += e * dep;
            ei += e * (float)d;
        }
        inv = 1.f / S;
        ed *= inv;
        ei *= inv;
    } else {
        float S = 1.f;
        if (!is_prob) {
            m = -INFINITY;
            for (int d = 0; d < D; ++d) m = fmaxf(m, __ldg(lp + (size_t)d * P));
            S = 0.f;
            for (int d = 0; d < D; ++d) S += __expf(__ldg(lp + (size_t)d * P) - m);
        }
        inv = 1.f / S;
        for (int d = 0; d < D; ++d) {
            float l = __ldg(lp + (size_t)d * P);
            float pr = is_prob ? l : __expf(l - m) * inv;
            float dep = __ldg(dpp + (size_t)d * dstride);
            ed += pr * dep;
            ei += pr * (float)d;
            if (prob_out) prob_out[(size_t)b * D * P + (size_t)d * P + p] = pr;
        }
    }
    if (depth_out) depth_out[i] = ed;
    if (conf_out) {
        int idx = min(max((int)ei, 0), D - 1);  // .long() truncates toward zero
        float c = 0.f;
        for (int j = idx - 1; j <= idx + 2; ++j)
            if (j >= 0 && j < D) {
                float l = __ldg(lp + (size_t)j * P);
                c += is_prob ? l : __expf(l - m) * inv;
            }
        conf_out[i] = c;
    }
}

// ---------------------------------------------------------------------------------------------
template <typename T, int CIN, int COUT>
int launch_conv(const void* in, const float* w, const float* bias, int B, int Di, int Hi, int Wi, int stride, int relu,
                void* out, cudaStream_t st) {
    constexpr int VPT = COUT <= 8 ? 4 : (COUT <= 16 ? 2 : 1);
    int Do = (Di + stride - 1) / stride, Ho = (Hi + stride - 1) / stride, Wo = (Wi + stride - 1) / stride;
    long long M = (long long)B * Do * Ho * Wo;
    conv3d_kernel<T, CIN, COUT, VPT><<<cds_div_up(M, 128 * VPT), 128, 0, st>>>((const T*)in, w, bias, B, Di, Hi, Wi, Do, Ho,
                                                                             Wo, stride, relu, (T*)out);
    return cds_check_launch("cds_conv3d_k3");
}

template <typename T>
int dispatch_conv(const void* in, const float* w, const float* bias, int B, int Cin, int Cout, int Di, int Hi, int Wi,
                  int stride, int relu, void* out, cudaStream_t st) {
#define CDS_CASE(ci, co) \
    if (Cin == ci && Cout == co) return launch_conv<T, ci, co>(in, w, bias, B, Di, Hi, Wi, stride, relu, out, st);
    CDS_CASE(8, 8) CDS_CASE(16, 8) CDS_CASE(32, 8) CDS_CASE(8, 16) CDS_CASE(16, 16) CDS_CASE(16, 32) CDS_CASE(32, 32)
    CDS_CASE(32, 64) CDS_CASE(64, 64)
#undef CDS_CASE
    cds_set_error("cds_conv3d_k3: unsupported channel pair Cin=%d Cout=%d", Cin, Cout);
    return CDS_EUNSUPPORTED;
}

template <typename T, int CIN, int COUT>
int launch_deconv(const void* in, const float* w, const float* bias, const void* skip, int B, int Di, int Hi, int Wi,
                  void* out, cudaStream_t st) {
    long long M = (long long)B * Di * Hi * Wi;
    deconv3d_kernel<T, CIN, COUT><<<cds_div_up(M, 128), 128, 0, st>>>((const T*)in, w, bias, (const T*)skip, B, Di, Hi, Wi, (T*)out);
    return cds_check_launch("cds_deconv3d_k3s2");
}

template <typename T>
int dispatch_deconv(const void* in, const float* w, const float* bias, const void* skip, int B, int Cin, int Cout, int Di,
                    int Hi, int Wi, void* out, cudaStream_t st) {
    if (Cin == 64 && Cout == 32) return launch_deconv<T, 64, 32>(in, w, bias, skip, B, Di, Hi, Wi, out, st);
    if (Cin == 32 && Cout == 16) return launch_deconv<T, 32, 16>(in, w, bias, skip, B, Di, Hi, Wi, out, st);
    if (Cin == 16 && Cout == 8) return launch_deconv<T, 16, 8>(in, w, bias, skip, B, Di, Hi, Wi, out, st);
    cds_set_error("cds_deconv3d_k3s2: unsupported channel pair Cin=%d Cout=%d", Cin, Cout);
    return CDS_EUNSUPPORTED;
}

}  // namespace

extern "C" {

int cds_conv3d_k3(const void* in, const float* wgt, const float* bias, int B, int Cin, int Cout, int D, int H, int W,
                  int stride, int relu, int dtype, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt && bias && out, CDS_EARG, "cds_conv3d_k3: null pointer");
    CDS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0 && (stride == 1 || stride == 2), CDS_ESHAPE,
                "cds_conv3d_k3: bad shape B=%d D=%d H=%d W=%d stride=%d", B, D, H, W, stride);
    if (dtype == CDS_F16) return dispatch_conv<__half>(in, wgt, bias, B, Cin, Cout, D, H, W, stride, relu, out, stream);
    if (dtype == CDS_F32) return dispatch_conv<float>(in, wgt, bias, B, Cin, Cout, D, H, W, stride, relu, out, stream);
    cds_set_error("cds_conv3d_k3: unknown dtype %d", dtype);
    return CDS_EARG;
}

int cds_deconv3d_k3s2(const void* in, const float* wgt, const float* bias, const void* skip, int B, int Cin, int Cout,
                      int D, int H, int W, int dtype, void* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt && bias && out, CDS_EARG, "cds_deconv3d_k3s2: null pointer");
    CDS_REQUIRE(B > 0 && D > 0 && H > 0 && W > 0, CDS_ESHAPE, "cds_deconv3d_k3s2: bad shape");
    if (dtype == CDS_F16) return dispatch_deconv<__half>(in, wgt, bias, skip, B, Cin, Cout, D, H, W, out, stream);
    if (dtype == CDS_F32) return dispatch_deconv<float>(in, wgt, bias, skip, B, Cin, Cout, D, H, W, out, stream);
    cds_set_error("cds_deconv3d_k3s2: unknown dtype %d", dtype);
    return CDS_EARG;
}

int cds_prob_conv(const void* in, const float* wgt, int B, int Cin, int D, int H, int W, int dtype, float* logits,
                  cudaStream_t stream) {
    CDS_REQUIRE(in && wgt && logits, CDS_EARG, "cds_prob_conv: null pointer");
    CDS_REQUIRE(Cin == 8, CDS_EUNSUPPORTED, "cds_prob_conv: Cin must be 8 (got %d)", Cin);
    long long M = (long long)B * D * H * W;
    if (dtype == CDS_F16)
        prob_conv_kernel<__half, 8><<<cds_div_up(M, 256), 256, 0, stream>>>((const __half*)in, wgt, B, D, H, W, logits);
    else if (dtype == CDS_F32)
        prob_conv_kernel<float, 8><<<cds_div_up(M, 256), 256, 0, stream>>>((const float*)in, wgt, B, D, H, W, logits);
    else { cds_set_error("cds_prob_conv: unknown dtype %d", dtype); return CDS_EARG; }
    return cds_check_launch("cds_prob_conv");
}

int cds_softmax_regress(const float* logits, const float* depth, int depth_per_pixel, int input_is_prob, int B, int D,
                        int h, int w, float* depth_out, float* conf_out, float* prob_out, cudaStream_t stream) {
    CDS_REQUIRE(logits && (depth || !depth_out), CDS_EARG, "cds_softmax_regress: null pointer");
    CDS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, CDS_ESHAPE, "cds_softmax_regress: bad shape");
    long long P = (long long)h * w;
    const float* dep = depth ? depth : logits;  // unused when depth_out is null
    softmax_regress_kernel<<<cds_div_up(B * P, 256), 256, 0, stream>>>(logits, dep, depth ? depth_per_pixel : 1, input_is_prob,
                                                                       B, D, P, depth_out, conf_out, prob_out);
    return cds_check_launch("cds_softmax_regress");
}

}  // extern "C"
