// Depth-map refinement network: the reference's `Refinement` (models/module.py:318-370), the last step of
// CDSMVSNet.forward when refine=True (models/model.py:209-216) -- the configuration of all three pretrained checkpoints.
//
//   depth_n = (depth_0 - lo) / (hi - lo) * 10                                   [B,1,h,w]   (h = H/2)
//   conv0   = ConvBnReLU(3 -> 8)(img)                                            [B,8,H,W]
//   deconv  = ReLU(BN(ConvTranspose2d(8 -> 8, k3 s2 p1 op1)(ConvBnReLU(8->8)(ConvBnReLU(1->8)(depth_n)))))   [B,8,H,W]
//   res     = Conv2d(8 -> 1, k3 p1)(ConvBnReLU(16 -> 8)(cat(deconv, conv0)))     [B,1,H,W]
//   out     = ((bilinear_up2(depth_n, align_corners=True) + res) / 10) * (hi - lo) + lo
// Everything is fp32 planar NCHW (this net sees the raw image and produces the final depth; it is ~3.5 kFLOP per pixel, far
// from any roofline that matters next to the cascade), BatchNorm folded by the host.  Four small kernels, each one thread
// per output pixel with the folded weights in shared memory.
#include "cds_common.cuh"

namespace {

__global__ void refine_prescale_kernel(const float* __restrict__ depth0, const float* __restrict__ lo, const float* __restrict__ hi,
                                       long long hw, float* __restrict__ out) {
    const int b = blockIdx.y;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= hw) return;
    const float l = __ldg(lo + b), h = __ldg(hi + b);
    out[(size_t)b * hw + i] = (__ldg(depth0 + (size_t)b * hw + i) - l) / (h - l) * 10.f;
}

// 3x3 pad-1 conv over cat(a [B,CA,H,W], b [B,CB,H,W]) -> [B,COUT,H,W], + bias (folded BN), optional ReLU
template <int CA, int CB, int COUT>
__global__ void __launch_bounds__(256) conv3x3_f32_kernel(const float* __restrict__ a, const float* __restrict__ bsrc,
                                                          const float* __restrict__ wgt /*[COUT][CA+CB][9]*/, const float* __restrict__ bias,
                                                          int H, int W, int relu, float* __restrict__ out) {
    constexpr int CIN = CA + CB;
    __shared__ float s_w[COUT * CIN * 9];
    __shared__ float s_b[COUT];
    for (int i = threadIdx.x; i < COUT * CIN * 9; i += blockDim.x) s_w[i] = __ldg(wgt + i);
    if (threadIdx.x < COUT) s_b[threadIdx.x] = bias ? __ldg(bias + threadIdx.x) : 0.f;
    __syncthreads();
    const int n = blockIdx.y;
    const int P = H * W;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const int x = pix % W, y = pix / W;
    float acc[COUT];
#pragma unroll
    for (int c = 0; c < COUT; ++c) acc[c] = s_b[c];
    for (int ci = 0; ci < CIN; ++ci) {
        const float* src = ci < CA ? a + ((size_t)n * CA + ci) * P : bsrc + ((size_t)n * CB + (ci - CA)) * P;
        float v[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int yy = y - 1 + k / 3, xx = x - 1 + k % 3;
            v[k] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(src + (size_t)yy * W + xx) : 0.f;
        }
#pragma unroll
        for (int c = 0; c < COUT; ++c) {
            const float* w9 = s_w + (c * CIN + ci) * 9;
#pragma unroll
            for (int k = 0; k < 9; ++k) acc[c] += v[k] * w9[k];
        }
    }
#pragma unroll
    for (int c = 0; c < COUT; ++c) out[((size_t)n * COUT + c) * P + pix] = relu ? fmaxf(acc[c], 0.f) : acc[c];
}

// ConvTranspose2d(C -> C, k3, stride 2, pad 1, output_padding 1) + folded BN + ReLU: out[2i - 1 + k] += in[i] * w[k] per axis
template <int C>
__global__ void __launch_bounds__(256) deconv3x3s2_f32_kernel(const float* __restrict__ in, const float* __restrict__ wgt /*[Cin][Cout][9] folded*/,
                                                              const float* __restrict__ bias, int h, int w, float* __restrict__ out) {
    __shared__ float s_w[C * C * 9];
    __shared__ float s_b[C];
    for (int i = threadIdx.x; i < C * C * 9; i += blockDim.x) s_w[i] = __ldg(wgt + i);
    if (threadIdx.x < C) s_b[threadIdx.x] = __ldg(bias + threadIdx.x);
    __syncthreads();
    const int n = blockIdx.y, H = 2 * h, W = 2 * w;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= H * W) return;
    const int ox = pix % W, oy = pix / W;
    float acc[C];
#pragma unroll
    for (int c = 0; c < C; ++c) acc[c] = s_b[c];
    for (int ky = 0; ky < 3; ++ky) {
        const int ty = oy + 1 - ky;
        if (ty < 0 || (ty & 1) || ty / 2 >= h) continue;
        for (int kx = 0; kx < 3; ++kx) {
            const int tx = ox + 1 - kx;
            if (tx < 0 || (tx & 1) || tx / 2 >= w) continue;
            const size_t ip = (size_t)(ty / 2) * w + tx / 2;
#pragma unroll
            for (int ci = 0; ci < C; ++ci) {
                const float v = __ldg(in + ((size_t)n * C + ci) * h * w + ip);
#pragma unroll
                for (int co = 0; co < C; ++co) acc[co] += v * s_w[(ci * C + co) * 9 + ky * 3 + kx];
            }
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) out[((size_t)n * C + c) * H * W + pix] = fmaxf(acc[c], 0.f);
}

// res = Conv2d(8 -> 1)(x); out = ((up2(depth_n) + res) / 10 * (hi - lo) + lo) * post
__global__ void __launch_bounds__(256) refine_final_kernel(const float* __restrict__ x /*[B,8,H,W]*/, const float* __restrict__ wgt /*[8][9]*/,
                                                           const float* __restrict__ depth_n /*[B,h,w]*/, const float* __restrict__ lo,
                                                           const float* __restrict__ hi, const float* __restrict__ post, int h, int w,
                                                           float* __restrict__ out /*[B,H,W]*/) {
    __shared__ float s_w[72];
    if (threadIdx.x < 72) s_w[threadIdx.x] = __ldg(wgt + threadIdx.x);
    __syncthreads();
    const int n = blockIdx.y, H = 2 * h, W = 2 * w, P = H * W;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const int ox = pix % W, oy = pix / W;
    float res = 0.f;
    for (int ci = 0; ci < 8; ++ci) {
        const float* src = x + ((size_t)n * 8 + ci) * P;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int yy = oy - 1 + k / 3, xx = ox - 1 + k % 3;
            if (yy >= 0 && yy < H && xx >= 0 && xx < W) res += __ldg(src + (size_t)yy * W + xx) * s_w[ci * 9 + k];
        }
    }
    // F.interpolate(scale_factor=2, mode="bilinear", align_corners=True): src = dst * (in - 1) / (out - 1)
    const float sy = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f, sx = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const float fy = sy * (float)oy, fx = sx * (float)ox;
    const int y0 = (int)fy, x0 = (int)fx;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float ly = fy - (float)y0, lx = fx - (float)x0;
    const float* d = depth_n + (size_t)n * h * w;
    const float up = (1.f - ly) * ((1.f - lx) * __ldg(d + (size_t)y0 * w + x0) + lx * __ldg(d + (size_t)y0 * w + x1)) +
                     ly * ((1.f - lx) * __ldg(d + (size_t)y1 * w + x0) + lx * __ldg(d + (size_t)y1 * w + x1));
    const float l = __ldg(lo + n), hh = __ldg(hi + n);
    float v = (up + res) / 10.f * (hh - l) + l;
    if (post) v *= __ldg(post + n);
    out[(size_t)n * P + pix] = v;
}

}  // namespace

extern "C" {

// depth_n [B,h,w] = (depth_0 - lo[b]) / (hi[b] - lo[b]) * 10     (module.py:350-352)
int cds_refine_prescale(const float* depth0, const float* lo, const float* hi, int B, int h, int w, float* depth_n, cudaStream_t stream) {
    CDS_REQUIRE(depth0 && lo && hi && depth_n, CDS_EARG, "cds_refine_prescale: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0, CDS_ESHAPE, "cds_refine_prescale: bad shape");
    const long long hw = (long long)h * w;
    refine_prescale_kernel<<<dim3(cds_div_up(hw, 256), B), 256, 0, stream>>>(depth0, lo, hi, hw, depth_n);
    return cds_check_launch("cds_refine_prescale");
}

// fp32 planar 3x3 pad-1 conv over cat(a [B,Ca,H,W], b [B,Cb,H,W]) -> out [B,Cout,H,W]; wgt [Cout][Ca+Cb][3][3], bias [Cout] or NULL.
// Channel triples of the refinement net: (3,0,8) conv0, (1,0,8) conv1, (8,0,8) conv2, (8,8,8) conv3.
int cds_conv2d_3x3_f32(const float* a, const float* b, const float* wgt, const float* bias, int B, int Ca, int Cb, int Cout, int H, int W,
                       int relu, float* out, cudaStream_t stream) {
    CDS_REQUIRE(a && wgt && out && (b || Cb == 0), CDS_EARG, "cds_conv2d_3x3_f32: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && H > 0 && W > 0 && (long long)H * W < (1ll << 31), CDS_ESHAPE, "cds_conv2d_3x3_f32: bad shape");
    dim3 grid(cds_div_up((long long)H * W, 256), B);
#define CDS_RC(ca, cb, co)                                                                                     \
    if (Ca == ca && Cb == cb && Cout == co) {                                                                  \
        conv3x3_f32_kernel<ca, cb, co><<<grid, 256, 0, stream>>>(a, b, wgt, bias, H, W, relu, out);           \
        return cds_check_launch("cds_conv2d_3x3_f32");                                                         \
    }
    CDS_RC(3, 0, 8) CDS_RC(1, 0, 8) CDS_RC(8, 0, 8) CDS_RC(8, 8, 8)
#undef CDS_RC
    cds_set_error("cds_conv2d_3x3_f32: unsupported channels Ca=%d Cb=%d Cout=%d", Ca, Cb, Cout);
    return CDS_EUNSUPPORTED;
}

// ConvTranspose2d(8 -> 8, k3 s2 p1 op1) + bias (folded BN) + ReLU: in [B,8,h,w] -> out [B,8,2h,2w]; wgt [Cin][Cout][3][3] folded
int cds_deconv2d_k3s2_f32(const float* in, const float* wgt, const float* bias, int B, int C, int h, int w, float* out, cudaStream_t stream) {
    CDS_REQUIRE(in && wgt && bias && out, CDS_EARG, "cds_deconv2d_k3s2_f32: null pointer");
    CDS_REQUIRE(C == 8, CDS_EUNSUPPORTED, "cds_deconv2d_k3s2_f32: C must be 8 (got %d)", C);
    CDS_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0 && 4ll * h * w < (1ll << 31), CDS_ESHAPE, "cds_deconv2d_k3s2_f32: bad shape");
    deconv3x3s2_f32_kernel<8><<<dim3(cds_div_up(4ll * h * w, 256), B), 256, 0, stream>>>(in, wgt, bias, h, w, out);
    return cds_check_launch("cds_deconv2d_k3s2_f32");
}

// out [B,2h,2w] = ((bilinear_up2(depth_n) + Conv2d(8->1)(x)) / 10 * (hi - lo) + lo) * post[b]   (post may be NULL)
int cds_refine_final(const float* x, const float* res_wgt, const float* depth_n, const float* lo, const float* hi, const float* post,
                     int B, int h, int w, float* out, cudaStream_t stream) {
    CDS_REQUIRE(x && res_wgt && depth_n && lo && hi && out, CDS_EARG, "cds_refine_final: null pointer");
    CDS_REQUIRE(B > 0 && B <= 65535 && h > 0 && w > 0 && 4ll * h * w < (1ll << 31), CDS_ESHAPE, "cds_refine_final: bad shape");
    refine_final_kernel<<<dim3(cds_div_up(4ll * h * w, 256), B), 256, 0, stream>>>(x, res_wgt, depth_n, lo, hi, post, h, w, out);
    return cds_check_launch("cds_refine_final");
}

}  // extern "C"
