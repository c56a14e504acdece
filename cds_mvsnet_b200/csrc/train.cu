// Training support for the op-level drop-ins (first slice): the backward passes autograd needs.
//
// Reference semantics restated (file:line into TruongKhang/cds-mvsnet):
//   models/utils/warping.py:79,100-101  the sampling grid is built under no_grad, so homo_warping_3D only propagates
//                                       a gradient to src_fea: the adjoint of the bilinear zero-padded gather
//   models/module.py:373-379            depth = sum_d p_d * depth_d  ->  d/dp = g * depth_d, d/d depth = g * p_d
//   models/losses.py:14-23,36-37        per stage: smooth-L1 (beta 1, mean over mask > 0.5) of depth / interval, and
//                                       the masked mean of norm_curv
//   models/losses.py:25-35              binary cross entropy with logits of feat_distance against feat_target over the
//                                       mask repeated across the planes, positives weighted by neg / pos
#include "cds_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// A1 backward: grad_src[b,c,tap] += weight * grad_out[b,c,d,y,x]
// One thread per (b, d, y, x) voxel: the footprint is computed once (same arithmetic as the forward
// kernel, geometry.cu) and shared by all C channels; consecutive lanes are consecutive x, so the
// grad_out reads are coalesced and neighbouring lanes' reductions land in neighbouring addresses.
// grad_src must be zero on entry (red.global.add.f32).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) homo_warp_backward_kernel(const float* __restrict__ grad_out, const float* __restrict__ coef,
                                                                 const float* __restrict__ depth, int per_pixel, int B, int C,
                                                                 int D, int h, int w, float* __restrict__ grad_src) {
    const long long P = (long long)h * w;
    const long long total = (long long)B * D * P;
    const float half_w = (float)((double)(w - 1) / 2.0), half_h = (float)((double)(h - 1) / 2.0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        const int y = (int)((i / w) % h);
        const int d = (int)((i / P) % D);
        const int b = (int)(i / (P * D));
        const WarpCoef k = load_coef(coef + b * 12);
        const float dep = per_pixel ? __ldg(depth + ((size_t)b * D + d) * P + (size_t)y * w + x) : __ldg(depth + b * D + d);
        float rx, ry, rz, u, v;
        pixel_ray(k, (float)x, (float)y, rx, ry, rz);
        project(k, rx, ry, rz, dep, u, v);
        u = ((u / half_w - 1.f) + 1.f) / 2.f * (float)(w - 1);
        v = ((v / half_h - 1.f) + 1.f) / 2.f * (float)(h - 1);
        const Taps t = make_taps(u, v, w, h);
        if (t.w00 == 0.f && t.w01 == 0.f && t.w10 == 0.f && t.w11 == 0.f) continue;
        const int xa = min(max(t.x0, 0), w - 1), xb = min(max(t.x0 + 1, 0), w - 1);
        const int ya = min(max(t.y0, 0), h - 1), yb = min(max(t.y0 + 1, 0), h - 1);
        const size_t o00 = (size_t)ya * w + xa, o01 = (size_t)ya * w + xb, o10 = (size_t)yb * w + xa, o11 = (size_t)yb * w + xb;
        const float* gb = grad_out + ((size_t)b * C * D + d) * P + (size_t)y * w + x;
        float* sb = grad_src + (size_t)b * C * P;
        float g = __ldg(gb);
        for (int c = 0; c < C; ++c) {
            const float gn = c + 1 < C ? __ldg(gb + (size_t)(c + 1) * D * P) : 0.f;   // next channel in flight under the reductions
            float* sc = sb + (size_t)c * P;
            if (t.w00 != 0.f) atomicAdd(sc + o00, t.w00 * g);
            if (t.w01 != 0.f) atomicAdd(sc + o01, t.w01 * g);
            if (t.w10 != 0.f) atomicAdd(sc + o10, t.w10 * g);
            if (t.w11 != 0.f) atomicAdd(sc + o11, t.w11 * g);
            g = gn;
        }
    }
}

// Channels-last form for C % 4 == 0: the reductions go to a zeroed [B,h,w,C] fp32 workspace as 16-byte vector
// reductions (red.global.add.v4.f32, sm_90+) -- a quarter of the reduction instructions and one L2 sector per tap and
// channel quad instead of four -- and nhwc_to_nchw_kernel then lays the result out as the reference's [B,C,h,w].
// Neighbouring lanes (consecutive x) usually land on overlapping footprints: lane i's right-hand column is lane i+1's
// left-hand column whenever the views' scales are close.  Such a lane hands its right-hand column to its neighbour by
// shuffle and the neighbour reduces the sum, which halves the reductions that reach L2 (the bound of this kernel).
__device__ __forceinline__ float4 scale4(float w, float4 g) { return make_float4(w * g.x, w * g.y, w * g.z, w * g.w); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 shfl_up4(float4 v) {
    return make_float4(__shfl_up_sync(0xffffffffu, v.x, 1), __shfl_up_sync(0xffffffffu, v.y, 1), __shfl_up_sync(0xffffffffu, v.z, 1),
                       __shfl_up_sync(0xffffffffu, v.w, 1));
}

__global__ void __launch_bounds__(256) homo_warp_backward_nhwc_kernel(const float* __restrict__ grad_out, const float* __restrict__ coef,
                                                                      const float* __restrict__ depth, int per_pixel, int B, int C,
                                                                      int D, int h, int w, float* __restrict__ ws) {
    const long long P = (long long)h * w;
    const long long total = (long long)B * D * P;
    const float half_w = (float)((double)(w - 1) / 2.0), half_h = (float)((double)(h - 1) / 2.0);
    const int lane = threadIdx.x & 31;
    // warp-uniform trip count: every lane takes part in the shuffles of every iteration
    for (long long base = blockIdx.x * (long long)blockDim.x; base < total; base += (long long)gridDim.x * blockDim.x) {
        const long long i = base + threadIdx.x;
        const bool live = i < total;
        const long long ii = live ? i : total - 1;
        const int x = (int)(ii % w);
        const int y = (int)((ii / w) % h);
        const int d = (int)((ii / P) % D);
        const int b = (int)(ii / (P * D));
        const WarpCoef k = load_coef(coef + b * 12);
        const float dep = per_pixel ? __ldg(depth + ((size_t)b * D + d) * P + (size_t)y * w + x) : __ldg(depth + b * D + d);
        float rx, ry, rz, u, v;
        pixel_ray(k, (float)x, (float)y, rx, ry, rz);
        project(k, rx, ry, rz, dep, u, v);
        u = ((u / half_w - 1.f) + 1.f) / 2.f * (float)(w - 1);
        v = ((v / half_h - 1.f) + 1.f) / 2.f * (float)(h - 1);
        Taps t = make_taps(u, v, w, h);
        if (!live) t.w00 = t.w01 = t.w10 = t.w11 = 0.f;
        const int xa = min(max(t.x0, 0), w - 1), xb = min(max(t.x0 + 1, 0), w - 1);
        const int ya = min(max(t.y0, 0), h - 1), yb = min(max(t.y0 + 1, 0), h - 1);
        // pixel indices of the four taps inside the whole workspace (batch included, so equal index == equal address)
        const long long pb = (long long)b * P;
        const long long o00 = pb + (long long)ya * w + xa, o01 = pb + (long long)ya * w + xb;
        const long long o10 = pb + (long long)yb * w + xa, o11 = pb + (long long)yb * w + xb;
        // does my left-hand column coincide with the previous lane's right-hand column?
        const long long p01 = __shfl_up_sync(0xffffffffu, o01, 1), p11 = __shfl_up_sync(0xffffffffu, o11, 1);
        const bool take_left = lane > 0 && p01 == o00 && p11 == o10;
        const bool give_right = __shfl_down_sync(0xffffffffu, (int)take_left, 1) != 0 && lane < 31;
        const bool any = __any_sync(0xffffffffu, t.w00 != 0.f || t.w01 != 0.f || t.w10 != 0.f || t.w11 != 0.f);
        if (!any) continue;   // warp-uniform
        float4* q00 = reinterpret_cast<float4*>(ws + (size_t)o00 * C);
        float4* q01 = reinterpret_cast<float4*>(ws + (size_t)o01 * C);
        float4* q10 = reinterpret_cast<float4*>(ws + (size_t)o10 * C);
        float4* q11 = reinterpret_cast<float4*>(ws + (size_t)o11 * C);
        const float* gb = grad_out + ((size_t)b * C * D + d) * P + (size_t)y * w + x;
        const size_t cs = (size_t)D * P;
        float4 g = make_float4(__ldg(gb), __ldg(gb + cs), __ldg(gb + 2 * cs), __ldg(gb + 3 * cs));
        for (int c4 = 0; c4 < C / 4; ++c4) {
            float4 gn = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c4 + 1 < C / 4) {   // next channel quad in flight under the reductions
                const float* gq = gb + (size_t)(c4 + 1) * 4 * cs;
                gn = make_float4(__ldg(gq), __ldg(gq + cs), __ldg(gq + 2 * cs), __ldg(gq + 3 * cs));
            }
            float4 v00 = scale4(t.w00, g), v10 = scale4(t.w10, g);
            const float4 v01 = scale4(t.w01, g), v11 = scale4(t.w11, g);
            const float4 r01 = shfl_up4(v01), r11 = shfl_up4(v11);   // the previous lane's right-hand column
            bool left0 = t.w00 != 0.f, left1 = t.w10 != 0.f;
            if (take_left) {
                v00 = add4(v00, r01);
                v10 = add4(v10, r11);
                left0 = left1 = true;   // the neighbour's share may be non-zero where mine is zero (adding 0 is harmless)
            }
            if (left0) atomicAdd(q00 + c4, v00);
            if (left1) atomicAdd(q10 + c4, v10);
            if (!give_right) {
                if (t.w01 != 0.f) atomicAdd(q01 + c4, v01);
                if (t.w11 != 0.f) atomicAdd(q11 + c4, v11);
            }
            g = gn;
        }
    }
}

// [B,P,C] -> [B,C,P]: consecutive lanes are consecutive pixels, so the plane writes are coalesced; a lane's C floats are
// contiguous reads (C % 4 == 0).
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const float* __restrict__ in, int B, int C, long long P, float* __restrict__ out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * P) return;
    const int b = (int)(i / P);
    const long long p = i % P;
    const float4* q = reinterpret_cast<const float4*>(in + (size_t)i * C);
    float* o = out + (size_t)b * C * P + p;
    for (int c4 = 0; c4 < C / 4; ++c4) {
        const float4 v = __ldg(q + c4);
        o[(size_t)(4 * c4) * P] = v.x;
        o[(size_t)(4 * c4 + 1) * P] = v.y;
        o[(size_t)(4 * c4 + 2) * P] = v.z;
        o[(size_t)(4 * c4 + 3) * P] = v.w;
    }
}

// ------------------------------------------------------------------------------------------
// A5 backward: depth = sum_d p_d * depth_d.  One thread per pixel, planes strided by P (coalesced).
// grad_p / grad_dv [B,D,h,w] (either may be NULL); depth is [B,D] or [B,D,h,w].
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) depth_regress_backward_kernel(const float* __restrict__ grad_depth, const float* __restrict__ prob,
                                                                     const float* __restrict__ depth, int per_pixel, int B, int D,
                                                                     long long P, float* __restrict__ grad_p,
                                                                     float* __restrict__ grad_dv) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= (long long)B * P) return;
    const int b = (int)(i / P);
    const long long p = i % P;
    const float g = __ldg(grad_depth + i);
    const size_t base = (size_t)b * D * P + p;
    for (int d = 0; d < D; ++d) {
        if (grad_p) {
            const float dep = per_pixel ? __ldg(depth + base + (size_t)d * P) : __ldg(depth + (size_t)b * D + d);
            grad_p[base + (size_t)d * P] = g * dep;
        }
        if (grad_dv) grad_dv[base + (size_t)d * P] = g * __ldg(prob + base + (size_t)d * P);
    }
}

// ------------------------------------------------------------------------------------------
// Stage loss (models/losses.py:14-23): sums[0] = sum over mask of smooth_l1(est/iv - gt/iv), sums[1] = |mask|,
// sums[2] = sum over mask of norm_curv.  fp64 accumulators, one atomic triple per block.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ double block_sum(double v, double* red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += red[k];
    return s;
}

__global__ void __launch_bounds__(256) stage_loss_forward_kernel(const float* __restrict__ est, const float* __restrict__ gt,
                                                                 const float* __restrict__ mask, const float* __restrict__ interval,
                                                                 const float* __restrict__ curv, int B, long long P,
                                                                 double* __restrict__ sums) {
    __shared__ double red[8];
    double l = 0.0, n = 0.0, c = 0.0;
    const long long total = (long long)B * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (!(__ldg(mask + i) > 0.5f)) continue;
        const float iv = __ldg(interval + i / P);
        const float diff = __ldg(est + i) / iv - __ldg(gt + i) / iv;
        const float a = fabsf(diff);
        l += (double)(a < 1.f ? 0.5f * diff * diff : a - 0.5f);
        n += 1.0;
        if (curv) c += (double)__ldg(curv + i);
    }
    l = block_sum(l, red);
    n = block_sum(n, red);
    c = block_sum(c, red);
    if (threadIdx.x == 0) {
        atomicAdd(sums, l);
        atomicAdd(sums + 1, n);
        atomicAdd(sums + 2, c);
    }
}

// d/d est = g_depth * smooth_l1'(diff) / (iv * |mask|), d/d curv = g_curv / |mask| on the mask, 0 elsewhere.
// g_depth / g_curv: the upstream gradients of the two means (device scalars).
__global__ void __launch_bounds__(256) stage_loss_backward_kernel(const float* __restrict__ est, const float* __restrict__ gt,
                                                                  const float* __restrict__ mask, const float* __restrict__ interval,
                                                                  const double* __restrict__ sums, const float* __restrict__ g_depth,
                                                                  const float* __restrict__ g_curv, int B, long long P,
                                                                  float* __restrict__ grad_est, float* __restrict__ grad_curv) {
    const long long total = (long long)B * P;
    const float inv_n = (float)(1.0 / sums[1]);
    const float gd = g_depth ? __ldg(g_depth) : 0.f, gc = g_curv ? __ldg(g_curv) : 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const bool on = __ldg(mask + i) > 0.5f;
        float ge = 0.f;
        if (on && grad_est) {
            const float iv = __ldg(interval + i / P);
            const float diff = __ldg(est + i) / iv - __ldg(gt + i) / iv;
            const float s = fabsf(diff) < 1.f ? diff : (diff > 0.f ? 1.f : (diff < 0.f ? -1.f : 0.f));
            ge = gd * s * inv_n / iv;
        }
        if (grad_est) grad_est[i] = ge;
        if (grad_curv) grad_curv[i] = on ? gc * inv_n : 0.f;
    }
}

// ------------------------------------------------------------------------------------------
// feat_distance term of final_loss (models/losses.py:25-35): binary cross entropy with logits over the mask repeated
// across the D planes, positives weighted by neg / pos.  counts[0] = sum of target over the selection, counts[1] = its size.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) feat_loss_count_kernel(const float* __restrict__ target, const float* __restrict__ mask, int B, int D,
                                                              long long P, double* __restrict__ counts) {
    __shared__ double red[8];
    double pos = 0.0, n = 0.0;
    const long long total = (long long)B * D * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / (D * P), p = i % P;
        if (!(__ldg(mask + b * P + p) > 0.5f)) continue;
        pos += (double)__ldg(target + i);
        n += 1.0;
    }
    pos = block_sum(pos, red);
    n = block_sum(n, red);
    if (threadIdx.x == 0) {
        atomicAdd(counts, pos);
        atomicAdd(counts + 1, n);
    }
}

// loss_i = (1 - y) x + (1 + (pw - 1) y) softplus(-x), pw = (n - pos) / pos; softplus(-x) = max(-x, 0) + log1p(exp(-|x|))
__global__ void __launch_bounds__(256) feat_loss_forward_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                                const float* __restrict__ mask, const double* __restrict__ counts, int B,
                                                                int D, long long P, double* __restrict__ loss_sum) {
    __shared__ double red[8];
    const float pw = (float)((counts[1] - counts[0]) / counts[0]);
    double l = 0.0;
    const long long total = (long long)B * D * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / (D * P), p = i % P;
        if (!(__ldg(mask + b * P + p) > 0.5f)) continue;
        const float x = __ldg(logits + i), y = __ldg(target + i);
        const float sp = fmaxf(-x, 0.f) + log1pf(expf(-fabsf(x)));
        l += (double)((1.f - y) * x + (1.f + (pw - 1.f) * y) * sp);
    }
    l = block_sum(l, red);
    if (threadIdx.x == 0) atomicAdd(loss_sum, l);
}

// d/dx = g * ((1 - y) - (1 + (pw - 1) y) sigmoid(-x)) / n on the selection, 0 elsewhere
__global__ void __launch_bounds__(256) feat_loss_backward_kernel(const float* __restrict__ logits, const float* __restrict__ target,
                                                                 const float* __restrict__ mask, const double* __restrict__ counts,
                                                                 const float* __restrict__ g, int B, int D, long long P,
                                                                 float* __restrict__ grad) {
    const float pw = (float)((counts[1] - counts[0]) / counts[0]);
    const float scale = __ldg(g) * (float)(1.0 / counts[1]);
    const long long total = (long long)B * D * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / (D * P), p = i % P;
        float r = 0.f;
        if (__ldg(mask + b * P + p) > 0.5f) {
            const float x = __ldg(logits + i), y = __ldg(target + i);
            const float sg = 1.f / (1.f + expf(x));   // sigmoid(-x)
            r = scale * ((1.f - y) - (1.f + (pw - 1.f) * y) * sg);
        }
        grad[i] = r;
    }
}

}  // namespace

extern "C" {

int cds_homo_warp_backward(const float* grad_out, const float* coef, const float* depth, int depth_per_pixel, int B, int C,
                           int D, int h, int w, float* grad_src, float* workspace, cudaStream_t stream) {
    CDS_REQUIRE(grad_out && coef && depth && grad_src, CDS_EARG, "cds_homo_warp_backward: null pointer");
    CDS_REQUIRE(B > 0 && C > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE, "cds_homo_warp_backward: bad shape B=%d C=%d D=%d h=%d w=%d", B, C, D, h, w);
    CDS_REQUIRE(!workspace || C % 4 == 0, CDS_EUNSUPPORTED, "cds_homo_warp_backward: the channels-last workspace form needs C %% 4 == 0 (C=%d)", C);
    long long total = (long long)B * D * h * w;
    int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
    if (!workspace) {
        homo_warp_backward_kernel<<<blocks, 256, 0, stream>>>(grad_out, coef, depth, depth_per_pixel, B, C, D, h, w, grad_src);
        return cds_check_launch("cds_homo_warp_backward");
    }
    homo_warp_backward_nhwc_kernel<<<blocks, 256, 0, stream>>>(grad_out, coef, depth, depth_per_pixel, B, C, D, h, w, workspace);
    int rc = cds_check_launch("cds_homo_warp_backward");
    if (rc) return rc;
    long long P = (long long)h * w;
    nhwc_to_nchw_kernel<<<cds_div_up(B * P, 256), 256, 0, stream>>>(workspace, B, C, P, grad_src);
    return cds_check_launch("cds_homo_warp_backward(layout)");
}

int cds_depth_regress_backward(const float* grad_depth, const float* prob, const float* depth, int depth_per_pixel, int B, int D,
                               int h, int w, float* grad_p, float* grad_dv, cudaStream_t stream) {
    CDS_REQUIRE(grad_depth && (grad_p || grad_dv), CDS_EARG, "cds_depth_regress_backward: null pointer");
    CDS_REQUIRE((!grad_p || depth) && (!grad_dv || prob), CDS_EARG, "cds_depth_regress_backward: grad_p needs depth, grad_dv needs prob");
    CDS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, CDS_ESHAPE, "cds_depth_regress_backward: bad shape");
    long long P = (long long)h * w;
    depth_regress_backward_kernel<<<cds_div_up(B * P, 256), 256, 0, stream>>>(grad_depth, prob, depth, depth_per_pixel, B, D, P, grad_p,
                                                                             grad_dv);
    return cds_check_launch("cds_depth_regress_backward");
}

int cds_stage_loss_forward(const float* est, const float* gt, const float* mask, const float* interval, const float* norm_curv,
                           int B, int h, int w, double* sums, cudaStream_t stream) {
    CDS_REQUIRE(est && gt && mask && interval && sums, CDS_EARG, "cds_stage_loss_forward: null pointer");
    CDS_REQUIRE(B > 0 && h > 0 && w > 0, CDS_ESHAPE, "cds_stage_loss_forward: bad shape");
    long long P = (long long)h * w, total = (long long)B * P;
    int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    stage_loss_forward_kernel<<<blocks, 256, 0, stream>>>(est, gt, mask, interval, norm_curv, B, P, sums);
    return cds_check_launch("cds_stage_loss_forward");
}

int cds_stage_loss_backward(const float* est, const float* gt, const float* mask, const float* interval, const double* sums,
                            const float* g_depth, const float* g_curv, int B, int h, int w, float* grad_est, float* grad_curv,
                            cudaStream_t stream) {
    CDS_REQUIRE(est && gt && mask && interval && sums && (grad_est || grad_curv), CDS_EARG, "cds_stage_loss_backward: null pointer");
    CDS_REQUIRE((!grad_est || g_depth) && (!grad_curv || g_curv), CDS_EARG, "cds_stage_loss_backward: missing upstream gradient");
    CDS_REQUIRE(B > 0 && h > 0 && w > 0, CDS_ESHAPE, "cds_stage_loss_backward: bad shape");
    long long P = (long long)h * w, total = (long long)B * P;
    int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    stage_loss_backward_kernel<<<blocks, 256, 0, stream>>>(est, gt, mask, interval, sums, g_depth, g_curv, B, P, grad_est, grad_curv);
    return cds_check_launch("cds_stage_loss_backward");
}

int cds_feat_loss_forward(const float* logits, const float* target, const float* mask, int B, int D, int h, int w, double* sums,
                          cudaStream_t stream) {
    CDS_REQUIRE(logits && target && mask && sums, CDS_EARG, "cds_feat_loss_forward: null pointer");
    CDS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, CDS_ESHAPE, "cds_feat_loss_forward: bad shape");
    long long P = (long long)h * w, total = (long long)B * D * P;
    int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    feat_loss_count_kernel<<<blocks, 256, 0, stream>>>(target, mask, B, D, P, sums);
    int rc = cds_check_launch("cds_feat_loss_forward(count)");
    if (rc) return rc;
    feat_loss_forward_kernel<<<blocks, 256, 0, stream>>>(logits, target, mask, sums, B, D, P, sums + 2);
    return cds_check_launch("cds_feat_loss_forward");
}

int cds_feat_loss_backward(const float* logits, const float* target, const float* mask, const double* sums, const float* g, int B,
                           int D, int h, int w, float* grad, cudaStream_t stream) {
    CDS_REQUIRE(logits && target && mask && sums && g && grad, CDS_EARG, "cds_feat_loss_backward: null pointer");
    CDS_REQUIRE(B > 0 && D > 0 && h > 0 && w > 0, CDS_ESHAPE, "cds_feat_loss_backward: bad shape");
    long long P = (long long)h * w, total = (long long)B * D * P;
    int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    feat_loss_backward_kernel<<<blocks, 256, 0, stream>>>(logits, target, mask, sums, g, B, D, P, grad);
    return cds_check_launch("cds_feat_loss_backward");
}

}  // extern "C"
