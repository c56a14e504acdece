// Cost-volume build (A2) fused with the plane-sweep warp (A1): the warped source features are
// never materialised.  Two sweeps around the visibility net (A3):
//   pass 1  per source view: similarity sum_c ref*warp per plane -> online softmax entropy over D
//   pass 2  visibility-weighted mean over views of ref (.) warp, all C channels kept, written as
//           a channel-blocked channels-last [B, C/8, D, h, w, 8] volume
// Reference: models/model.py:34-60,74 (loop), models/utils/warping.py:84-101 (coordinates, gather).
//
// Layout: features are channels-last [V, B, h, w, C] (one ref and one src map per source view,
// because the ref features depend on the pair's epipole, model.py:154-161); a thread owns one
// (pixel, 8-channel chunk) so each bilinear tap is one 16-byte (fp16) load and the C/8 lanes of
// a pixel reduce the channel sum with warp shuffles.
#include "cds_common.cuh"

namespace {

constexpr int kMaxViews = 8;

template <typename T>
__device__ __forceinline__ void gather8(const T* __restrict__ fea, int w, int h, int C, const Taps& t, float (&out)[8]) {
    int xa = min(max(t.x0, 0), w - 1), xb = min(max(t.x0 + 1, 0), w - 1);
    int ya = min(max(t.y0, 0), h - 1), yb = min(max(t.y0 + 1, 0), h - 1);
    float a[8], b[8], c[8], d[8];
    Vec8<T>::load(fea + ((size_t)ya * w + xa) * C, a);
    Vec8<T>::load(fea + ((size_t)ya * w + xb) * C, b);
    Vec8<T>::load(fea + ((size_t)yb * w + xa) * C, c);
    Vec8<T>::load(fea + ((size_t)yb * w + xb) * C, d);
#pragma unroll
    for (int i = 0; i < 8; ++i) out[i] = t.w00 * a[i] + t.w01 * b[i] + t.w10 * c[i] + t.w11 * d[i];
}

// ---------------------------------------------------------------------------------------------
// pass 1: entropy[v, b, y, x] = H(softmax_d(sum_c ref[c] * warp_d[c]))
// ---------------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256, 4) entropy_kernel(const T* __restrict__ ref_fea, const T* __restrict__ src_fea,
                                                      const float* __restrict__ coef, const float* __restrict__ depth,
                                                      int V, int B, int D, int h, int w, float* __restrict__ entropy) {
    constexpr int LPP = C / 8;  // lanes per pixel
    const long long P = (long long)h * w;
    const long long total = (long long)V * B * P * LPP;
    long long gid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    // total is padded by the host to a multiple of the block size only through this guard;
    // lanes of one pixel are always in the same warp because LPP divides 32
    bool live = gid < total;
    long long g = live ? gid : total - 1;
    int chunk = (int)(g % LPP);
    long long pix = g / LPP;
    int x = (int)(pix % w), y = (int)((pix / w) % h);
    int b = (int)((pix / P) % B), v = (int)(pix / (P * B));

    const T* rf = ref_fea + (((size_t)v * B + b) * P + (size_t)y * w + x) * C + chunk * 8;
    const T* sf = src_fea + ((size_t)v * B + b) * P * C + chunk * 8;
    float ref[8];
    Vec8<T>::load(rf, ref);
    WarpCoef k = load_coef(coef + ((size_t)b * V + v) * 12);
    float rx, ry, rz;
    pixel_ray(k, (float)x, (float)y, rx, ry, rz);
    const float* dp = depth + (size_t)b * D * P + (size_t)y * w + x;

    // online softmax statistics: m = running max, S = sum e^(s-m), A = sum (s-m) e^(s-m)
    float m = -INFINITY, S = 0.f, A = 0.f;
    for (int d = 0; d < D; ++d) {
        float dep = __ldg(dp + (size_t)d * P);
        float u, vv;
        project_fast(k, rx, ry, rz, dep, u, vv);
        Taps t = make_taps(u, vv, w, h);
        float wv[8];
        gather8<T>(sf, w, h, C, t, wv);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) s += ref[i] * wv[i];
#pragma unroll
        for (int o = 1; o < LPP; o <<= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (s > m) {
            float delta = m - s;  // <= 0 (or -inf on the first plane)
            float e = __expf(delta);
            A = (d == 0) ? 0.f : e * (A + delta * S);
            S = (d == 0) ? 0.f : S * e;
            m = s;
        }
        float z = s - m;
        float e = __expf(z);
        S += e;
        A += z * e;
    }
    if (live && chunk == 0) entropy[((size_t)v * B + b) * P + (size_t)y * w + x] = __logf(S) - A / S;
}

// ---------------------------------------------------------------------------------------------
// pass 2: volume[b, d, y, x, :] = sum_v vis_v * ref_v (.) warp_{v,d} / (sum_v vis_v + 1e-6)
// ---------------------------------------------------------------------------------------------
template <typename T, int C>
__global__ void __launch_bounds__(256, 4) aggregate_kernel(const T* __restrict__ ref_fea, const T* __restrict__ src_fea,
                                                        const float* __restrict__ coef, const float* __restrict__ depth,
                                                        const float* __restrict__ vis, int V, int B, int D, int h, int w,
                                                        T* __restrict__ volume) {
    constexpr int LPP = C / 8;
    const long long P = (long long)h * w;
    const long long total = (long long)B * P * LPP;
    long long g = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (g >= total) return;
    int chunk = (int)(g % LPP);
    long long pix = g / LPP;
    int x = (int)(pix % w), y = (int)((pix / w) % h), b = (int)(pix / P);
    size_t pofs = (size_t)y * w + x;

    float vw[kMaxViews];
    float vsum = 0.f;
#pragma unroll
    for (int v = 0; v < kMaxViews; ++v) {
        vw[v] = (v < V) ? __ldg(vis + ((size_t)v * B + b) * P + pofs) : 0.f;
        if (v < V) vsum += vw[v];  // same accumulation order as the reference loop (model.py:59)
    }
    float inv = 1.f / (vsum + 1e-6f);
    const float* dp = depth + (size_t)b * D * P + pofs;
    // channel-blocked volume [B][C/8][D][h][w][8]: one 8-channel slab of a row is contiguous (what conv0's TMA wants)
    T* outp = volume + (((size_t)b * LPP + chunk) * D * P + pofs) * 8;

    for (int d = 0; d < D; ++d) {
        float dep = __ldg(dp + (size_t)d * P);
        float acc[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
        for (int v = 0; v < kMaxViews; ++v) {
            if (v < V) {
                WarpCoef k = load_coef(coef + ((size_t)b * V + v) * 12);
                float rx, ry, rz, u, vv;
                pixel_ray(k, (float)x, (float)y, rx, ry, rz);
                project_fast(k, rx, ry, rz, dep, u, vv);
                Taps t = make_taps(u, vv, w, h);
                float wv[8], ref[8];
                gather8<T>(src_fea + ((size_t)v * B + b) * P * C + chunk * 8, w, h, C, t, wv);
                Vec8<T>::load(ref_fea + (((size_t)v * B + b) * P + pofs) * C + chunk * 8, ref);
#pragma unroll
                for (int i = 0; i < 8; ++i) acc[i] += (ref[i] * wv[i]) * vw[v];
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] *= inv;
        Vec8<T>::store(outp + (size_t)d * P * 8, acc);
    }
}

// mean over views of (ref_nc_sum + src_nc_sum) / 2   (model.py:60,79)
__global__ void nc_mean_kernel(const float* __restrict__ ref_nc, const float* __restrict__ src_nc, int V, long long n,
                               float* __restrict__ out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int v = 0; v < V; ++v) s += (__ldg(ref_nc + (size_t)v * n + i) + __ldg(src_nc + (size_t)v * n + i)) / 2.f;
    out[i] = s / (float)V;
}

template <typename T>
int launch_entropy(const void* ref, const void* src, const float* coef, const float* depth, int V, int B, int C, int D,
                   int h, int w, float* entropy, cudaStream_t st) {
    long long total = (long long)V * B * h * w * (C / 8);
    int blocks = cds_div_up(total, 256);
    const T* r = (const T*)ref;
    const T* s = (const T*)src;
    switch (C) {
        case 8: entropy_kernel<T, 8><<<blocks, 256, 0, st>>>(r, s, coef, depth, V, B, D, h, w, entropy); break;
        case 16: entropy_kernel<T, 16><<<blocks, 256, 0, st>>>(r, s, coef, depth, V, B, D, h, w, entropy); break;
        case 32: entropy_kernel<T, 32><<<blocks, 256, 0, st>>>(r, s, coef, depth, V, B, D, h, w, entropy); break;
        default: cds_set_error("cds_costvol_entropy: C must be 8, 16 or 32 (got %d)", C); return CDS_EUNSUPPORTED;
    }
    return cds_check_launch("cds_costvol_entropy");
}

template <typename T>
int launch_aggregate(const void* ref, const void* src, const float* coef, const float* depth, const float* vis, int V,
                     int B, int C, int D, int h, int w, void* volume, cudaStream_t st) {
    long long total = (long long)B * h * w * (C / 8);
    int blocks = cds_div_up(total, 256);
    const T* r = (const T*)ref;
    const T* s = (const T*)src;
    T* o = (T*)volume;
    switch (C) {
        case 8: aggregate_kernel<T, 8><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o); break;
        case 16: aggregate_kernel<T, 16><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o); break;
        case 32: aggregate_kernel<T, 32><<<blocks, 256, 0, st>>>(r, s, coef, depth, vis, V, B, D, h, w, o); break;
        default: cds_set_error("cds_costvol_aggregate: C must be 8, 16 or 32 (got %d)", C); return CDS_EUNSUPPORTED;
    }
    return cds_check_launch("cds_costvol_aggregate");
}

}  // namespace

extern "C" {

int cds_costvol_entropy(const void* ref_fea, const void* src_fea, const float* coef, const float* depth, int V, int B,
                        int C, int D, int h, int w, int dtype, float* entropy, cudaStream_t stream) {
    CDS_REQUIRE(ref_fea && src_fea && coef && depth && entropy, CDS_EARG, "cds_costvol_entropy: null pointer");
    CDS_REQUIRE(V >= 1 && V <= kMaxViews && B > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE,
                "cds_costvol_entropy: bad shape V=%d B=%d D=%d h=%d w=%d (1 <= V <= %d)", V, B, D, h, w, kMaxViews);
    if (dtype == CDS_F16) return launch_entropy<__half>(ref_fea, src_fea, coef, depth, V, B, C, D, h, w, entropy, stream);
    if (dtype == CDS_F32) return launch_entropy<float>(ref_fea, src_fea, coef, depth, V, B, C, D, h, w, entropy, stream);
    cds_set_error("cds_costvol_entropy: unknown dtype %d", dtype);
    return CDS_EARG;
}

int cds_costvol_aggregate(const void* ref_fea, const void* src_fea, const float* coef, const float* depth,
                          const float* vis, int V, int B, int C, int D, int h, int w, int dtype, void* volume,
                          cudaStream_t stream) {
    CDS_REQUIRE(ref_fea && src_fea && coef && depth && vis && volume, CDS_EARG, "cds_costvol_aggregate: null pointer");
    CDS_REQUIRE(V >= 1 && V <= kMaxViews && B > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE,
                "cds_costvol_aggregate: bad shape V=%d B=%d D=%d h=%d w=%d (1 <= V <= %d)", V, B, D, h, w, kMaxViews);
    if (dtype == CDS_F16) return launch_aggregate<__half>(ref_fea, src_fea, coef, depth, vis, V, B, C, D, h, w, volume, stream);
    if (dtype == CDS_F32) return launch_aggregate<float>(ref_fea, src_fea, coef, depth, vis, V, B, C, D, h, w, volume, stream);
    cds_set_error("cds_costvol_aggregate: unknown dtype %d", dtype);
    return CDS_EARG;
}

int cds_nc_mean(const float* ref_nc_sum, const float* src_nc_sum, int V, long long n, float* out, cudaStream_t stream) {
    CDS_REQUIRE(ref_nc_sum && src_nc_sum && out && V >= 1 && n > 0, CDS_EARG, "cds_nc_mean: bad arguments");
    nc_mean_kernel<<<cds_div_up(n, 256), 256, 0, stream>>>(ref_nc_sum, src_nc_sum, V, n, out);
    return cds_check_launch("cds_nc_mean");
}

}  // extern "C"
