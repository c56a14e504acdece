// Geometric-consistency filter: the step right after the depth-inference path.
// Reference: fusion.py:49-117 (get_reproj / project_img / vis_filter / ave_fusion), driven by test.py:326-352.
//
// For a reference view with depth map d_ref and V source views with depth maps d_v:
//   reproj_xyd[v](p) = bilinear sample, at the projection of reference pixel p (lifted with d_ref) into source view v, of the
//                      map  q -> (x, y, depth) of source pixel q (lifted with d_v) seen from the reference camera;
//   in_range[v](p)   = that projection lies inside the source image;
//   masks[v](p)      = in_range & |reproj_xy - p| < img_dist_thresh & |d_ref - reproj_d| < depth_thresh * max(d_ref, reproj_d);
//   vis_mask(p)      = sum_v masks[v] >= vthresh - 1.1;   ave(p) = (sum_v reproj_d * masks[v] + d_ref) / (sum_v masks[v] + 1);
//   points(p)        = ave(p) back-projected to the world frame.
// The reference materialises the (x, y, depth) map of every source view and samples it with grid_sample; here one thread owns
// a reference pixel, walks the views, and lifts the four bilinear taps of each view on the fly (one 4-byte gather per tap), so
// nothing but the requested outputs touches HBM.  The arithmetic follows the reference step by step in fp32 (pixel centres at
// +0.5, 1e-9 added to every homogeneous division, the warp coordinate normalised by W / H, clamped to +-1.1 and un-normalised
// the align_corners=True way), so masks agree except where a test sits within rounding of its threshold.  The camera
// inverses are taken in fp64 by cds_fusion_setup.
#include "cds_common.cuh"

namespace {

constexpr int kMats = 100;   // per (batch item, view): Kr^-1 9 | Er^-1 16 | Es 16 | Ks 9 | Ks^-1 9 | Es^-1 16 | Er 16 | Kr 9
constexpr int oKRI = 0, oERI = 9, oES = 25, oKS = 41, oKSI = 50, oESI = 59, oER = 75, oKR = 91;
constexpr int kMaxFusionViews = 16;

__device__ void inv3(const double* a, double* o) {
    const double c00 = a[4] * a[8] - a[5] * a[7], c01 = a[5] * a[6] - a[3] * a[8], c02 = a[3] * a[7] - a[4] * a[6];
    const double det = a[0] * c00 + a[1] * c01 + a[2] * c02;
    const double id = 1.0 / det;
    o[0] = c00 * id; o[1] = (a[2] * a[7] - a[1] * a[8]) * id; o[2] = (a[1] * a[5] - a[2] * a[4]) * id;
    o[3] = c01 * id; o[4] = (a[0] * a[8] - a[2] * a[6]) * id; o[5] = (a[2] * a[3] - a[0] * a[5]) * id;
    o[6] = c02 * id; o[7] = (a[1] * a[6] - a[0] * a[7]) * id; o[8] = (a[0] * a[4] - a[1] * a[3]) * id;
}
// Gauss-Jordan with partial pivoting
__device__ void inv4(const double* a, double* o) {
    double m[4][8];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) { m[i][j] = a[i * 4 + j]; m[i][4 + j] = i == j ? 1.0 : 0.0; }
    for (int c = 0; c < 4; ++c) {
        int piv = c;
        for (int r = c + 1; r < 4; ++r) if (fabs(m[r][c]) > fabs(m[piv][c])) piv = r;
        for (int j = 0; j < 8; ++j) { double t = m[c][j]; m[c][j] = m[piv][j]; m[piv][j] = t; }
        const double id = 1.0 / m[c][c];
        for (int j = 0; j < 8; ++j) m[c][j] *= id;
        for (int r = 0; r < 4; ++r) {
            if (r == c) continue;
            const double f = m[r][c];
            for (int j = 0; j < 8; ++j) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) o[i * 4 + j] = m[i][4 + j];
}

// cams: [., 2, 4, 4] = (extrinsic, intrinsic in [1, :3, :3])   (fusion.py:22,30)
__global__ void fusion_setup_kernel(const float* __restrict__ ref_cam, const float* __restrict__ srcs_cam, int n, int v,
                                    float* __restrict__ mats) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * v) return;
    const float* rc = ref_cam + (size_t)(i / v) * 32;
    const float* sc = srcs_cam + (size_t)i * 32;
    double Er[16], Es[16], Kr[9], Ks[9], t16[16], t9[9];
    for (int k = 0; k < 16; ++k) { Er[k] = rc[k]; Es[k] = sc[k]; }
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) { Kr[r * 3 + c] = rc[16 + r * 4 + c]; Ks[r * 3 + c] = sc[16 + r * 4 + c]; }
    float* o = mats + (size_t)i * kMats;
    inv3(Kr, t9);  for (int k = 0; k < 9; ++k) o[oKRI + k] = (float)t9[k];
    inv4(Er, t16); for (int k = 0; k < 16; ++k) o[oERI + k] = (float)t16[k];
    for (int k = 0; k < 16; ++k) o[oES + k] = (float)Es[k];
    for (int k = 0; k < 9; ++k) o[oKS + k] = (float)Ks[k];
    inv3(Ks, t9);  for (int k = 0; k < 9; ++k) o[oKSI + k] = (float)t9[k];
    inv4(Es, t16); for (int k = 0; k < 16; ++k) o[oESI + k] = (float)t16[k];
    for (int k = 0; k < 16; ++k) o[oER + k] = (float)Er[k];
    for (int k = 0; k < 9; ++k) o[oKR + k] = (float)Kr[k];
}

struct V3 { float x, y, z; };
struct V4 { float x, y, z, w; };
__device__ __forceinline__ V3 mul3(const float* m, float x, float y, float z) {
    return {m[0] * x + m[1] * y + m[2] * z, m[3] * x + m[4] * y + m[5] * z, m[6] * x + m[7] * y + m[8] * z};
}
__device__ __forceinline__ V4 mul4(const float* m, const V4& a) {
    return {m[0] * a.x + m[1] * a.y + m[2] * a.z + m[3] * a.w, m[4] * a.x + m[5] * a.y + m[6] * a.z + m[7] * a.w,
            m[8] * a.x + m[9] * a.y + m[10] * a.z + m[11] * a.w, m[12] * a.x + m[13] * a.y + m[14] * a.z + m[15] * a.w};
}
// a / (a.w + 1e-9): one correctly rounded reciprocal + multiplies instead of four divisions (<= 1 ulp from the reference's
// quotient; the kernel is bound by its ~60 homogeneous divisions per pixel and view otherwise)
__device__ __forceinline__ V4 hdiv(const V4& a) {
    const float r = __frcp_rn(a.w + 1e-9f);
    return {a.x * r, a.y * r, a.z * r, a.w * r};
}
// pixel (centre px, py) with depth, camera A -> camera frame of B and its image: fusion.py idx_img2cam -> idx_cam2world ->
// idx_world2cam -> idx_cam2img.  Returns image (x, y) in B and the depth in B's camera frame.
__device__ __forceinline__ void transfer(const float* KAi, const float* EAi, const float* EB, const float* KB, float px, float py,
                                         float depth, float& ix, float& iy, float& zc) {
    V3 c = mul3(KAi, px, py, 1.f);
    const float rz = __frcp_rn(c.z + 1e-9f);
    V4 cam = {c.x * rz * depth, c.y * rz * depth, c.z * rz * depth, 1.f};
    V4 wv = hdiv(mul4(EAi, cam));
    V4 cb = hdiv(mul4(EB, wv));
    const float r3 = __frcp_rn(cb.w + 1e-9f);
    V3 im = mul3(KB, cb.x * r3, cb.y * r3, cb.z * r3);
    const float ri = __frcp_rn(im.z + 1e-9f);
    ix = im.x * ri;
    iy = im.y * ri;
    zc = cb.z;
}

__global__ void __launch_bounds__(256) geometric_filter_kernel(const float* __restrict__ ref_depth, const float* __restrict__ srcs_depth,
                                                               const float* __restrict__ mats, int v, int h, int w, float dist_t,
                                                               float depth_t, float vthresh, float* __restrict__ reproj_xyd,
                                                               float* __restrict__ in_range_o, float* __restrict__ masks_o,
                                                               unsigned char* __restrict__ vis_mask, float* __restrict__ ave,
                                                               float* __restrict__ points) {
    extern __shared__ float s_m[];   // [v][kMats]
    const int n = blockIdx.y;
    for (int i = threadIdx.x; i < v * kMats; i += blockDim.x) s_m[i] = __ldg(mats + (size_t)n * v * kMats + i);
    __syncthreads();
    const int P = h * w;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const int x = pix % w, y = pix / w;
    const float pxc = (float)x + 0.5f, pyc = (float)y + 0.5f;
    const float dref = __ldg(ref_depth + (size_t)n * P + pix);
    float msum = 0.f, dsum = 0.f;
    for (int vi = 0; vi < v; ++vi) {
        const float* m = s_m + vi * kMats;
        // reference pixel -> source image (fusion.py:53-61)
        float sx, sy, sz;
        transfer(m + oKRI, m + oERI, m + oES, m + oKS, pxc, pyc, dref, sx, sy, sz);
        float cx = fminf(fmaxf(sx / (float)w * 2.f - 1.f, -1.1f), 1.1f);
        float cy = fminf(fmaxf(sy / (float)h * 2.f - 1.f, -1.1f), 1.1f);
        const bool inr = -1.f <= cx && cx <= 1.f && -1.f <= cy && cy <= 1.f;
        // grid_sample(bilinear, zeros, align_corners=True) of the source view's (x, y, depth)-in-reference map
        const float gx = (cx + 1.f) / 2.f * (float)(w - 1), gy = (cy + 1.f) / 2.f * (float)(h - 1);
        const float fx0 = floorf(gx), fy0 = floorf(gy);
        const int x0 = (int)fx0, y0 = (int)fy0;
        const float tx = gx - fx0, ty = gy - fy0;
        const float wgt[4] = {(1.f - tx) * (1.f - ty), tx * (1.f - ty), (1.f - tx) * ty, tx * ty};
        float rx = 0.f, ry = 0.f, rd = 0.f;
        const float* sd = srcs_depth + ((size_t)n * v + vi) * P;
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int xi = x0 + (t & 1), yi = y0 + (t >> 1);
            if (xi < 0 || xi >= w || yi < 0 || yi >= h) continue;   // zero padding
            float qx, qy, qz;
            transfer(m + oKSI, m + oESI, m + oER, m + oKR, (float)xi + 0.5f, (float)yi + 0.5f, __ldg(sd + (size_t)yi * w + xi), qx, qy, qz);
            rx += qx * wgt[t];
            ry += qy * wgt[t];
            rd += qz * wgt[t];
        }
        const size_t o = ((size_t)n * v + vi);
        if (reproj_xyd) {
            reproj_xyd[(o * 3 + 0) * P + pix] = rx;
            reproj_xyd[(o * 3 + 1) * P + pix] = ry;
            reproj_xyd[(o * 3 + 2) * P + pix] = rd;
        }
        if (in_range_o) in_range_o[o * P + pix] = inr ? 1.f : 0.f;
        // vis_filter (fusion.py:103-112)
        const float ex = rx - pxc, ey = ry - pyc;
        const bool dist_ok = sqrtf(ex * ex + ey * ey) < dist_t;
        const bool dep_ok = fabsf(dref - rd) < fmaxf(dref, rd) * depth_t;
        const float mk = (inr && dist_ok && dep_ok) ? 1.f : 0.f;
        if (masks_o) masks_o[o * P + pix] = mk;
        msum += mk;
        dsum += rd * mk;
    }
    const float av = (dsum + dref) / (msum + 1.f);   // ave_fusion (fusion.py:115-117)
    if (vis_mask) vis_mask[(size_t)n * P + pix] = msum >= vthresh - 1.1f ? 1 : 0;
    if (ave) ave[(size_t)n * P + pix] = av;
    if (points) {   // test.py:345-347: the fused depth back-projected with the reference camera
        const float* m = s_m;
        V3 c = mul3(m + oKRI, pxc, pyc, 1.f);
        const float rz = __frcp_rn(c.z + 1e-9f);
        V4 wv = hdiv(mul4(m + oERI, V4{c.x * rz * av, c.y * rz * av, c.z * rz * av, 1.f}));
        points[((size_t)n * 3 + 0) * P + pix] = wv.x;
        points[((size_t)n * 3 + 1) * P + pix] = wv.y;
        points[((size_t)n * 3 + 2) * P + pix] = wv.z;
    }
}

// vis_filter + ave_fusion on materialised reprojections (the reference's op-level signatures, fusion.py:103-117)
__global__ void __launch_bounds__(256) vis_filter_kernel(const float* __restrict__ ref_depth, const float* __restrict__ reproj_xyd,
                                                         const float* __restrict__ in_range, int v, int h, int w, float dist_t,
                                                         float depth_t, float vthresh, const float* __restrict__ masks_in,
                                                         float* __restrict__ masks_o, unsigned char* __restrict__ vis_mask,
                                                         float* __restrict__ ave) {
    const int n = blockIdx.y, P = h * w;
    const int pix = blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= P) return;
    const float pxc = (float)(pix % w) + 0.5f, pyc = (float)(pix / w) + 0.5f;
    const float dref = __ldg(ref_depth + (size_t)n * P + pix);
    float msum = 0.f, dsum = 0.f;
    for (int vi = 0; vi < v; ++vi) {
        const size_t o = (size_t)n * v + vi;
        const float rd = __ldg(reproj_xyd + (o * 3 + 2) * P + pix);
        float mk;
        if (masks_in) {
            mk = __ldg(masks_in + o * P + pix);
        } else {
            const float ex = __ldg(reproj_xyd + (o * 3 + 0) * P + pix) - pxc, ey = __ldg(reproj_xyd + (o * 3 + 1) * P + pix) - pyc;
            const bool ok = sqrtf(ex * ex + ey * ey) < dist_t && fabsf(dref - rd) < fmaxf(dref, rd) * depth_t;
            mk = fminf(__ldg(in_range + o * P + pix), ok ? 1.f : 0.f);
            if (masks_o) masks_o[o * P + pix] = mk;
        }
        msum += mk;
        dsum += rd * mk;
    }
    if (vis_mask) vis_mask[(size_t)n * P + pix] = msum >= vthresh - 1.1f ? 1 : 0;
    if (ave) ave[(size_t)n * P + pix] = (dsum + dref) / (msum + 1.f);
}

// prob_filter (fusion.py:69-77): mask = AND_c prob[c] > thresh[c]; optionally depth *= mask (test.py:333-335)
__global__ void prob_filter_kernel(const float* __restrict__ prob, int C, long long HW, float t0, float t1, float t2, float t3,
                                   unsigned char* __restrict__ mask, float* __restrict__ depth_inout) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int n = blockIdx.y;
    if (i >= HW) return;
    const float th[4] = {t0, t1, t2, t3};
    bool ok = true;
    for (int c = 0; c < C; ++c) ok = ok && (__ldg(prob + ((size_t)n * C + c) * HW + i) > th[c]);
    if (mask) mask[(size_t)n * HW + i] = ok ? 1 : 0;
    if (depth_inout && !ok) depth_inout[(size_t)n * HW + i] *= 0.f;
}

}  // namespace

extern "C" {

int cds_fusion_mats_floats(int n, int v) { return n * v * kMats; }

// vis_filter (masks_in NULL: masks computed from reproj_xyd / in_range and written to masks_out if given) and / or ave_fusion
// (masks_in given: the reference's ave_fusion(ref_depth, reproj_xyd, masks)).  vis_mask [n,h,w] uint8, ave [n,h,w]; any NULL.
int cds_vis_filter(const float* ref_depth, const float* reproj_xyd, const float* in_range, const float* masks_in, int n, int v, int h,
                   int w, float img_dist_thresh, float depth_thresh, float vthresh, float* masks_out, unsigned char* vis_mask,
                   float* ave, cudaStream_t stream) {
    CDS_REQUIRE(ref_depth && reproj_xyd && (in_range || masks_in), CDS_EARG, "cds_vis_filter: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && v > 0 && h > 0 && w > 0 && (long long)h * w < (1ll << 30), CDS_ESHAPE, "cds_vis_filter: bad shape");
    dim3 grid(cds_div_up((long long)h * w, 256), n);
    vis_filter_kernel<<<grid, 256, 0, stream>>>(ref_depth, reproj_xyd, in_range, v, h, w, img_dist_thresh, depth_thresh, vthresh, masks_in,
                                                masks_out, vis_mask, ave);
    return cds_check_launch("cds_vis_filter");
}

// prob [n,C,h,w] (C <= 4), thresholds[C] on the HOST; mask [n,h,w] uint8 and / or depth_inout [n,h,w] *= mask
int cds_prob_filter(const float* prob, const float* thresholds, int n, int C, int h, int w, unsigned char* mask, float* depth_inout,
                    cudaStream_t stream) {
    CDS_REQUIRE(prob && thresholds && (mask || depth_inout), CDS_EARG, "cds_prob_filter: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && C >= 1 && C <= 4 && h > 0 && w > 0, CDS_ESHAPE, "cds_prob_filter: bad shape (1 <= C <= 4)");
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    for (int c = 0; c < C; ++c) t[c] = thresholds[c];
    const long long HW = (long long)h * w;
    dim3 grid(cds_div_up(HW, 256), n);
    prob_filter_kernel<<<grid, 256, 0, stream>>>(prob, C, HW, t[0], t[1], t[2], t[3], mask, depth_inout);
    return cds_check_launch("cds_prob_filter");
}


// ref_cam [n,2,4,4], srcs_cam [n,v,2,4,4] fp32 -> mats [n,v,100] fp32 (camera matrices and their fp64-computed inverses)
int cds_fusion_setup(const float* ref_cam, const float* srcs_cam, int n, int v, float* mats, cudaStream_t stream) {
    CDS_REQUIRE(ref_cam && srcs_cam && mats, CDS_EARG, "cds_fusion_setup: null pointer");
    CDS_REQUIRE(n > 0 && v > 0 && v <= kMaxFusionViews, CDS_ESHAPE, "cds_fusion_setup: need n > 0 and 1 <= v <= %d (got %d, %d)", kMaxFusionViews, n, v);
    fusion_setup_kernel<<<cds_div_up(n * v, 64), 64, 0, stream>>>(ref_cam, srcs_cam, n, v, mats);
    return cds_check_launch("cds_fusion_setup");
}

// ref_depth [n,h,w], srcs_depth [n,v,h,w], mats from cds_fusion_setup.  Every output may be NULL:
// reproj_xyd [n,v,3,h,w], in_range [n,v,h,w], masks [n,v,h,w] (fp32 0/1), vis_mask [n,h,w] (uint8), ave [n,h,w], points [n,3,h,w]
int cds_geometric_filter(const float* ref_depth, const float* srcs_depth, const float* mats, int n, int v, int h, int w,
                         float img_dist_thresh, float depth_thresh, float vthresh, float* reproj_xyd, float* in_range, float* masks,
                         unsigned char* vis_mask, float* ave, float* points, cudaStream_t stream) {
    CDS_REQUIRE(ref_depth && srcs_depth && mats, CDS_EARG, "cds_geometric_filter: null pointer");
    CDS_REQUIRE(n > 0 && n <= 65535 && v > 0 && v <= kMaxFusionViews && h > 1 && w > 1 && (long long)h * w < (1ll << 30), CDS_ESHAPE,
                "cds_geometric_filter: bad shape n=%d v=%d h=%d w=%d (1 <= v <= %d)", n, v, h, w, kMaxFusionViews);
    dim3 grid(cds_div_up((long long)h * w, 256), n);
    geometric_filter_kernel<<<grid, 256, (size_t)v * kMats * sizeof(float), stream>>>(ref_depth, srcs_depth, mats, v, h, w, img_dist_thresh,
                                                                                      depth_thresh, vthresh, reproj_xyd, in_range, masks,
                                                                                      vis_mask, ave, points);
    return cds_check_launch("cds_geometric_filter");
}

}  // extern "C"
