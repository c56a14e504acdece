// Camera algebra, plane-sweep warp (A1), depth hypotheses (A9).
//
// Reference semantics restated (file:line into TruongKhang/cds-mvsnet):
//   models/model.py:40-43          projection = K[:3,:3] @ E[:3,:4] in rows 0-2 of the extrinsic
//   models/utils/warping.py:80-82  proj = src_proj @ inv(ref_proj); rot = proj[:3,:3]; trans = proj[:3,3]
//   models/dynamic_conv.py:19-47   fundamental matrix and the c=1e3 epipole solve
//   models/utils/warping.py:84-101 per-plane projective coordinates + bilinear zero-padded gather
//   models/module.py:394-439, models/model.py:174-193  depth hypotheses
#include "cds_common.cuh"

namespace {

// ------------------------------------------------------------------------------------------
// small dense algebra in fp64 (one thread per matrix; these are 3x3 / 4x4 / 2x2 problems)
// ------------------------------------------------------------------------------------------
template <int N>
__device__ bool invert(const double* a, double* inv) {
    double m[N][2 * N];
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            m[i][j] = a[i * N + j];
            m[i][N + j] = (i == j) ? 1.0 : 0.0;
        }
    for (int c = 0; c < N; ++c) {
        int p = c;
        double best = fabs(m[c][c]);
        for (int r = c + 1; r < N; ++r)
            if (fabs(m[r][c]) > best) { best = fabs(m[r][c]); p = r; }
        if (best == 0.0) return false;
        if (p != c)
            for (int j = 0; j < 2 * N; ++j) { double t = m[c][j]; m[c][j] = m[p][j]; m[p][j] = t; }
        double d = 1.0 / m[c][c];
        for (int j = 0; j < 2 * N; ++j) m[c][j] *= d;
        for (int r = 0; r < N; ++r) {
            if (r == c) continue;
            double f = m[r][c];
            if (f != 0.0)
                for (int j = 0; j < 2 * N; ++j) m[r][j] -= f * m[c][j];
        }
    }
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) inv[i * N + j] = m[i][N + j];
    return true;
}

template <int N>
__device__ void matmul(const double* a, const double* b, double* c) {
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < N; ++j) {
            double s = 0.0;
            for (int k = 0; k < N; ++k) s += a[i * N + k] * b[k * N + j];
            c[i * N + j] = s;
        }
}

// cam = [2,4,4] (extrinsic, intrinsic) -> 4x4 projection with rows 0..2 = K @ E[:3,:4]
__device__ void compose(const float* cam, double* P) {
    const float* E = cam;
    const float* K = cam + 16;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 4; ++j) {
            double s = 0.0;
            for (int k = 0; k < 3; ++k) s += (double)K[i * 4 + k] * (double)E[k * 4 + j];
            P[i * 4 + j] = s;
        }
    for (int j = 0; j < 4; ++j) P[12 + j] = E[12 + j];
}

__device__ void write_coef(const double* src_P, const double* ref_P, float* coef) {
    double inv[16], M[16];
    if (!invert<4>(ref_P, inv)) {
        for (int i = 0; i < 12; ++i) coef[i] = __int_as_float(0x7fc00000);
        return;
    }
    matmul<4>(src_P, inv, M);
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) coef[i * 3 + j] = (float)M[i * 4 + j];
        coef[9 + i] = (float)M[i * 4 + 3];
    }
}

__global__ void warp_coeffs_kernel(const float* __restrict__ src_proj, const float* __restrict__ ref_proj, int B,
                                   float* __restrict__ coef) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double s[16], r[16];
    for (int i = 0; i < 16; ++i) { s[i] = src_proj[b * 16 + i]; r[i] = ref_proj[b * 16 + i]; }
    write_coef(s, r, coef + b * 12);
}

// epipole of F: rows c*F0 +- (F1+F2), 2x2 solve (dynamic_conv.py:41-47)
__device__ void epipole(const double* F, bool transpose, float* e) {
    double G[9];
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) G[i * 3 + j] = transpose ? F[j * 3 + i] : F[i * 3 + j];
    const double c = 1e3;
    double r1[3], r2[3];
    for (int j = 0; j < 3; ++j) {
        r1[j] = c * G[j] + G[3 + j] + G[6 + j];
        r2[j] = c * G[j] - G[3 + j] - G[6 + j];
    }
    double A[4] = {r1[0], r1[1], r2[0], r2[1]}, Ai[4];
    if (!invert<2>(A, Ai)) { e[0] = e[1] = __int_as_float(0x7fc00000); return; }
    e[0] = (float)(-(Ai[0] * r1[2] + Ai[1] * r2[2]));
    e[1] = (float)(-(Ai[2] * r1[2] + Ai[3] * r2[2]));
}

struct StagePtrs {
    const float* p[4];
};

// one thread per (b, source view): warp coefficients for every stage + the two epipoles
__global__ void camera_setup_kernel(StagePtrs proj, int n_stages, int epi_stage, int B, int N,
                                    float* __restrict__ coef, float* __restrict__ epipoles) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    int V = N - 1;
    if (t >= B * V) return;
    int b = t / V, v = t % V;
    for (int s = 0; s < n_stages; ++s) {
        const float* ref = proj.p[s] + ((size_t)b * N + 0) * 32;
        const float* src = proj.p[s] + ((size_t)b * N + (v + 1)) * 32;
        double Pr[16], Ps[16];
        compose(ref, Pr);
        compose(src, Ps);
        write_coef(Ps, Pr, coef + (((size_t)s * B + b) * V + v) * 12);
    }
    if (epipoles) {
        const float* c1 = proj.p[epi_stage] + ((size_t)b * N + 0) * 32;
        const float* c2 = proj.p[epi_stage] + ((size_t)b * N + (v + 1)) * 32;
        double R1[9], R2[9], K1[9], K2[9], t1[3], t2[3];
        for (int i = 0; i < 3; ++i) {
            for (int j = 0; j < 3; ++j) {
                R1[i * 3 + j] = c1[i * 4 + j];
                R2[i * 3 + j] = c2[i * 4 + j];
                K1[i * 3 + j] = c1[16 + i * 4 + j];
                K2[i * 3 + j] = c2[16 + i * 4 + j];
            }
            t1[i] = c1[i * 4 + 3];
            t2[i] = c2[i * 4 + 3];
        }
        double R1i[9], R2i[9], P1[9], P2[9], P1i[9];
        // layout [side (0 = ref image, 1 = src image)][v][b][2]: the order of the feature-extractor batch
        float* e = epipoles + ((size_t)v * B + b) * 2;
        float* e2 = e + (size_t)V * B * 2;
        bool ok = invert<3>(R1, R1i) && invert<3>(R2, R2i);
        matmul<3>(K1, R1, P1);
        matmul<3>(K2, R2, P2);
        ok = ok && invert<3>(P1, P1i);
        if (!ok) { e[0] = e[1] = e2[0] = e2[1] = __int_as_float(0x7fc00000); return; }
        double cd[3], e12[3];
        for (int i = 0; i < 3; ++i) {
            double a = 0, c = 0;
            for (int k = 0; k < 3; ++k) { a += R1i[i * 3 + k] * t1[k]; c += R2i[i * 3 + k] * t2[k]; }
            cd[i] = -a + c;  // c1 - c2, with c_i = -R_i^-1 t_i
        }
        for (int i = 0; i < 3; ++i) e12[i] = P2[i * 3] * cd[0] + P2[i * 3 + 1] * cd[1] + P2[i * 3 + 2] * cd[2];
        double S[9] = {0, -e12[2], e12[1], e12[2], 0, -e12[0], -e12[1], e12[0], 0};
        double SP[9], F[9];
        matmul<3>(S, P2, SP);
        matmul<3>(SP, P1i, F);
        epipole(F, false, e);   // epipole in the reference image
        epipole(F, true, e2);   // epipole in the source image
    }
}

// ------------------------------------------------------------------------------------------
// A1: materialised warp, exact reference contract (fp32 NCHW in, fp32 NCDHW out)
// ------------------------------------------------------------------------------------------
__global__ void homo_warp_kernel(const float* __restrict__ src, const float* __restrict__ coef,
                                 const float* __restrict__ depth, int per_pixel, int B, int C, int D, int h, int w,
                                 float* __restrict__ out) {
    long long P = (long long)h * w;
    long long total = (long long)B * D * P;
    float half_w = (float)((double)(w - 1) / 2.0), half_h = (float)((double)(h - 1) / 2.0);
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int x = (int)(i % w);
        int y = (int)((i / w) % h);
        int d = (int)((i / P) % D);
        int b = (int)(i / (P * D));
        WarpCoef k = load_coef(coef + b * 12);
        float dep = per_pixel ? __ldg(depth + ((size_t)b * D + d) * P + (size_t)y * w + x) : __ldg(depth + b * D + d);
        float rx, ry, rz, u, v;
        pixel_ray(k, (float)x, (float)y, rx, ry, rz);
        project(k, rx, ry, rz, dep, u, v);
        // the reference normalises to [-1,1] and grid_sample maps back; keep that round trip
        u = ((u / half_w - 1.f) + 1.f) / 2.f * (float)(w - 1);
        v = ((v / half_h - 1.f) + 1.f) / 2.f * (float)(h - 1);
        Taps t = make_taps(u, v, w, h);
        int xa = min(max(t.x0, 0), w - 1), xb = min(max(t.x0 + 1, 0), w - 1);
        int ya = min(max(t.y0, 0), h - 1), yb = min(max(t.y0 + 1, 0), h - 1);
        const float* sb = src + (size_t)b * C * P;
        float* ob = out + ((size_t)b * C * D + d) * P + (size_t)y * w + x;
        for (int c = 0; c < C; ++c) {
            const float* sc = sb + (size_t)c * P;
            float val = t.w00 * __ldg(sc + (size_t)ya * w + xa) + t.w01 * __ldg(sc + (size_t)ya * w + xb) +
                        t.w10 * __ldg(sc + (size_t)yb * w + xa) + t.w11 * __ldg(sc + (size_t)yb * w + xb);
            ob[(size_t)c * D * P] = val;
        }
    }
}

// ------------------------------------------------------------------------------------------
// A9: depth hypotheses at the stage resolution
// ------------------------------------------------------------------------------------------
// bilinear up-sample with half-pixel centres (F.interpolate(..., align_corners=False))
__device__ __forceinline__ float upsample_at(const float* __restrict__ prev, int hp, int wp, float sy, float sx, int Y, int X) {
    float cy = fmaxf(sy * ((float)Y + 0.5f) - 0.5f, 0.f);
    float cx = fmaxf(sx * ((float)X + 0.5f) - 0.5f, 0.f);
    int y0 = min((int)cy, hp - 1), x0 = min((int)cx, wp - 1);
    int y1 = min(y0 + 1, hp - 1), x1 = min(x0 + 1, wp - 1);
    float ly = cy - (float)y0, lx = cx - (float)x0;
    float a = __ldg(prev + (size_t)y0 * wp + x0), b = __ldg(prev + (size_t)y0 * wp + x1);
    float c = __ldg(prev + (size_t)y1 * wp + x0), d = __ldg(prev + (size_t)y1 * wp + x1);
    return (1.f - ly) * ((1.f - lx) * a + lx * b) + ly * ((1.f - lx) * c + lx * d);
}

__global__ void hypotheses_kernel(const float* __restrict__ depth_values, int Dtot, const float* __restrict__ prev,
                                  int hp, int wp, int B, int D, float ratio, int H, int W, int scale,
                                  float* __restrict__ out) {
    int h = H / scale, w = W / scale;
    long long P = (long long)h * w;
    long long total = (long long)B * P;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int x = (int)(i % w), y = (int)((i / w) % h), b = (int)(i / P);
        const float* dv = depth_values + (size_t)b * Dtot;
        float dmin = __ldg(dv), dmax = __ldg(dv + Dtot - 1);
        float* o = out + (size_t)b * D * P + (size_t)y * w + x;
        if (prev == nullptr) {
            float step = (dmax - dmin) / (float)(D - 1);
            for (int d = 0; d < D; ++d) o[(size_t)d * P] = dmin + (float)d * step;
            continue;
        }
        float step = ratio * (__ldg(dv + 1) - dmin);
        int nl = (D - 1) / 2;
        const float* pb = prev + (size_t)b * hp * wp;
        float sy = (float)hp / (float)H, sx = (float)wp / (float)W;
        // centre taps of the scale-wide block per axis (2 per axis, coincident when scale == 1)
        int oa = scale / 2 - 1, ob = scale / 2;
        if (scale == 1) oa = ob = 0;
        float cur[4];
        cur[0] = upsample_at(pb, hp, wp, sy, sx, y * scale + oa, x * scale + oa);
        if (scale > 1) {
            cur[1] = upsample_at(pb, hp, wp, sy, sx, y * scale + oa, x * scale + ob);
            cur[2] = upsample_at(pb, hp, wp, sy, sx, y * scale + ob, x * scale + oa);
            cur[3] = upsample_at(pb, hp, wp, sy, sx, y * scale + ob, x * scale + ob);
        }
        for (int d = 0; d < D; ++d) {
            float s[4];
            int n = scale > 1 ? 4 : 1;
            for (int j = 0; j < n; ++j) {
                float v = (cur[j] - (float)nl * step) + (float)d * step;
                v = dmin + fmaxf(v - dmin, 0.f);
                v = dmax + fminf(v - dmax, 0.f);
                s[j] = v;
            }
            float r = s[0];
            if (scale > 1) r = 0.5f * (0.5f * s[0] + 0.5f * s[1]) + 0.5f * (0.5f * s[2] + 0.5f * s[3]);
            o[(size_t)d * P] = r;
        }
    }
}

}  // namespace

extern "C" {

int cds_warp_coeffs(const float* src_proj, const float* ref_proj, int B, float* coef, cudaStream_t stream) {
    CDS_REQUIRE(src_proj && ref_proj && coef && B > 0, CDS_EARG, "cds_warp_coeffs: null pointer or B <= 0");
    warp_coeffs_kernel<<<cds_div_up(B, 64), 64, 0, stream>>>(src_proj, ref_proj, B, coef);
    return cds_check_launch("cds_warp_coeffs");
}

int cds_camera_setup(const float* const* proj_stages, int n_stages, int epi_stage, int B, int N, float* coef,
                     float* epipoles, cudaStream_t stream) {
    CDS_REQUIRE(proj_stages && coef, CDS_EARG, "cds_camera_setup: null pointer");
    CDS_REQUIRE(n_stages >= 1 && n_stages <= 4 && epi_stage >= 0 && epi_stage < n_stages, CDS_EARG,
                "cds_camera_setup: n_stages must be 1..4 and epi_stage inside it");
    CDS_REQUIRE(B > 0 && N >= 2, CDS_ESHAPE, "cds_camera_setup: need B > 0 and at least one source view");
    StagePtrs p{};
    for (int s = 0; s < n_stages; ++s) {
        CDS_REQUIRE(proj_stages[s], CDS_EARG, "cds_camera_setup: null stage pointer");
        p.p[s] = proj_stages[s];
    }
    int n = B * (N - 1);
    camera_setup_kernel<<<cds_div_up(n, 64), 64, 0, stream>>>(p, n_stages, epi_stage, B, N, coef, epipoles);
    return cds_check_launch("cds_camera_setup");
}

int cds_homo_warp(const float* src_fea, const float* coef, const float* depth, int depth_per_pixel, int B, int C, int D,
                  int h, int w, float* out, cudaStream_t stream) {
    CDS_REQUIRE(src_fea && coef && depth && out, CDS_EARG, "cds_homo_warp: null pointer");
    CDS_REQUIRE(B > 0 && C > 0 && D > 0 && h > 1 && w > 1, CDS_ESHAPE, "cds_homo_warp: bad shape B=%d C=%d D=%d h=%d w=%d", B, C, D, h, w);
    long long total = (long long)B * D * h * w;
    int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
    homo_warp_kernel<<<blocks, 256, 0, stream>>>(src_fea, coef, depth, depth_per_pixel, B, C, D, h, w, out);
    return cds_check_launch("cds_homo_warp");
}

int cds_depth_hypotheses(const float* depth_values, int Dtot, const float* prev_depth, int hp, int wp, int B, int D,
                         float ratio, int H, int W, int scale, float* out, cudaStream_t stream) {
    CDS_REQUIRE(depth_values && out, CDS_EARG, "cds_depth_hypotheses: null pointer");
    CDS_REQUIRE(Dtot >= 2 && D >= 2 && B > 0, CDS_ESHAPE, "cds_depth_hypotheses: need Dtot >= 2, D >= 2");
    CDS_REQUIRE((scale == 1 || scale == 2 || scale == 4) && H % scale == 0 && W % scale == 0, CDS_ESHAPE,
                "cds_depth_hypotheses: scale must be 1, 2 or 4 and divide H, W");
    CDS_REQUIRE(prev_depth == nullptr || (hp > 0 && wp > 0), CDS_ESHAPE, "cds_depth_hypotheses: bad previous-depth shape");
    long long total = (long long)B * (H / scale) * (W / scale);
    int blocks = (int)min((long long)148 * 8, (total + 255) / 256);
    hypotheses_kernel<<<blocks, 256, 0, stream>>>(depth_values, Dtot, prev_depth, hp, wp, B, D, ratio, H, W, scale, out);
    return cds_check_launch("cds_depth_hypotheses");
}

}  // extern "C"
