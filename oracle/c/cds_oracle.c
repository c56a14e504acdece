/* Plain-C restatement of the third-party operators the oracle delegates to (TEST INFRASTRUCTURE ONLY).
 *
 * oracle/oracle.py computes the dense arithmetic of the path with the torch (ATen, CPU, fp32) operators the
 * reference itself calls -- conv2d / conv3d / conv_transpose3d / grid_sample / softmax -- because that code is a
 * third-party dependency of the reference (PyTorch, pinned at 1.6.0 in the reference's README.md:15) and is not
 * vendored in its tree.  This file restates the published definitions of those operators as direct loops so
 * that tests/test_oracle_c.py can check "ATen == the definition" on small cases; it is never linked into the
 * product.  Layouts are the reference's: NCHW / NCDHW, fp32, contiguous.
 *
 *   cds_c_conv2d            models/dynamic_conv.py:85-86,112,116 ; models/module.py:49,191 (nn.Conv2d)
 *   cds_c_conv3d            models/module.py:102 (nn.Conv3d k3, stride 1|2, pad 1)
 *   cds_c_conv_transpose3d  models/module.py:146 (nn.ConvTranspose3d k3 s2 p1 op1)
 *   cds_c_bilinear_zeros    models/utils/warping.py:100-101 (grid_sample bilinear / zeros / align_corners=True,
 *                           addressed directly in pixel coordinates)
 *   cds_c_softmax_regress   models/model.py:90-92 ; models/module.py:373-391
 */
#include <math.h>
#include <stdlib.h>
#include <stddef.h>

/* y[n,co,ho,wo] = b[co] + sum_{ci,ky,kx} x[n,ci,ho*s-p+ky,wo*s-p+kx] * w[co,ci,ky,kx] */
void cds_c_conv2d(const float* x, const float* w, const float* b, int N, int Ci, int H, int W, int Co, int k, int stride,
                  int pad, float* y) {
    int Ho = (H + 2 * pad - k) / stride + 1, Wo = (W + 2 * pad - k) / stride + 1;
    for (int n = 0; n < N; ++n)
        for (int co = 0; co < Co; ++co)
            for (int ho = 0; ho < Ho; ++ho)
                for (int wo = 0; wo < Wo; ++wo) {
                    double acc = b ? b[co] : 0.0;
                    for (int ci = 0; ci < Ci; ++ci)
                        for (int ky = 0; ky < k; ++ky) {
                            int iy = ho * stride - pad + ky;
                            if (iy < 0 || iy >= H) continue;
                            for (int kx = 0; kx < k; ++kx) {
                                int ix = wo * stride - pad + kx;
                                if (ix < 0 || ix >= W) continue;
                                acc += (double)x[(((size_t)n * Ci + ci) * H + iy) * W + ix] *
                                       (double)w[(((size_t)co * Ci + ci) * k + ky) * k + kx];
                            }
                        }
                    y[(((size_t)n * Co + co) * Ho + ho) * Wo + wo] = (float)acc;
                }
}

/* k = 3, pad = 1; output size ceil(n / stride) per axis */
void cds_c_conv3d(const float* x, const float* w, int N, int Ci, int D, int H, int W, int Co, int stride, float* y) {
    int Do = (D - 1) / stride + 1, Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
    for (int n = 0; n < N; ++n)
        for (int co = 0; co < Co; ++co)
            for (int od = 0; od < Do; ++od)
                for (int oh = 0; oh < Ho; ++oh)
                    for (int ow = 0; ow < Wo; ++ow) {
                        double acc = 0.0;
                        for (int ci = 0; ci < Ci; ++ci)
                            for (int kd = 0; kd < 3; ++kd)
                                for (int kh = 0; kh < 3; ++kh)
                                    for (int kw = 0; kw < 3; ++kw) {
                                        int id = od * stride - 1 + kd, ih = oh * stride - 1 + kh, iw = ow * stride - 1 + kw;
                                        if (id < 0 || id >= D || ih < 0 || ih >= H || iw < 0 || iw >= W) continue;
                                        acc += (double)x[((((size_t)n * Ci + ci) * D + id) * H + ih) * W + iw] *
                                               (double)w[((((size_t)co * Ci + ci) * 3 + kd) * 3 + kh) * 3 + kw];
                                    }
                        y[((((size_t)n * Co + co) * Do + od) * Ho + oh) * Wo + ow] = (float)acc;
                    }
}

/* k = 3, stride 2, pad 1, output_padding 1: y[2i - 1 + k] += x[i] * w[ci,co,k] per axis; output size 2n. y must be zeroed. */
void cds_c_conv_transpose3d(const float* x, const float* w, int N, int Ci, int D, int H, int W, int Co, float* y) {
    int Do = 2 * D, Ho = 2 * H, Wo = 2 * W;
    for (int n = 0; n < N; ++n)
        for (int ci = 0; ci < Ci; ++ci)
            for (int id = 0; id < D; ++id)
                for (int ih = 0; ih < H; ++ih)
                    for (int iw = 0; iw < W; ++iw) {
                        float v = x[((((size_t)n * Ci + ci) * D + id) * H + ih) * W + iw];
                        for (int co = 0; co < Co; ++co)
                            for (int kd = 0; kd < 3; ++kd)
                                for (int kh = 0; kh < 3; ++kh)
                                    for (int kw = 0; kw < 3; ++kw) {
                                        int od = 2 * id - 1 + kd, oh = 2 * ih - 1 + kh, ow = 2 * iw - 1 + kw;
                                        if (od < 0 || od >= Do || oh < 0 || oh >= Ho || ow < 0 || ow >= Wo) continue;
                                        y[((((size_t)n * Co + co) * Do + od) * Ho + oh) * Wo + ow] +=
                                            v * w[((((size_t)ci * Co + co) * 3 + kd) * 3 + kh) * 3 + kw];
                                    }
                    }
}

/* out[c,m] = bilinear sample of fea[c,:,:] at pixel coordinates (u[m], v[m]); taps outside the image contribute 0 */
void cds_c_bilinear_zeros(const float* fea, int C, int h, int w, const float* u, const float* v, int M, float* out) {
    for (int m = 0; m < M; ++m) {
        float x0f = floorf(u[m]), y0f = floorf(v[m]);
        float fx = u[m] - x0f, fy = v[m] - y0f;
        int x0 = (int)x0f, y0 = (int)y0f;
        for (int c = 0; c < C; ++c) {
            float acc = 0.f;
            for (int dy = 0; dy < 2; ++dy)
                for (int dx = 0; dx < 2; ++dx) {
                    int xi = x0 + dx, yi = y0 + dy;
                    if (xi < 0 || xi >= w || yi < 0 || yi >= h) continue;
                    float wgt = (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy);
                    acc += wgt * fea[((size_t)c * h + yi) * w + xi];
                }
            out[(size_t)c * M + m] = acc;
        }
    }
}

/* Adjoint of cds_c_bilinear_zeros in fea: grad[c,m] at (u[m], v[m]) is added, with the same weights, to the in-image taps of
 * grad_fea[c,:,:] (zeroed here).  What autograd derives for grid_sample(bilinear, zeros, align_corners=True) when only the
 * input carries a gradient (models/utils/warping.py:79,100-101).  fp64 accumulation. */
void cds_c_bilinear_zeros_backward(const float* grad, int C, int h, int w, const float* u, const float* v, int M, float* grad_fea) {
    double* acc = (double*)calloc((size_t)C * h * w, sizeof(double));
    for (int m = 0; m < M; ++m) {
        float x0f = floorf(u[m]), y0f = floorf(v[m]);
        float fx = u[m] - x0f, fy = v[m] - y0f;
        int x0 = (int)x0f, y0 = (int)y0f;
        for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx) {
                int xi = x0 + dx, yi = y0 + dy;
                if (xi < 0 || xi >= w || yi < 0 || yi >= h) continue;
                float wgt = (dx ? fx : 1.f - fx) * (dy ? fy : 1.f - fy);
                for (int c = 0; c < C; ++c) acc[((size_t)c * h + yi) * w + xi] += (double)(wgt * grad[(size_t)c * M + m]);
            }
    }
    for (size_t i = 0; i < (size_t)C * h * w; ++i) grad_fea[i] = (float)acc[i];
    free(acc);
}

/* One stage of final_loss (models/losses.py:14-23,25-35) on flat arrays of n pixels (interval already per pixel):
 * out[0] = smooth-L1 mean of est/iv - gt/iv over mask > 0.5, out[1] = masked mean of curv. */
void cds_c_stage_loss(const float* est, const float* gt, const float* mask, const float* iv, const float* curv, long n, float* out) {
    double l = 0.0, c = 0.0, cnt = 0.0;
    for (long i = 0; i < n; ++i) {
        if (!(mask[i] > 0.5f)) continue;
        float d = est[i] / iv[i] - gt[i] / iv[i], a = fabsf(d);
        l += a < 1.f ? 0.5f * d * d : a - 0.5f;
        c += curv[i];
        cnt += 1.0;
    }
    out[0] = (float)(l / cnt);
    out[1] = (float)(c / cnt);
}

/* logits [D,P] (one batch item), depth [D,P]: softmax over D, expectation depth, 4-plane confidence window */
void cds_c_softmax_regress(const float* logits, const float* depth, int D, int P, float* depth_out, float* conf_out) {
    for (int p = 0; p < P; ++p) {
        float m = -INFINITY;
        for (int d = 0; d < D; ++d) m = fmaxf(m, logits[(size_t)d * P + p]);
        double S = 0.0;
        for (int d = 0; d < D; ++d) S += exp((double)logits[(size_t)d * P + p] - m);
        double ed = 0.0, ei = 0.0;
        for (int d = 0; d < D; ++d) {
            double pr = exp((double)logits[(size_t)d * P + p] - m) / S;
            ed += pr * depth[(size_t)d * P + p];
            ei += pr * d;
        }
        int idx = (int)ei;
        if (idx < 0) idx = 0;
        if (idx > D - 1) idx = D - 1;
        double c = 0.0;
        for (int j = idx - 1; j <= idx + 2; ++j)
            if (j >= 0 && j < D) c += exp((double)logits[(size_t)j * P + p] - m) / S;
        depth_out[p] = (float)ed;
        conf_out[p] = (float)c;
    }
}
