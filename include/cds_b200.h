/* cds_b200.h -- C ABI of libcds_b200.so: the CDS-MVSNet depth-inference hot path on B200 (sm_100a).
 *
 * The reference (TruongKhang/cds-mvsnet) has no FFI/plugin layer: its hot path is Python calling
 * ATen/cuDNN.  This header is the boundary a replacement binds instead -- plain pointers, sizes and
 * a CUDA stream, no torch types.  Each entry names the reference code it replaces (file:line into
 * the reference tree).  The Python shims in cds_mvsnet_b200/ (ctypes) are the intended callers;
 * INTEGRATION.md shows the reference-side binding.
 *
 * Conventions
 *   - every function returns 0 on success, a positive cudaError_t, or a negative CDS_E* argument
 *     code; cds_last_error_string() describes the last failure on the calling thread
 *   - all pointers are DEVICE pointers unless stated; the caller owns every buffer, nothing is
 *     allocated or freed inside, inputs are never written
 *   - work is enqueued on `stream` and is CUDA-graph capturable (no host sync inside)
 *   - thread-safe / re-entrant: no mutable global state
 *   - 2-D activation tensors are channels-last [n,H,W,C]; 3-D ones are channel-BLOCKED channels-last
 *     [B,C/8,D,H,W,8] (plain NDHWC when C = 8) so an 8-channel slab of a voxel row is contiguous -- the unit
 *     the TMA/tcgen05 kernels consume; all in the storage type `dtype`
 *     (CDS_F16: fp16 storage with fp32 accumulation -- the production setting; CDS_F32: fp32
 *     storage for tight parity checks); per-pixel maps, hypotheses and outputs are fp32
 */
#ifndef CDS_B200_H
#define CDS_B200_H

#include <cuda_runtime_api.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CDS_F32 0
#define CDS_F16 1

#define CDS_ACT_NONE 0
#define CDS_ACT_LRELU 1 /* LeakyReLU(0.1), models/module.py:69 */
#define CDS_ACT_TANH 2  /* models/module.py:223,230,232 */

#define CDS_EARG (-1)
#define CDS_ESHAPE (-2)
#define CDS_EUNSUPPORTED (-3)

int cds_version(void);
const char* cds_last_error_string(void);
/* 0 iff the current device is sm_100 (the only target this library is built for). */
int cds_check_device(void);

/* ---- camera algebra ------------------------------------------------------------------------ */
/* coef[b] = {R row-major 9, t 3} of src_proj[b] @ inv(ref_proj[b]); 4x4 row-major fp32 inputs.
 * Replaces models/utils/warping.py:80-82 (torch.inverse + matmul), evaluated in fp64. */
int cds_warp_coeffs(const float* src_proj, const float* ref_proj, int B, float* coef, cudaStream_t stream);

/* For the whole cascade at once: proj_stages is a HOST array of n_stages device pointers to the data
 * layer's [B,N,2,4,4] matrices.  Writes coef [n_stages,B,N-1,12] (projection composed as
 * K[:3,:3] @ E[:3,:4], models/model.py:40-43, then the warp coefficients above) and, if non-null,
 * epipoles [2,N-1,B,2] = xy of the epipole in (0: the ref image, 1: the src image) from stage `epi_stage`'s cameras
 * (models/dynamic_conv.py:19-47 compute_Fmatrix/compute_epipole, called at models/model.py:152-158). */
int cds_camera_setup(const float* const* proj_stages, int n_stages, int epi_stage, int B, int N, float* coef,
                     float* epipoles, cudaStream_t stream);

/* ---- A1: plane-sweep warp, exact reference contract ---------------------------------------- */
/* homo_warping_3D (models/utils/warping.py:69-104) with the 4x4 algebra already reduced to coef:
 * src_fea [B,C,h,w] fp32 NCHW, depth [B,D] (depth_per_pixel=0) or [B,D,h,w] (=1),
 * out [B,C,D,h,w] fp32.  Bilinear, zero padding, align_corners=True pixel coordinates. */
int cds_homo_warp(const float* src_fea, const float* coef, const float* depth, int depth_per_pixel, int B, int C, int D,
                  int h, int w, float* out, cudaStream_t stream);

/* ---- A9: depth hypotheses ------------------------------------------------------------------- */
/* depth_values [B,Dtot].  prev_depth == NULL: D planes uniformly spanning [dv[0], dv[-1]]
 * (models/module.py:425-433).  Otherwise prev_depth [B,hp,wp] is bilinearly up-sampled to (H,W)
 * (models/model.py:177-182), hypotheses cur - ((D-1)/2)*step + d*step with
 * step = ratio * (dv[1]-dv[0]) are clamped to [dv[0], dv[-1]] (models/module.py:398-417) and resized
 * to the stage grid (models/model.py:191-193).  out [B,D,H/scale,W/scale] fp32, scale in {1,2,4}. */
int cds_depth_hypotheses(const float* depth_values, int Dtot, const float* prev_depth, int hp, int wp, int B, int D,
                         float ratio, int H, int W, int scale, float* out, cudaStream_t stream);

/* ---- A2: fused warp + cost volume (two sweeps around the visibility net) --------------------- */
/* ref_fea/src_fea [V,B,h,w,C] channels-last (one ref map per source view: the ref features depend on
 * the pair's epipole, models/model.py:154-161), coef [B,V,12], depth [B,D,h,w].
 * entropy[v,b,y,x] = H(softmax_d(sum_c ref*warp_d))  (models/model.py:44-50). C in {8,16,32}. */
int cds_costvol_entropy(const void* ref_fea, const void* src_fea, const float* coef, const float* depth, int V, int B,
                        int C, int D, int h, int w, int dtype, float* entropy, cudaStream_t stream);
/* Same sweep for a consumer that only needs the entropy to ~5e-4 (the visibility net, models/model.py:51): on fp16 features the
 * four taps are blended in packed half arithmetic before the dot product (cds_costvol_entropy keeps fp32 tap weights: < 2e-4).
 * Other dtypes and the TMA-staged last-stage shape run cds_costvol_entropy itself. */
int cds_costvol_entropy_fast(const void* ref_fea, const void* src_fea, const float* coef, const float* depth, int V, int B,
                             int C, int D, int h, int w, int dtype, float* entropy, cudaStream_t stream);
/* volume[b,d,y,x,:] = sum_v vis_v * ref_v (.) warp_{v,d} / (sum_v vis_v + 1e-6), vis [V,B,h,w]
 * (models/model.py:57-59,74); volume [B,C/8,D,h,w,8] channel-blocked. */
int cds_costvol_aggregate(const void* ref_fea, const void* src_fea, const float* coef, const float* depth,
                          const float* vis, int V, int B, int C, int D, int h, int w, int dtype, void* volume,
                          cudaStream_t stream);
/* out[i] = sum_v (ref_nc_sum[v,i] + src_nc_sum[v,i]) / 2 / V   (models/model.py:60,79) */
/* cds_costvol_aggregate from fp32 features to a split-precision fp16 volume: vol_hi = fp16 value plane, vol_lo (optional) = its
 * fp16 rounding residual plane, both [B,C/8,D,h,w,8] (stage 1 of the cascade, whose rounding the later stages amplify). */
int cds_costvol_aggregate_split(const float* ref_fea, const float* src_fea, const float* coef, const float* depth, const float* vis,
                                int V, int B, int C, int D, int h, int w, void* vol_hi, void* vol_lo, cudaStream_t stream);
int cds_nc_mean(const float* ref_nc_sum, const float* src_nc_sum, int V, long long n, float* out, cudaStream_t stream);

/* ---- A3: visibility net ----------------------------------------------------------------------- */
/* StageNet.vis[s] (models/model.py:14,51; ConvBnReLU models/module.py:169-198), BN folded, as one
 * kernel.  entropy, curv, vis: [n,h,w] fp32.  wpack: cds_visnet_weight_floats() floats laid out as
 * L1[9][2][16] b1[16] L2[9][16][16] b2[16] L3[9][16][16] b3[16] w4[16] b4[1]. */
int cds_visnet_weight_floats(void);
int cds_visnet(const float* entropy, const float* curv, const float* wpack, int n, int h, int w, float* vis,
               cudaStream_t stream);

/* Tensor-core form (three tcgen05 tap-GEMM layers chained through shared memory), w >= 128.  wgt_packed: fp16 operand
 * image (cds_visnet_tc_weight_halfs() halfs); fparams: b1[16] b2[16] b3[16] w4[16] b4[1] fp32 (BN folded). */
int cds_visnet_tc_supported(int h, int w);
int cds_visnet_tc_weight_halfs(void);
int cds_visnet_tc(const float* entropy, const float* curv, const void* wgt_packed, const float* fparams, int n, int h, int w,
                  float* vis, cudaStream_t stream);

/* ---- A4: 3-D regulariser blocks --------------------------------------------------------------- */
/* Conv3d block (models/module.py:80-122): k3 p1, stride 1|2, BN folded into wgt [27][Cin][Cout] fp32
 * (tap = (kd*3+kh)*3+kw) and bias [Cout], optional ReLU.  in [B,Cin/8,D,H,W,8] -> out [B,Cout/8,ceil(D/s),..,8]. */
int cds_conv3d_k3(const void* in, const float* wgt, const float* bias, int B, int Cin, int Cout, int D, int H, int W,
                  int stride, int relu, int dtype, void* out, cudaStream_t stream);
/* Deconv3d block (models/module.py:125-166): ConvTranspose3d k3 s2 p1 op1 + BN + ReLU, then `+ skip`
 * (models/module.py:310-312; skip may be NULL).  wgt [27][Cin][Cout]; out/skip [B,Cout/8,2D,2H,2W,8]. */
int cds_deconv3d_k3s2(const void* in, const float* wgt, const float* bias, const void* skip, int B, int Cin, int Cout,
                      int D, int H, int W, int dtype, void* out, cudaStream_t stream);
/* Tensor-core (tcgen05 + TMEM) form of the stride-1 Conv3d block, fp16 storage only.  Same semantics as
 * cds_conv3d_k3; wgt_packed is the fp16 operand image built by the host (cds_conv3d_k3_tc_weight_halfs()
 * halfs, layout [mma][k-chunk 2][Npad/8][8 n][8 k], see csrc/conv3d_tc.cu).  cds_conv3d_k3_tc_supported()
 * tells whether a layer shape is covered (stride 1, W >= 8, channel pairs of the regulariser). */
int cds_conv3d_k3_tc_supported(int Cin, int Cout, int D, int H, int W, int stride);
int cds_conv3d_k3_tc_weight_halfs(int Cin, int Cout);
int cds_conv3d_k3_tc(const void* in, const void* wgt_packed, const float* bias, int B, int Cin, int Cout, int D, int H,
                     int W, int relu, void* out, cudaStream_t stream);
/* Tensor-core form of the Deconv3d block (same semantics as cds_deconv3d_k3s2, fp16 storage, input W >= 8).
 * wgt_packed: fp16 operand image (cds_deconv3d_k3s2_tc_weight_halfs() halfs, layout in csrc/conv3d_tc.cu). */
int cds_deconv3d_k3s2_tc_supported(int Cin, int Cout, int D, int H, int W);
int cds_deconv3d_k3s2_tc_weight_halfs(int Cin, int Cout);
int cds_deconv3d_k3s2_tc(const void* in, const void* wgt_packed, const float* bias, const void* skip, int B, int Cin, int Cout,
                         int D, int H, int W, void* out, cudaStream_t stream);
/* Gather-form tensor-core Conv3d / Deconv3d blocks (csrc/conv3d_gtc.cu): same semantics as cds_conv3d_k3 (stride 1|2) and
 * cds_deconv3d_k3s2, fp16 storage, any D/H/W; operands are gathered voxel by voxel (cp.async) so stride 2 and small deep
 * volumes are covered.  wgt_packed: fp16 image [mma][k-chunk 2][2*Cout/8][8 n][8 k] whose columns [Cout, 2*Cout) hold the
 * fp16 rounding residual of the folded weights (host: weights.py pack_conv3d_gtc / pack_deconv3d_gtc). */
int cds_conv3d_k3_gtc_supported(int Cin, int Cout, int stride);
int cds_conv3d_k3_gtc_weight_halfs(int Cin, int Cout);
int cds_conv3d_k3_gtc(const void* in, const void* wgt_packed, const float* bias, int B, int Cin, int Cout, int D, int H, int W,
                      int stride, int relu, void* out, cudaStream_t stream);
int cds_deconv3d_k3s2_gtc_supported(int Cin, int Cout);
int cds_deconv3d_k3s2_gtc_weight_halfs(int Cin, int Cout);
int cds_deconv3d_k3s2_gtc(const void* in, const void* wgt_packed, const float* bias, const void* skip, int B, int Cin, int Cout,
                          int D, int H, int W, void* out, cudaStream_t stream);
/* Persistent d-rolling tcgen05 form of the full-resolution stride-1 layers (csrc/conv3d_roll.cu): Cout == 8 (conv0, same
 * semantics as cds_conv3d_k3, fp16 storage) and Cin == 8, Cout == 1 (the prob head: out = fp32 logits [B,D,H,W], bias NULL).
 * wgt_packed: fp16 image [kd][mma][k-chunk 2][N/8][8 n][8 k], kw folded into N (host: weights.py pack_conv3d_roll). */
int cds_conv3d_k3_roll_supported(int Cin, int Cout, int D, int H, int W);
int cds_conv3d_k3_roll_weight_halfs(int Cin, int Cout);
int cds_conv3d_k3_roll(const void* in, const void* wgt_packed, const float* bias, int B, int Cin, int Cout, int D, int H, int W,
                       int relu, void* out, cudaStream_t stream);
/* The regulariser's tail in ONE kernel: prob head (models/module.py:303) with softmax over D, depth_regression and
 * conf_regression (models/model.py:85-92, models/module.py:373-391) fused into its epilogue.  in [B,D,H,W,8] fp16 (conv11's
 * output), samples [B,D,H,W] fp32 per-pixel hypotheses -> depth, conf [B,H,W] fp32; logits [B,D,H,W] fp32 is still written
 * (the confidence re-reads the four around the expected index).  Same results as cds_conv3d_k3_roll + cds_softmax_regress. */
int cds_prob_head_regress(const void* in, const void* wgt_packed, const float* samples, int B, int D, int H, int W, float* logits,
                          float* depth, float* conf, cudaStream_t stream);
/* prob head: plain Conv3d(8,1,3,p=1,bias=False) (models/module.py:303) -> fp32 logits [B,D,H,W]. */
int cds_prob_conv(const void* in, const float* wgt, int B, int Cin, int D, int H, int W, int dtype, float* logits,
                  cudaStream_t stream);

/* ---- A5: softmax over D + soft-argmin depth + 4-plane confidence ------------------------------ */
/* logits [B,D,h,w] fp32 (input_is_prob=1: already probabilities, as depth_regression/conf_regression
 * receive them, models/module.py:373-391).  depth [B,D] or [B,D,h,w].  Any of depth_out [B,h,w],
 * conf_out [B,h,w], prob_out [B,D,h,w] may be NULL.  (models/model.py:90-92) */
int cds_softmax_regress(const float* logits, const float* depth, int depth_per_pixel, int input_is_prob, int B, int D,
                        int h, int w, float* depth_out, float* conf_out, float* prob_out, cudaStream_t stream);

/* ---- A6/A7: DynamicConv and the feature extractor's 2-D plumbing ------------------------------ */
/* DynamicConv.forward (models/dynamic_conv.py:97-122) for n images in one launch.
 *   x: in_mode 0 -> [n,H,W,Cin] channels-last `dtype`; in_mode 1 -> planar fp32 images [*,3,H,W], item i
 *      reads image img_index[i] (NULL: i)
 *   in_stats/in_act: InstanceNorm statistics ([n,Cin,2] fp64 sum, sum of squares) + activation applied
 *      to x while loading (the producer's norm, models/module.py:66-69); NULL = x is used as is
 *   epipole [n,2] full-resolution (x,y), multiplied by epi_scale (1, 1/2, 1/4: models/module.py:239,242)
 *   w_att per branch [k*k][Cin][4] (a,b,c,0); w_conv per branch [k*k][Cin][Cout]; bias [K][Cout] or NULL
 *   gate: W1 [4][K] and b1 [4] with BatchNorm2d folded, W2 [K][4] (models/dynamic_conv.py:89-92)
 *   kernel_sizes: HOST array of num_kernels odd sizes <= 11
 *   out_raw [n,H,W,Cout] (pre-norm), out_stats [n,Cout,2] fp64 ACCUMULATED (zero them first)
 *   norm_curv / nc_abs [n,H,W] optional; nc_sq accumulates the stage curvature term
 *      (nc_a^2+nc_b^2+nc_c^2)/3 (models/module.py:250,257,264): nc_mode 0 = write, 1 = add, 2 = add and /3 */
int cds_dynamic_conv(const void* x, int in_mode, const int* img_index, const double* in_stats, int in_act,
                     const float* epipole, float epi_scale, const float* w_att, const float* w_conv, const float* bias,
                     const float* gate, int n, int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes,
                     float temperature, int dtype, void* out_raw, double* out_stats, float* norm_curv, float* nc_sq,
                     int nc_mode, float* nc_abs, cudaStream_t stream);
/* Tensor-core (tcgen05 + TMEM + TMA) DynamicConv for the feature extractor's layer shapes (8->8 with kernels
 * (3,7,11) [conv00, image padded to 8 channels by cds_image_to_nhwc8], (3,5,7), (1,3); 16->16 with (3,5), (1,3);
 * 32->32 with (1,3)), fp16 storage, W >= 8.  Same semantics and outputs as cds_dynamic_conv.
 * x: [n_images,H,W,Cin] fp16; item i reads image img_index[i] (NULL: i); wgt_packed: fp16 operand image from the host
 * (cds_dynamic_conv_tc_weight_halfs() halfs, layout in csrc/dynconv_tc.cu).  kernel_sizes is a HOST array.
 * Split-precision activations (value + fp16 rounding residual as two fp16 planes, ~22 bits): split_in = 1 means x is
 * [2,n_images,H,W,Cin] (residual plane second; 16->16 (3,5) layers only); out_lo, if non-NULL, receives the residual
 * plane of out_raw.  cds_conv2d_3x3s2 has the same out_lo. */
int cds_image_to_nhwc8(const float* img, int n, int H, int W, void* out, cudaStream_t stream);
/* 8-bit images as the data layer reads them: out[i] = (float)in[i] / 255.f, the IEEE quotient of datasets/general_eval.py:91
 * (np.float32(img) / 255.) evaluated on the device, so a caller may upload bytes instead of floats.  count elements. */
int cds_image_u8_to_f32(const unsigned char* in, long long count, float* out, cudaStream_t stream);
int cds_dynamic_conv_tc_supported(int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes);
int cds_dynamic_conv_tc_weight_halfs(int Cin, int Cout, int num_kernels, const int* kernel_sizes);
int cds_dynamic_conv_tc(const void* x, int n_images, const int* img_index, const double* in_stats, int in_act,
                        const float* epipole, float epi_scale, const void* wgt_packed, const float* bias, const float* gate,
                        int n, int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes, float temperature,
                        int split_in, void* out_raw, void* out_lo, double* out_stats, float* norm_curv, float* nc_sq,
                        int nc_mode, float* nc_abs, cudaStream_t stream);
/* FeatureNet.downsample1/2 (models/module.py:214,218): 3x3 stride 2 pad 1, wgt [9][Cin][Cout]. */
/* The image layer (conv00) over the cascade's pair batch: n = 2*V*B items ordered (side, v, b), item (0,v,b) = the reference
 * image of batch item b seen with pair v's epipole (models/model.py:154-161 recomputes it per pair), item (1,v,b) = source
 * image v; img_index maps the V side-0 items of a batch item to one image.  Results equal cds_dynamic_conv_tc on the same
 * batch; the reference image's branch convolutions (which do not depend on the epipole) run once per batch item. */
int cds_dynamic_conv_tc_pairs(const void* x, int n_images, const int* img_index, const float* epipole, float epi_scale,
                              const void* wgt_packed, const float* bias, const float* gate, int V, int B, int Cin, int Cout, int H, int W,
                              int num_kernels, const int* kernel_sizes, float temperature, void* out_raw, void* out_lo, double* out_stats,
                              float* norm_curv, float* nc_sq, int nc_mode, float* nc_abs, cudaStream_t stream);
/* Second tensor-core formulation for the TRUNK layers (conv00, conv01, conv10/11, conv20/21, out1): kernel rows folded into the
 * GEMM's N dimension, persistent row-streaming pipeline (csrc/dynconv_kh.cu).  Same contract as cds_dynamic_conv_tc, plus the
 * pair batch of cds_dynamic_conv_tc_pairs when pair_v > 0 (then n = 2*pair_v*pair_b).  split_in: x holds the residual plane
 * after the n_images value images.  wgt_packed: cds_dynamic_conv_kh_weight_halfs() halfs (host: weights.py
 * pack_dynamic_conv_kh): weights AND their fp16 rounding residuals, multiplied as separate accumulating products. */
int cds_dynamic_conv_kh_supported(int Cin, int Cout, int H, int W, int num_kernels, const int* kernel_sizes);
int cds_dynamic_conv_kh_weight_halfs(int Cin, int Cout, int num_kernels, const int* kernel_sizes);
/* The image layer on 8-BIT images (the bytes the reference's loader reads before np.float32(img) / 255.,
 * datasets/general_eval.py:88-91): cds_image_u8_to_px2 builds pixel-pair operand slots -- (RGB of the pixel, RGB of its right
 * neighbour, 0, 0) as byte / 256 (exact in fp16), rows padded by cds_dynamic_conv_kh_u8_pad() pixels on either side -- so that one
 * K = 16 MMA covers four horizontal taps; the factor 256 / 255 is folded into the weights.  Same outputs as the fp32-image path to fp32 rounding. */
int cds_dynamic_conv_kh_u8_weight_halfs(int Cout, int num_kernels, const int* kernel_sizes);
int cds_dynamic_conv_kh_u8_pad(void);
int cds_image_u8_to_px2(const unsigned char* img, int n_images, int H, int W, void* out, cudaStream_t stream);
int cds_dynamic_conv_kh_u8(const void* px2, int n_images, const int* img_index, const float* epipole, float epi_scale, const void* wgt_packed,
                           const float* gate, int n, int Cout, int H, int W, int num_kernels, const int* kernel_sizes, float temperature,
                           void* out_raw, void* out_lo, double* out_stats, float* norm_curv, float* nc_sq, int nc_mode, float* nc_abs,
                           int pair_v, int pair_b, cudaStream_t stream);
/* layout of the packed images: columns of one kernel-row group (Cout + 3 curvature columns, rounded up), columns of one image
 * of a k x k branch (k groups + zero padding) */
int cds_dynamic_conv_kh_group_cols(int Cout);
int cds_dynamic_conv_kh_image_cols(int Cout, int k);
int cds_dynamic_conv_kh(const void* x, int n_images, const int* img_index, const double* in_stats, int in_act, const float* epipole,
                        float epi_scale, const void* wgt_packed, const float* bias, const float* gate, int n, int Cin, int Cout, int H,
                        int W, int num_kernels, const int* kernel_sizes, float temperature, int split_in, void* out_raw, void* out_lo,
                        double* out_stats, float* norm_curv, float* nc_sq, int nc_mode, float* nc_abs, int pair_v, int pair_b,
                        cudaStream_t stream);
int cds_conv2d_3x3s2(const void* in, const double* in_stats, int in_act, const float* wgt, int n, int Cin, int Cout, int H,
                     int W, int dtype, void* out, void* out_lo, double* out_stats, cudaStream_t stream);
/* FeatureNet.inner1/2 (models/module.py:253-254,260-261): 1x1 conv over cat(nearest-up2(a), b);
 * a [n,H/2,W/2,Ca], b [n,H,W,Cb], wgt [Ca+Cb][Cout]. */
int cds_conv2d_1x1_cat(const void* a, const double* a_stats, int a_act, const void* b, const double* b_stats, int b_act,
                       const float* wgt, int n, int Ca, int Cb, int Cout, int H, int W, int dtype, void* out,
                       double* out_stats, cudaStream_t stream);
/* Tensor-core (tcgen05) forms of the two plain 2-D convs above, fp16 storage (csrc/conv2d_gtc.cu): same semantics, the
 * operand is gathered + normalised per output pixel; wgt_packed = fp16 image [mma][k-chunk 2][2*Cout/8][8 n][8 k] with the
 * weights' fp16 rounding residual in columns [Cout, 2*Cout) (host: weights.py pack_conv2d_gtc). */
int cds_conv2d_3x3s2_tc_supported(int Cin, int Cout);
int cds_conv2d_3x3s2_tc_weight_halfs(int Cin, int Cout);
/* in_lo (optional, shape of in): the fp16 rounding residual plane of in; out_lo (optional, shape of out) receives that of out. */
int cds_conv2d_3x3s2_tc(const void* in, const void* in_lo, const double* in_stats, int in_act, const void* wgt_packed, int n, int Cin,
                        int Cout, int H, int W, void* out, void* out_lo, double* out_stats, cudaStream_t stream);
/* The same stride-2 layer as a persistent row-streaming kernel (csrc/conv2d_s2rows.cu; models/module.py:214,218): every
 * input row is loaded once by TMA (even / odd pixel phases as dense operand slabs) and normalised once.  Needs W even and the
 * input statistics; wgt_packed = [kernel row 3][image][k-chunk 2][2*Cout/8][8 n][8 k] fp16 (host: weights.py
 * pack_conv2d_s2rows).  Same tensors as cds_conv2d_3x3s2_tc. */
int cds_conv2d_3x3s2_rows_supported(int Cin, int Cout, int H, int W);
int cds_conv2d_3x3s2_rows_weight_halfs(int Cin, int Cout);
int cds_conv2d_3x3s2_rows(const void* in, const void* in_lo, const double* in_stats, int in_act, const void* wgt_packed, int n, int Cin,
                          int Cout, int H, int W, void* out, void* out_lo, double* out_stats, cudaStream_t stream);
int cds_conv2d_1x1_cat_tc_supported(int Ca, int Cb, int Cout);
int cds_conv2d_1x1_cat_tc_weight_halfs(int Ca, int Cb, int Cout);
int cds_conv2d_1x1_cat_tc(const void* a, const double* a_stats, int a_act, const void* b, const double* b_stats, int b_act,
                          const void* wgt_packed, int n, int Ca, int Cb, int Cout, int H, int W, void* out, double* out_stats,
                          cudaStream_t stream);
/* InstanceNorm2d(affine=False, eps 1e-5, biased variance) + activation, materialised. */
int cds_instnorm_act(const void* raw, const double* stats, int act, int n, int C, int H, int W, int dtype, void* out,
                     cudaStream_t stream);
/* Same from split-precision fp16 storage: raw + raw_lo (its fp16 rounding residual plane, may be NULL) -> fp32 out [n,H,W,C]
 * (the stage-1 feature of the cascade, whose rounding the later stages amplify: DESIGN.md section 3); out_f16 (optional,
 * same shape) receives the fp16 rounding of out. */
int cds_instnorm_act_split_f32(const void* raw, const void* raw_lo, const double* stats, int act, int n, int C, int H, int W,
                               float* out, void* out_f16, cudaStream_t stream);
/* fp32 NCHW <-> channels-last storage type, for the op-level drop-ins' public signatures. */
int cds_nchw_to_nhwc(const float* in, int n, int C, int H, int W, int dtype, void* out, cudaStream_t stream);
int cds_nhwc_to_nchw(const void* in, int n, int C, int H, int W, int dtype, float* out, cudaStream_t stream);

/* ---- next row: geometric-consistency filter, the step after the depth-inference path ----------- */
/* Reference: fusion.py:49-117 (get_reproj, project_img, vis_filter, ave_fusion, prob_filter), driven by test.py:326-352.
 * cams are [.,2,4,4] = (extrinsic, intrinsic in [1,:3,:3]).  cds_fusion_setup turns (ref_cam [n,2,4,4], srcs_cam [n,v,2,4,4])
 * into mats [n,v,100] (matrices + fp64-computed inverses, cds_fusion_mats_floats(n, v) floats).  cds_geometric_filter does
 * get_reproj + vis_filter + ave_fusion + back-projection in one pass; every output may be NULL: reproj_xyd [n,v,3,h,w],
 * in_range [n,v,h,w], masks [n,v,h,w] (fp32 0/1), vis_mask [n,h,w] uint8, ave [n,h,w], points [n,3,h,w] (world frame). */
int cds_fusion_mats_floats(int n, int v);
int cds_fusion_setup(const float* ref_cam, const float* srcs_cam, int n, int v, float* mats, cudaStream_t stream);
int cds_geometric_filter(const float* ref_depth, const float* srcs_depth, const float* mats, int n, int v, int h, int w,
                         float img_dist_thresh, float depth_thresh, float vthresh, float* reproj_xyd, float* in_range, float* masks,
                         unsigned char* vis_mask, float* ave, float* points, cudaStream_t stream);
/* vis_filter / ave_fusion on materialised reprojections (masks_in NULL: masks computed, else used as given). */
int cds_vis_filter(const float* ref_depth, const float* reproj_xyd, const float* in_range, const float* masks_in, int n, int v, int h,
                   int w, float img_dist_thresh, float depth_thresh, float vthresh, float* masks_out, unsigned char* vis_mask,
                   float* ave, cudaStream_t stream);
/* prob_filter: mask = AND_c prob[:,c] > thresholds[c] (thresholds on the HOST, C <= 4); optional depth_inout *= mask. */
int cds_prob_filter(const float* prob, const float* thresholds, int n, int C, int h, int w, unsigned char* mask, float* depth_inout,
                    cudaStream_t stream);

/* ---- next row: the Refinement network (models/module.py:318-370; models/model.py:209-216) ------------ */
/* fp32 planar NCHW, BatchNorm folded by the host (weights.py pack_refinement).  Refinement.forward is the sequence
 * cds_refine_prescale -> cds_conv2d_3x3_f32 x3 (conv0 on the image; conv1, conv2 on the depth) -> cds_deconv2d_k3s2_f32 ->
 * cds_conv2d_3x3_f32 (conv3 over cat(deconv, conv0)) -> cds_refine_final (res conv + bilinear up2 of the normalised depth +
 * rescale; post[b] = optional per-item multiplier, the depth interval of models/model.py:216). */
int cds_refine_prescale(const float* depth0, const float* lo, const float* hi, int B, int h, int w, float* depth_n, cudaStream_t stream);
int cds_conv2d_3x3_f32(const float* a, const float* b, const float* wgt, const float* bias, int B, int Ca, int Cb, int Cout, int H, int W,
                       int relu, float* out, cudaStream_t stream);
int cds_deconv2d_k3s2_f32(const float* in, const float* wgt, const float* bias, int B, int C, int h, int w, float* out, cudaStream_t stream);
int cds_refine_final(const float* x, const float* res_wgt, const float* depth_n, const float* lo, const float* hi, const float* post,
                     int B, int h, int w, float* out, cudaStream_t stream);

/* ---- next row: CostRegNet in TRAINING mode and its backward (csrc/train3d.cu) ----------------------------
 * fp32 planar NCDHW tensors, weights tap-major [Cin][27][Cout] (permuted on the device by the host side).  Reference:
 * models/module.py:80-122 (Conv3d k3 p1 s1|2 -> BatchNorm3d with batch statistics -> ReLU), :125-166 (ConvTranspose3d k3 s2 p1
 * op1 -> BN -> ReLU), :303-315 (wiring; skips are added after the ReLU); gradients as torch.autograd computes them. */
/* direct convolution, out [B,Cout,ceil(D/s),ceil(H/s),ceil(W/s)]; also the input gradient of a stride-1 block (flipped /
 * transposed weights) and of a transposed block (its weight as a stride-2 conv weight) */
int cds_train_conv3d(const float* x, const float* wgt, int B, int Cin, int Cout, int D, int H, int W, int stride, float* out,
                     cudaStream_t stream);
/* direct transposed convolution k3 s2 p1 op1, out [B,Cout,2D,2H,2W]; also the input gradient of a stride-2 block */
int cds_train_deconv3d(const float* x, const float* wgt, int B, int Cin, int Cout, int D, int H, int W, float* out, cudaStream_t stream);
/* dw [Cin][27][Cout] = sum_{b,o} g[b,co,o] x[b,ci,o*s-1+tap] (x [B,Cin,D,H,W], g [B,Cout,ceil(D/s),..]); the transposed
 * block's weight gradient is the same call with input and gradient swapped */
int cds_train_conv3d_wgrad(const float* x, const float* g, int B, int Cin, int Cout, int D, int H, int W, int stride, float* dw,
                           cudaStream_t stream);
/* BatchNorm3d, batch statistics: sums [C][2] fp64 = (sum x, sum x^2) over (b, V voxels) */
int cds_train_bn_stats(const float* x, int B, int C, long long V, double* sums, cudaStream_t stream);
/* y = relu?(gamma (x - mean) rstd + beta) (+ skip after the ReLU; skip may be NULL) */
int cds_train_bn_apply(const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta, const float* skip,
                       int relu, int B, int C, long long V, float* y, cudaStream_t stream);
/* backward of bn_apply w.r.t. x: dx; sums [C][2] fp64 receives (dbeta, dgamma) = (sum dz, sum dz xhat), dz = dy [z > 0] */
int cds_train_bn_backward(const float* dy, const float* x, const float* mean, const float* rstd, const float* gamma, const float* beta,
                          int relu, int B, int C, long long V, double* sums, float* dx, cudaStream_t stream);

/* DynamicConv in TRAINING mode (csrc/train2d.cu): the k x k convolutions of its branches, fp32 planar NCHW, stride 1, pad (k-1)/2,
 * odd k <= 11; weights tap-major [Cin][k*k][Cout].  Reference: models/dynamic_conv.py:84-88,112-116.  The same call with flipped,
 * transposed weights is the input gradient. */
int cds_train_conv2d(const float* x, const float* wgt, int B, int Cin, int Cout, int H, int W, int k, float* out, cudaStream_t stream);
/* dw [Cin][k*k][Cout] = sum_{b,p} g[b,co,p] x[b,ci,p + tap - (k-1)/2] */
int cds_train_conv2d_wgrad(const float* x, const float* g, int B, int Cin, int Cout, int H, int W, int k, float* dw, cudaStream_t stream);

/* ---- next row (first slice): backward passes of the op-level drop-ins and the stage loss ---------------- */
/* Adjoint of cds_homo_warp in src_fea (the sampling grid carries no gradient, models/utils/warping.py:79): grad_out
 * [B,C,D,h,w] fp32 is scattered with the forward's bilinear weights into grad_src [B,C,h,w] fp32 (float reductions: the
 * order of the adds, hence the last bits, varies from run to run).  workspace == NULL: scalar reductions straight into
 * grad_src, which the CALLER ZEROES first (any C).  workspace != NULL (C % 4 == 0): B*h*w*C floats ZEROED by the caller
 * receive 16-byte vector reductions in channels-last order and a second kernel writes grad_src (need not be zeroed). */
int cds_homo_warp_backward(const float* grad_out, const float* coef, const float* depth, int depth_per_pixel, int B, int C,
                           int D, int h, int w, float* grad_src, float* workspace, cudaStream_t stream);
/* depth_regression backward (models/module.py:373-379): grad_p[b,d] = grad_depth[b] * depth[b,d] and
 * grad_dv[b,d] = grad_depth[b] * prob[b,d], both [B,D,h,w]; either output may be NULL (then its input may be too). */
int cds_depth_regress_backward(const float* grad_depth, const float* prob, const float* depth, int depth_per_pixel, int B, int D,
                               int h, int w, float* grad_p, float* grad_dv, cudaStream_t stream);
/* One stage of final_loss (models/losses.py:14-23): ADDS to sums[0..2] (fp64, zeroed by the caller) the smooth-L1 sum of
 * est/interval - gt/interval over mask > 0.5, the mask count, and the masked sum of norm_curv (NULL: skipped).
 * est, gt, mask, norm_curv [B,h,w] fp32; interval [B]. */
int cds_stage_loss_forward(const float* est, const float* gt, const float* mask, const float* interval, const float* norm_curv,
                           int B, int h, int w, double* sums, cudaStream_t stream);
/* Its backward: grad_est = *g_depth * smooth_l1'(.) / (interval * count), grad_curv = *g_curv / count on the mask, 0 off it
 * (g_depth, g_curv: device scalars, the upstream gradients of the two means; sums from the forward call). */
int cds_stage_loss_backward(const float* est, const float* gt, const float* mask, const float* interval, const double* sums,
                            const float* g_depth, const float* g_curv, int B, int h, int w, float* grad_est, float* grad_curv,
                            cudaStream_t stream);
/* The feat_distance term of final_loss (models/losses.py:25-35): binary cross entropy with logits over mask > 0.5 repeated
 * across the D planes, positives weighted by neg / pos.  logits, target [B,D,h,w]; mask [B,h,w].  ADDS to sums[0..2] (fp64,
 * zeroed by the caller): sum of target over the selection, size of the selection, sum of the weighted loss terms (two
 * kernels: the weight needs the counts).  Backward: grad [B,D,h,w] = *g * d(mean)/d(logits), 0 off the selection. */
int cds_feat_loss_forward(const float* logits, const float* target, const float* mask, int B, int D, int h, int w, double* sums,
                          cudaStream_t stream);
int cds_feat_loss_backward(const float* logits, const float* target, const float* mask, const double* sums, const float* g, int B,
                           int D, int h, int w, float* grad, cudaStream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CDS_B200_H */
