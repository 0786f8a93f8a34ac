/*
 * m4d.h - C ABI of libm4d.so, the B200-native (sm_100a) implementation of M4Depth's per-frame
 * parallax-inference hot path.
 *
 * This is the drop-in boundary: every entry point below is what a binding for the reference's
 * operator / function would call (TF custom-op shim, ctypes, cgo ...).  Plain pointers and sizes only.
 * Reference citations are relative to the M4Depth repository root.
 *
 * Conventions
 *   - All pointers are DEVICE pointers unless the name ends in _host.  All tensors are dense NHWC fp32
 *     exactly like the reference's tensors; "pix_stride" arguments (in floats) let a kernel write its
 *     channels straight into a wider buffer (the refiner input) instead of a compact tensor.
 *   - The caller owns every buffer.  The library never allocates device memory, keeps no hidden state,
 *     never synchronises and enqueues everything (including zero fills) on the caller's stream, so every
 *     call is CUDA-graph capturable and re-entrant.  (The reference memsets on the legacy default stream
 *     and exit(-1)s on launch errors: cuda_backproject/backproject_op_gpu.cu.cc:91-100.)
 *   - Return value: 0 on success, a negative M4D_E* code otherwise; m4d_last_error_string() returns a
 *     thread-local description of the last failure.  Nothing ever calls exit().
 *   - stream is a cudaStream_t passed as void* (NULL = legacy default stream).
 *   - rot is [b, rot_dim] with rot_dim 4 = quaternion (w,x,y,z), 3 = small angles (x,y,z)
 *     (utils/depth_operations.py:18-53); trans is [b,3]; cam_f / cam_c are [b,2] = (fx,fy) / (cx,cy) ALREADY
 *     divided by 2^level as DepthEstimatorPyramid.call does (m4depth_network.py:300-302).
 *   - Floating-point contract: geometry is evaluated with one rounded fp32 operation per reference TF op
 *     (no FMA contraction, IEEE division and square root) so the integer tap grids x0,x1,y0,y1 are
 *     bit-identical to the oracle's; see DESIGN.md "Numerics".
 */
#ifndef M4D_H_
#define M4D_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define M4D_OK 0
#define M4D_EINVAL (-1)   /* bad argument (null pointer, non-positive size, unsupported shape) */
#define M4D_ECUDA (-2)    /* CUDA runtime / launch error; see m4d_last_error_string() */
#define M4D_ENOTSUP (-3)  /* shape outside what the kernel family supports */

#define M4D_ABI_VERSION 1

int m4d_abi_version(void);
const char* m4d_last_error_string(void);
/* Number of kernels launched by this library since load (all threads); used by bench.py for "gpu_launches". */
uint64_t m4d_launch_count(void);

/* ---- L0: BackProject op ------------------------------------------------------------------------
 * Replaces BackProjectForwardLauncher (cuda_backproject/backproject_op_gpu.h:17-18, kernel
 * backproject_op_gpu.cu.cc:19-79) = TF op "BackProject" (backproject_op.cc:32-35).
 * dim = {B,H,W,S,F,C}; input [B,H,W,F,C]; coords [B,H,W,S,F,2] (x,y); out [B,H,W,S,F,C].
 * out is fully written (zero where the coordinate is outside [0,W-1]x[0,H-1] or NaN); no separate memset.
 * idx_dbg (nullable) int32 [B,H,W,S,F,4] receives the tap grid x0,x1,y0,y1 (-1 where outside). */
int m4d_backproject_fwd(const float* input, const float* coords, const int32_t dim[6], float* out,
                        int32_t* idx_dbg, void* stream);
/* Replaces BackProjectBackwardLauncher (backproject_op_gpu.h:20-22, kernel backproject_op_gpu.cu.cc:108-196) = TF op
 * "BackProjectGrad" (backproject_op.cc:37-42), the gradient registered for BackProject (utils/dense_image_warp.py:46-52).
 * grad [B,H,W,S,F,C] -> input_grad [B,H,W,F,C] (zeroed here, on the caller's stream, then scatter-added with
 * floating-point atomics as in the reference: summation order not fixed) and coords_grad [B,H,W,S,F,2] (d/dx, d/dy; fully
 * written, zero where the coordinate is outside the image or NaN; deterministic). */
int m4d_backproject_bwd(const float* grad, const float* input, const float* coords, const int32_t dim[6],
                        float* input_grad, float* coords_grad, void* stream);

/* ---- L1: dense_image_warp (utils/dense_image_warp.py:195-268, BackProject branch :246-253) -------
 * image [b,h,w,c], flow [b,h,w,2] (row, col); query = grid + flow, clipped to the image. */
int m4d_dense_image_warp(const float* image, const float* flow, int b, int h, int w, int c, float* out,
                         void* stream);

/* ---- L2: geometry (utils/depth_operations.py) ---------------------------------------------------*/
/* get_rot_mat :18-53  -> out [b,9] row-major */
int m4d_get_rot_mat(const float* rot, int b, int rot_dim, float* out, void* stream);
/* prev_d2para :196-215 ; parallax2depth :140-166 ; depth2parallax :168-194 ; maps are [b,h,w,1] */
int m4d_prev_d2para(const float* prev_d, const float* rot, int rot_dim, const float* trans, const float* cam_f,
                    const float* cam_c, int b, int h, int w, float* out, void* stream);
int m4d_parallax2depth(const float* para, const float* rot, int rot_dim, const float* trans, const float* cam_f,
                       const float* cam_c, int b, int h, int w, float* out, void* stream);
int m4d_depth2parallax(const float* depth, const float* rot, int rot_dim, const float* trans, const float* cam_f,
                       const float* cam_c, int b, int h, int w, float* out, void* stream);

/* ---- L2: fused backproject + parallax-sweeping cost volume --------------------------------------
 * Replaces get_parallax_sweeping_cv (utils/depth_operations.py:223-281) together with the
 * tile_in_batch copies (:217-221, :267-268), dense_image_warp and BackProject it calls, and the fp16
 * correlate (:276-278).  K = 2*search_range+1.
 *   c1, c2        [b,h,w,c]   current / previous (already group-normalised) feature maps
 *   para_prev_t   [b,h,w]     (disp_prev_t) sampled alongside c2
 *   para_prev_l   [b,h,w]     (disp) the parallax the sweep is centred on
 *   cv            [b,h,w,cuts*K] cut-major (channel = cut*K + k), row stride cv_pix_stride floats
 *   prev_disp     nullable, [b,h,w,K] row stride pd_pix_stride floats
 *   centre_log    nullable, receives log(prev_disp[...,search_range] * centre_log_scale) at
 *                 centre_log[pixel*centre_log_pix_stride] (m4depth_network.py:238); with prev_disp NULL
 *                 only this consumed channel is produced
 *   idx_dbg       nullable int32 [b,h,w,K,4] = x0,x1,y0,y1 of the BackProject convention (-1 outside)
 * fp16 contract: products of fp16-rounded operands rounded to fp16, summed in fp32, /n, rounded once to fp16. */
int m4d_pscv_fused_fwd(const float* c1, const float* c2, const float* para_prev_t, const float* para_prev_l,
                       const float* rot, int rot_dim, const float* trans, const float* cam_f, const float* cam_c,
                       int b, int h, int w, int c, int cuts, int search_range,
                       float* cv, int cv_pix_stride, float* prev_disp, int pd_pix_stride,
                       float* centre_log, int centre_log_pix_stride, float centre_log_scale,
                       int32_t* idx_dbg, void* stream);

/* Bilinear-sampling convention of the warp inside the PSCV.  The reference has two code paths for the same op:
 *   GATHER  utils/dense_image_warp.py:127-190,255-259 - what runs on the TF CPU path and whenever backproject.so is
 *           not found (the default: make.sh writes the .so where dense_image_warp.py:38 does not look);
 *   BP      dense_image_warp.py:246-253 + backproject_op_gpu.cu.cc:44-76 with every product / sum rounded separately;
 *   BP_FMA  the same kernel expression (:74) as nvcc compiles it by default (mul, fma, fma, fma) - the arithmetic of
 *           the reference's own GPU binary.
 * m4d_pscv_fused_fwd uses GATHER; m4d_pscv_fused_fwd_ex takes the convention explicitly. */
#define M4D_INTERP_GATHER 0
#define M4D_INTERP_BP 1
#define M4D_INTERP_BP_FMA 2
/* OR-ed into interp: use the shape-generic kernel even where the specialised one (search_range 4, c in
 * {16,32,64,96,128,192}) applies.  Same results bit for bit; tests cross-check the two implementations with it. */
#define M4D_INTERP_FLAG_GENERIC 0x100
/* OR-ed into interp: use the CTA-tile kernel (pscv9_kernel) where the warp-autonomous one (pscv9w_kernel; the network's
 * (c, cuts) pairs) would run.  Same results bit for bit.  Bits 12-15, when non-zero, override the resident CTAs per SM
 * of the persistent grid (tuning experiments only). */
#define M4D_INTERP_FLAG_TILE 0x200
/* OR-ed into interp: use the warp-autonomous LDG-gather kernel (pscv9w_kernel) where the shared-memory staged one
 * (pscv9s_kernel: gather convention, c = 32 / cuts = 2) would run.  Same results bit for bit. */
#define M4D_INTERP_FLAG_WARP 0x400
int m4d_pscv_fused_fwd_ex(const float* c1, const float* c2, const float* para_prev_t, const float* para_prev_l,
                          const float* rot, int rot_dim, const float* trans, const float* cam_f, const float* cam_c,
                          int b, int h, int w, int c, int cuts, int search_range,
                          float* cv, int cv_pix_stride, float* prev_disp, int pd_pix_stride,
                          float* centre_log, int centre_log_pix_stride, float centre_log_scale,
                          int32_t* idx_dbg, int interp, void* stream);

/* Debug / test hook: the branch-free IEEE division of the shared-memory staged PSCV kernel (csrc/pscv_smem.cu, phase 0) next
 * to __fdiv_rn on n operand pairs; unsafe_flag[i] = 1 where the kernel would fall back to __fdiv_rn.  Tests require
 * out_fast == out_ieee bit for bit wherever the flag is 0. */
int m4d_debug_div_check(const float* a, const float* b, int n, float* out_fast, float* out_ieee, int32_t* unsafe_flag, void* stream);

/* Backward of m4d_pscv_fused_fwd in the gather convention (what TensorFlow's autodiff makes of
 * utils/depth_operations.py:223-281 + utils/dense_image_warp.py:127-190, needed by train_step, m4depth_network.py:371-399):
 * (d_cv [b,h,w,>=cuts*K], d_prev_disp [b,h,w,>=K] or NULL) -> d_c1 [b,h,w,c] and d_para_prev_l [b,h,w,1] (per-pixel sums,
 * deterministic), d_c2 [b,h,w,c] and d_para_prev_t [b,h,w,1] (zeroed here, then scatter-added onto the bilinear taps with
 * floating-point atomics).  The pose and camera get no gradient (they are data).  A parity implementation (one thread per
 * pixel), not a tuned kernel. */
int m4d_pscv_fused_bwd(const float* c1, const float* c2, const float* para_prev_t, const float* para_prev_l,
                       const float* rot, int rot_dim, const float* trans, const float* cam_f, const float* cam_c,
                       int b, int h, int w, int c, int cuts, int search_range,
                       const float* d_cv, int d_cv_pix_stride, const float* d_prev_disp, int d_pd_pix_stride,
                       float* d_c1, float* d_c2, float* d_para_prev_t, float* d_para_prev_l, void* stream);

/* ---- L2: spatial self-correlation cost volume ---------------------------------------------------
 * Replaces cost_volume (utils/depth_operations.py:283-313), dilation 1.
 * out[b,y,x,(dy*(2r+1)+dx)*cuts + k] = leaky_0.1(mean_{j in group k} c1[b,y,x,j]*c2pad[b,y+dy-r,x+dx-r,j]) */
int m4d_sncv_fwd(const float* c1, const float* c2, int b, int h, int w, int c, int cuts, int search_range,
                 float* out, int out_pix_stride, void* stream);
/* variant: AUTO = the column-strip kernel where it applies (group width 16 / 24 / 32, cuts <= 4), else the (pixel, dy)
 * kernel; PIXEL_DY forces the latter.  Both produce the same bits; tests cross-check them. */
#define M4D_SNCV_AUTO 0
#define M4D_SNCV_PIXEL_DY 1
#define M4D_SNCV_COLUMN_STRIP 2 /* column-strip kernel whenever the shape allows, however few tiles */
int m4d_sncv_fwd_ex(const float* c1, const float* c2, int b, int h, int w, int c, int cuts, int search_range,
                    float* out, int out_pix_stride, int variant, void* stream);

/* ---- L3: layer pieces (m4depth_network.py) ------------------------------------------------------*/
/* tf.linalg.normalize per feature group, no epsilon (:180,185-186).  in/out [b,h,w,c] (may alias). */
int m4d_group_l2norm(const float* in, int npix, int c, int cuts, float* out, void* stream);

/* DomainNormalization.call (:44-48).  x [b,h,w,c] (c <= 64); stats_ws: caller-provided workspace of
 * 2*b*c doubles (need not be initialised); leaky_alpha != 1 additionally applies leaky_relu (:84). */
int m4d_domain_norm(const float* x, int b, int h, int w, int c, const float* scale, const float* bias,
                    float leaky_alpha, double* stats_ws, float* out, void* stream);
/* FeaturePyramid level 0 in one call (m4depth_network.py:79-84): y = leaky(DN(conv3x3_same(x) + conv_bias)) for the 3-channel
 * image and the 16-channel first layer.  The conv output is never stored: both DN passes recompute it from the image (same FMA
 * chain), so the statistics describe exactly the values that are normalised.  x [b,h,w,>=3] with pixel stride x_pix_stride;
 * kernel HWIO [3,3,3,16]; stats_ws: 2*b*16 doubles (need not be initialised); out [b,h,w,16], 16-byte aligned. */
int m4d_rgb_conv_dn(const float* x, int x_pix_stride, const float* kernel_hwio, const float* conv_bias, int b, int h, int w,
                    const float* dn_scale, const float* dn_bias, float leaky_alpha, double* stats_ws, float* out, void* stream);
/* The same call with the conv kernel [3,3,3,16] and its bias given as HOST pointers: they travel in the kernel parameter block
 * and are read as constant operands of the FMAs (no shared-memory weight loads: ~2x faster than the call above, and faster
 * than storing the conv output).  Same results bit for bit.  Everything else as above (x, DN parameters, workspace: device). */
/* The two halves of the first encoder layer as separate calls: the conv (weights as HOST pointers, as below) evaluated once,
 * y = conv3x3_same(x) + bias stored [b,h,w,16] and its per-(image, channel) sum / sum of squares accumulated into stats_ws
 * (2*b*16 doubles, zeroed by the call); then the apply pass of DomainNormalization on those statistics (c = 16 or 32). */
int m4d_rgb_conv_stats_hostw(const float* x, int x_pix_stride, const float* kernel_hwio_host, const float* conv_bias_host, int b,
                             int h, int w, float* y, double* stats_ws, void* stream);
int m4d_domain_norm_apply(const float* x, int b, int h, int w, int c, const float* scale, const float* bias, float leaky_alpha,
                          const double* stats_ws, float* out, void* stream);
int m4d_rgb_conv_dn_hostw(const float* x, int x_pix_stride, const float* kernel_hwio_host, const float* conv_bias_host, int b, int h,
                          int w, const float* dn_scale, const float* dn_bias, float leaky_alpha, double* stats_ws, float* out,
                          void* stream);

/* Keras Conv2D(3x3, padding='same') + bias + optional leaky_relu (:63-72,104-114; TF SAME padding rule).
 * x [b,h,w,cin] with row stride x_pix_stride (>= cin); kernel HWIO [3,3,cin,cout]; y [b,oh,ow,cout] with row
 * stride y_pix_stride, oh = ceil(h/stride).  leaky_alpha = 1 means no activation.
 * algo: 0 = auto, 1 = FFMA2 direct (the only kernel that takes unpacked weights; 2 is reserved and returns M4D_ENOTSUP:
 * the tensor-core path is m4d_conv3x3_tc_fwd below). */
int m4d_conv3x3_nhwc(const float* x, int x_pix_stride, const float* kernel_hwio, const float* bias,
                     int b, int h, int w, int cin, int cout, int stride, float leaky_alpha,
                     float* y, int y_pix_stride, int algo, void* stream);

/* Tensor-core (tcgen05, 3xTF32) path of the same stride-1 convolution, for the refiner layers that hold ~95 % of a frame's
 * FLOPs (m4depth_network.py:104-114).  Weights are split into TF32 hi / lo planes and packed once per layer:
 *   m4d_conv3x3_tc_packed_floats  size (in floats) of the packed buffer; 0 if (cin, cout) is outside the path
 *                                 (1 <= cout <= 128, padded internally to a multiple of 16; any cin)
 *   m4d_conv3x3_tc_pack           kernel HWIO [3,3,cin,cout] -> packed (device buffer of that size, 16-byte aligned)
 *   m4d_conv3x3_tc_fwd            y = leaky(conv(x) + bias); the x pixel stride must be a multiple of 4 floats and x
 *                                 16-byte aligned (TMA), else M4D_ENOTSUP (callers fall back to m4d_conv3x3_nhwc).
 * Result: every product is evaluated as hi*hi + hi*lo + lo*hi with fp32 accumulation, i.e. to ~2^-22 relative - the same
 * accuracy class as the FFMA kernel, with a different summation order. */
int64_t m4d_conv3x3_tc_packed_floats(int cin, int cout);
int m4d_conv3x3_tc_pack(const float* kernel_hwio, int cin, int cout, float* packed, void* stream);
int m4d_conv3x3_tc_fwd(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w,
                       int cin, int cout, float leaky_alpha, float* y, int y_pix_stride, void* stream);
/* The same three calls with the stride explicit.  stride 2 = FeaturePyramid's down-sampling convs (m4depth_network.py:66-72)
 * as a 2x2-cell convolution on the tensor cores: needs even h and w (TF SAME then pads bottom / right only), cin % 16 == 0
 * and x_pix_stride == cin; odd sizes return M4D_ENOTSUP (callers fall back to m4d_conv3x3_nhwc).  cout up to 256 (beyond
 * 128 a pixel tile's output channels are split over 2 or 4 CTAs). */
int64_t m4d_conv3x3_tc_packed_floats_s(int cin, int cout, int stride);
int m4d_conv3x3_tc_pack_s(const float* kernel_hwio, int cin, int cout, int stride, float* packed, void* stream);
int m4d_conv3x3_tc_fwd_s(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w,
                         int cin, int cout, int stride, float leaky_alpha, float* y, int y_pix_stride, void* stream);
/* slices: 0 = let the library choose how many output-channel slices (1, 2 or 4 CTAs per pixel tile) a layer runs as -
 * layers with fewer tiles than SMs are sliced to shorten the serial MMA chain; 1 / 2 / 4 force it (tests, tuning). */
int m4d_conv3x3_tc_fwd_ex(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w,
                          int cin, int cout, int stride, float leaky_alpha, float* y, int y_pix_stride, int slices, void* stream);
/* Precision mode of the tensor-core path, chosen at pack time (the packed layouts differ) and passed again to the forward:
 *   3XTF32  operands split into TF32 hi + lo, kind::tf32 MMAs (the calls above)
 *   3XFP16  operands scaled by a power of two (per layer for the weights, per pixel tile and 32-channel block for the
 *           activations, from the data) and split into fp16 h1 + 2^-11 h2, kind::f16 MMAs at twice the TF32 rate; the same
 *           three products hi*hi + hi*lo + lo*hi with fp32 accumulation, i.e. the same ~2^-22 relative error class. */
/* Debug: in libraries built with -DM4D_TC_PROFILE (M4D_NVCC_EXTRA=-DM4D_TC_PROFILE python m4depth_b200/_build.py --force) every
 * following tensor-core conv launch writes, per CTA, 16 int64 clock64 sums of its warp roles' waits and work into device_buf
 * ([sm_count][16]; tools/conv_phases.py prints them); NULL switches it off.  Returns 0 in normal builds (no timers compiled in). */
int m4d_debug_conv_profile(long long* device_buf);
#define M4D_CONV_PREC_3XTF32 0
#define M4D_CONV_PREC_3XFP16 1
int64_t m4d_conv3x3_tc_packed_floats_p(int cin, int cout, int stride, int prec);
int m4d_conv3x3_tc_pack_p(const float* kernel_hwio, int cin, int cout, int stride, int prec, float* packed, void* stream);
/* Flags that may be OR-ed into `slices` (bits 0-3 = the slice count above):
 *   M4D_CONV_PDL                 launch with programmatic stream serialization: the kernel may be scheduled while the kernel
 *                                before it in the stream is still draining (its set-up then overlaps that tail) and waits
 *                                (griddepcontrol.wait) before it reads x or the weights - same results, same stream order.
 *   M4D_CONV_PDL_WEIGHTS_STABLE  with M4D_CONV_PDL: `packed` and `bias` were complete before the preceding kernel was
 *                                enqueued (true from a layer's second call on), so the weight loads need not wait. */
#define M4D_CONV_PDL (1 << 8)
#define M4D_CONV_PDL_WEIGHTS_STABLE (1 << 9)
int m4d_conv3x3_tc_fwd_p(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w,
                         int cin, int cout, int stride, int prec, float leaky_alpha, float* y, int y_pix_stride, int slices,
                         void* stream);

/* tf.compat.v1.image.resize_bilinear, align_corners=False, no half-pixel (:202-204); post_scale multiplies the
 * result (parallax is doubled after resizing).  in [b,ih,iw,c] -> out [b,oh,ow,c] with row stride. */
int m4d_resize_bilinear_legacy(const float* in, int b, int ih, int iw, int c, int oh, int ow, float post_scale,
                               float* out, int out_pix_stride, void* stream);
/* tf.image.resize(method=NEAREST) (:368-369) */
int m4d_resize_nearest(const float* in, int b, int ih, int iw, int c, int oh, int ow, float* out, void* stream);

/* Fused level prologue (:196-204, 218, 224, 227): from the previous (coarser) level's estimate
 * {other [b,ih,iw,4], parallax, depth [b,ih,iw]} (all NULL for the deepest level -> 1 / 1000 / 0) produce
 *   para_prev_l, depth_prev_l [b,h,w]; other_prev_l -> x_in[..., ch_other..ch_other+3];
 *   log(para_prev_l * log_scale) -> x_in[..., ch_logpara]; and, when state_depth != NULL,
 *   para_prev_t = prev_d2para(state_depth) [b,h,w].
 * x_in may be NULL (new-trajectory frame: only the pass-through values are needed); ch_other < 0 skips the
 * "other" channels (level_memory ablation); other_out (nullable) receives other_prev_l compactly [b,h,w,4]. */
int m4d_level_prologue(const float* prev_other, const float* prev_para, const float* prev_depth, int ih, int iw,
                       const float* state_depth, const float* rot, int rot_dim, const float* trans,
                       const float* cam_f, const float* cam_c, int b, int h, int w,
                       float* para_prev_l, float* depth_prev_l, float* other_out, float* para_prev_t,
                       float* x_in, int x_pix_stride, int ch_logpara, int ch_other, float log_scale, void* stream);

/* Fused level epilogue (:247-260): refiner output r [b,h,w,5] (row stride r_pix_stride) ->
 * parallax = exp(clip(r0,-7,7)) * inv_scale, depth = parallax2depth(parallax), other = r[1:5];
 * depth_state (nullable) receives a second copy of depth: what the level stores as depth_prev_t (:260). */
int m4d_level_epilogue(const float* r, int r_pix_stride, const float* rot, int rot_dim, const float* trans,
                       const float* cam_f, const float* cam_c, int b, int h, int w, float inv_scale,
                       float* parallax, float* depth, float* other, float* depth_state, void* stream);

/* Per-level intrinsics of DepthEstimatorPyramid.call (:300-302): out_f[l], out_c[l] ([nlevels,b,2]) = cam / 2^(l+1),
 * l = 0 .. nlevels-1 (exact: power-of-two divisions). */
int m4d_camera_pyramid(const float* cam_f, const float* cam_c, int b, int nlevels, float* out_f, float* out_c,
                       void* stream);

/* fill(n floats) on the stream (state reset: depth_prev_t <- 1000, m4depth_network.py:209) */
int m4d_fill(float* p, int64_t n, float value, void* stream);
/* x [b,h,w,c] (pixel stride x_pix_stride floats) -> the interior of y, a dense [b,h+shift_y,w+shift_x,c] tensor the caller zeroed
 * once, at offset (shift_y, shift_x).  Keras Conv2D(strides=2, padding='same') pads an ODD dimension one pixel on BOTH sides
 * (m4depth_network.py:66-72, SURVEY A.13: 15 -> 8); shifted by one pixel it is the even-sized problem (pad 0 / 1) that the
 * tensor-core stride-2 convolution handles.  c % 4 == 0. */
int m4d_pad_shift(const float* x, int x_pix_stride, int b, int h, int w, int c, int shift_y, int shift_x, float* y, void* stream);

/* ---- metrics (metrics.py:1-64 + clipping m4depth_network.py:465-467) ----------------------------
 * gt, est [n] -> out[7] = AbsRel, SqRel, RMSE, RMSE_log, Delta1, Delta2, Delta3 for this batch
 * (one keras Mean.update_state sample each).  ws: 16 doubles of caller workspace. */
int m4d_depth_metrics(const float* gt, const float* est, int64_t n, float max_d, double* ws, float* out,
                      void* stream);

#ifdef __cplusplus
}
#endif
#endif /* M4D_H_ */
