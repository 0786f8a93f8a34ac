// libm4d: fused backproject + parallax-sweeping cost volume (PSCV), one kernel per pyramid level.
//
// Replaces, in ONE pass over HBM, what utils/depth_operations.py:223-281 does with (2r+1)x tile_in_batch
// copies, dense_image_warp -> BackProject / 4x tf.gather, and the fp16 correlate:
//
//   phase 0a  one thread per pixel            : epipolar terms (alpha, p, delta, s)            -> smem
//   phase 0b  one thread per (pixel, k)       : query point, tap grid + weights               -> smem
//                                               + bilinear sample of para_prev_t (prev_disp) -> HBM
//   phase 1   one thread per (pixel, quad)    : for k in 0..K-1: 4 float4 tap loads of c2 (L1/L2 served,
//                                               a warp reads whole 128-B pixel rows), bilinear in packed
//                                               f32x2, fp16 products with the resident c1 quad, partial sum
//   phase 2   one thread per (pixel, cut, k)  : ordered sum of the group's partials, /n, fp16 round -> HBM
//
// HBM traffic is the algorithmic minimum: c1 and para maps read once, c2 read ~once (tap re-reads hit L1/L2),
// cv / prev_disp written once.  Arithmetic contract (DESIGN.md "Numerics"):
//   geometry    one rounded fp32 op per reference TF op (common.cuh) -> tap grids bit-identical to the oracle
//   interp mode 0 GATHER  utils/dense_image_warp.py:127-190 (TF CPU path): clamped floor, 3 un-fused lerps
//               1 BP      backproject_op_gpu.cu.cc:47-74 with separately rounded products / sums
//               2 BP_FMA  the same expression as nvcc contracts it (mul, fma, fma, fma) = the reference's GPU binary
//   correlate   fp16(c1) * fp16(c2w) rounded to fp16, summed in fp32 (4-channel partials in channel order, then
//               quads in order), / group width (IEEE), rounded once to fp16, widened to fp32.
#include "common.cuh"

namespace {

typedef unsigned long long u64;

// ---- packed fp32x2 arithmetic (FFMA2 on sm_100a).  Every helper is ONE correctly rounded IEEE operation per
// element; mul / add are expressed as fma with -0 / 1 so that ptxas cannot contract neighbouring ops.
__device__ __forceinline__ u64 pk(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) {
  u64 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}

struct __align__(16) TapRec {
  float w[4];      // BP modes: w00,w01,w10,w11 ; GATHER: ax, ay, -, -
  uint32_t i[4];   // float4 index of tap pixel (y0,x0),(y0,x1),(y1,x0),(y1,x1) channel 0; i[0]==kOutside -> sample is 0
};
constexpr uint32_t kOutside = 0xFFFFFFFFu;

struct PixRec {
  Epi e;
  float para_l;
  int x, y, b;     // b < 0: pixel beyond the end of the tensor
};

struct PscvArgs {
  const float *c1, *c2, *para_t, *para_l, *rot, *trans, *cam_f, *cam_c;
  float *cv, *prev_disp, *centre_log;
  int32_t* idx_dbg;
  int rot_dim, b, h, w, c, cuts, r, K, Q, TP;
  int cv_stride, pd_stride, cl_stride;
  float cl_scale;
  int64_t npix;
  // 1, -1 and -0 as RUN-TIME values: ptxas folds fma(fma(a,b,-0),1,c) into fma(a,b,c) when it can see the constants,
  // which removes a rounding the reference performs (observed in SASS; cv then differs by one fp16 ulp in ~1e-4 of
  // the outputs).  Read from the parameter bank, they are opaque to the optimiser.
  float one, neg_one, neg_zero;
};

enum { kGather = 0, kBP = 1, kBPFma = 2 };

// Bilinear sample of a scalar map with the same op order as the channel path.
template <int MODE>
__device__ __forceinline__ float sample_scalar(const float* __restrict__ img, const TapRec& t, int Q) {
  if (t.i[0] == kOutside) return 0.f;
  // tap indices are float4 indices of the c-channel tensor: pixel index = i / Q
  float v00 = __ldg(img + t.i[0] / Q), v01 = __ldg(img + t.i[1] / Q);
  float v10 = __ldg(img + t.i[2] / Q), v11 = __ldg(img + t.i[3] / Q);
  if (MODE == kGather) {
    float top = FADD(FMUL(t.w[0], FSUB(v01, v00)), v00);
    float bot = FADD(FMUL(t.w[0], FSUB(v11, v10)), v10);
    return FADD(FMUL(t.w[1], FSUB(bot, top)), top);
  } else if (MODE == kBP) {
    return FADD(FADD(FADD(FMUL(v00, t.w[0]), FMUL(v01, t.w[1])), FMUL(v10, t.w[2])), FMUL(v11, t.w[3]));
  } else {
    return __fmaf_rn(v11, t.w[3], __fmaf_rn(v10, t.w[2], __fmaf_rn(v01, t.w[1], FMUL(v00, t.w[0]))));
  }
}

template <int MODE>
__global__ void __launch_bounds__(256) pscv_kernel(PscvArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int K = a.K, Q = a.Q, TP = a.TP;
  TapRec* recs = reinterpret_cast<TapRec*>(smem_raw);                       // [TP*K]
  float* part = reinterpret_cast<float*>(recs + TP * K);                    // [TP*Q][K]
  PixRec* pix = reinterpret_cast<PixRec*>(part + TP * Q * K);               // [TP]

  const int tid = threadIdx.x;
  const int64_t pix0 = (int64_t)blockIdx.x * TP;
  const int H = a.h, W = a.w;

  // ---- phase 0a: per-pixel epipolar terms (depth_operations.py:239-259)
  if (tid < TP) {
    PixRec pr;
    int64_t p = pix0 + tid;
    if (p < a.npix) {
      pr.x = (int)(p % W);
      pr.y = (int)((p / W) % H);
      pr.b = (int)(p / ((int64_t)W * H));
      Pose P;
      load_pose(a.rot, a.rot_dim, a.trans, a.cam_f, a.cam_c, pr.b, P);
      pr.e = epipolar(P, pr.x, pr.y);
      pr.para_l = __ldg(a.para_l + p);
    } else {
      pr.b = -1; pr.x = pr.y = 0; pr.para_l = 1.f;
      pr.e = Epi();
    }
    pix[tid] = pr;
  }
  __syncthreads();

  // ---- phase 0b: query point and taps per (pixel, hypothesis) (:229-236, :261-265, dense_image_warp.py:238-253)
  for (int it = tid; it < TP * K; it += blockDim.x) {
    const int pl = it / K, k = it - pl * K;
    const PixRec& pr = pix[pl];
    TapRec t;
    t.i[0] = kOutside; t.i[1] = t.i[2] = t.i[3] = 0;
    t.w[0] = t.w[1] = t.w[2] = t.w[3] = 0.f;
    if (pr.b >= 0) {
      float rho = FADD(pr.para_l, (float)(k - a.r));
      rho = (rho != rho) ? rho : fminf(fmaxf(rho, 1e-6f), 1e6f);          // tf.clip_by_value :236
      const float div = FDIV(pr.e.s, rho);                                  // :262
      const float ex = FDIV(pr.e.dx, div), ey = FDIV(pr.e.dy, div);         // :263
      const float flx = FSUB(FADD(pr.e.px, ex), pr.e.sx);                   // :264
      const float fly = FSUB(FADD(pr.e.py, ey), pr.e.sy);
      const float qy = FADD((float)pr.y, fly), qx = FADD((float)pr.x, flx); // dense_image_warp.py:244
      // BackProject-convention grid (the "integer index grids" of the op); also what BP modes sample with
      const float cqx = clip_keep_nan(qx, (float)(W - 1)), cqy = clip_keep_nan(qy, (float)(H - 1));   // :248
      const Tap bt = make_tap(cqx, cqy, W, H);
      const int64_t p = pix0 + pl;
      if (a.idx_dbg) {
        int4 v = bt.inside ? make_int4(bt.x0, bt.x0 + bt.dxo, bt.y0, bt.y0 + bt.dyo) : make_int4(-1, -1, -1, -1);
        reinterpret_cast<int4*>(a.idx_dbg)[p * K + k] = v;
      }
      const uint32_t img = (uint32_t)pr.b * (uint32_t)(H * W);
      if (MODE == kGather) {
        if (qx == qx && qy == qy) {
          // floor clamped to [0, size-2], ceil = floor + 1, alpha clamped to [0,1] (dense_image_warp.py:135-149)
          const float fx0 = fminf(fmaxf(0.f, floorf(qx)), (float)(W - 2));
          const float fy0 = fminf(fmaxf(0.f, floorf(qy)), (float)(H - 2));
          const int x0 = (int)fx0, y0 = (int)fy0;
          t.w[0] = fminf(fmaxf(FSUB(qx, fx0), 0.f), 1.f);
          t.w[1] = fminf(fmaxf(FSUB(qy, fy0), 0.f), 1.f);
          const uint32_t base = (img + (uint32_t)(y0 * W + x0)) * (uint32_t)Q;
          t.i[0] = base; t.i[1] = base + Q; t.i[2] = base + (uint32_t)W * Q; t.i[3] = base + (uint32_t)(W + 1) * Q;
        }
      } else if (bt.inside) {
        tap_weights(bt.wx, bt.wy, t.w[0], t.w[1], t.w[2], t.w[3]);
        const uint32_t base = (img + (uint32_t)(bt.y0 * W + bt.x0)) * (uint32_t)Q;
        t.i[0] = base;
        t.i[1] = base + (uint32_t)bt.dxo * Q;
        t.i[2] = base + (uint32_t)(bt.dyo * W) * Q;
        t.i[3] = t.i[2] + (uint32_t)bt.dxo * Q;
      }
      // warped previous-frame parallax (:268, :280); only what the caller asked for
      const bool want_pd = a.prev_disp != nullptr;
      const bool want_cl = a.centre_log != nullptr && k == a.r;
      if (want_pd || want_cl) {
        const float pd = sample_scalar<MODE>(a.para_t, t, Q);
        if (want_pd) a.prev_disp[p * a.pd_stride + k] = pd;
        if (want_cl) a.centre_log[p * a.cl_stride] = logf(FMUL(pd, a.cl_scale));     // m4depth_network.py:238
      }
    }
    recs[it] = t;
  }
  __syncthreads();

  // ---- phase 1: gather + bilinear + fp16 products, one (pixel, channel quad) per thread
  const u64 NZ2 = pk(a.neg_zero, a.neg_zero), ONE2 = pk(a.one, a.one), NEG2 = pk(a.neg_one, a.neg_one);
  const float4* __restrict__ c1v = reinterpret_cast<const float4*>(a.c1);
  const float4* __restrict__ c2v = reinterpret_cast<const float4*>(a.c2);
  for (int it = tid; it < TP * Q; it += blockDim.x) {
    const int pl = it / Q, q = it - pl * Q;
    float* my_part = part + (size_t)it * K;
    if (pix[pl].b < 0) {
      for (int k = 0; k < K; ++k) my_part[k] = 0.f;
      continue;
    }
    const float4 cc = __ldg(c1v + (pix0 + pl) * Q + q);
    const __half2 h01 = __floats2half2_rn(cc.x, cc.y), h23 = __floats2half2_rn(cc.z, cc.w);     // :276 tf.cast(c1, fp16)
    const TapRec* rp = recs + pl * K;
#pragma unroll 3
    for (int k = 0; k < K; ++k) {
      const float4 wv = *reinterpret_cast<const float4*>(rp[k].w);
      const uint4 iv = *reinterpret_cast<const uint4*>(rp[k].i);
      float s = 0.f;
      if (iv.x != kOutside) {
        const float4 t00 = __ldg(c2v + iv.x + q), t01 = __ldg(c2v + iv.y + q);
        const float4 t10 = __ldg(c2v + iv.z + q), t11 = __ldg(c2v + iv.w + q);
        const u64 a00 = pk(t00.x, t00.y), b00 = pk(t00.z, t00.w), a01 = pk(t01.x, t01.y), b01 = pk(t01.z, t01.w);
        const u64 a10 = pk(t10.x, t10.y), b10 = pk(t10.z, t10.w), a11 = pk(t11.x, t11.y), b11 = pk(t11.z, t11.w);
        u64 va, vb;
        if (MODE == kGather) {
          const u64 ax = pk(wv.x, wv.x), ay = pk(wv.y, wv.y);
          // top = ax*(TR-TL)+TL ; bot = ax*(BR-BL)+BL ; out = ay*(bot-top)+top, every op rounded (dense_image_warp.py:188-190)
          u64 topa = fma2(fma2(ax, fma2(a00, NEG2, a01), NZ2), ONE2, a00);
          u64 bota = fma2(fma2(ax, fma2(a10, NEG2, a11), NZ2), ONE2, a10);
          va = fma2(fma2(ay, fma2(topa, NEG2, bota), NZ2), ONE2, topa);
          u64 topb = fma2(fma2(ax, fma2(b00, NEG2, b01), NZ2), ONE2, b00);
          u64 botb = fma2(fma2(ax, fma2(b10, NEG2, b11), NZ2), ONE2, b10);
          vb = fma2(fma2(ay, fma2(topb, NEG2, botb), NZ2), ONE2, topb);
        } else {
          const u64 w00 = pk(wv.x, wv.x), w01 = pk(wv.y, wv.y), w10 = pk(wv.z, wv.z), w11 = pk(wv.w, wv.w);
          if (MODE == kBP) {
            va = fma2(fma2(fma2(fma2(a00, w00, NZ2), ONE2, fma2(a01, w01, NZ2)), ONE2, fma2(a10, w10, NZ2)), ONE2, fma2(a11, w11, NZ2));
            vb = fma2(fma2(fma2(fma2(b00, w00, NZ2), ONE2, fma2(b01, w01, NZ2)), ONE2, fma2(b10, w10, NZ2)), ONE2, fma2(b11, w11, NZ2));
          } else {
            va = fma2(a11, w11, fma2(a10, w10, fma2(a01, w01, fma2(a00, w00, NZ2))));
            vb = fma2(b11, w11, fma2(b10, w10, fma2(b01, w01, fma2(b00, w00, NZ2))));
          }
        }
        float v0, v1, v2, v3;
        upk(va, v0, v1);
        upk(vb, v2, v3);
        // fp16 correlate (:276): both operands and the product rounded to fp16; partial sum in fp32
        const float2 p01 = __half22float2(__hmul2(h01, __floats2half2_rn(v0, v1)));
        const float2 p23 = __half22float2(__hmul2(h23, __floats2half2_rn(v2, v3)));
        s = FADD(FADD(FADD(p01.x, p01.y), p23.x), p23.y);
      }
      my_part[k] = s;
    }
  }
  __syncthreads();

  // ---- phase 2: group means (:277) -> cv, cut-major channel = cut*K + k (:278)
  const int gq = Q / a.cuts;
  const float gw = (float)(4 * gq);
  const int per_pix = a.cuts * K;
  for (int it = tid; it < TP * per_pix; it += blockDim.x) {
    const int pl = it / per_pix, rem = it - pl * per_pix;
    if (pix[pl].b < 0) continue;
    const int cut = rem / K, k = rem - cut * K;
    const float* src = part + ((size_t)pl * Q + (size_t)cut * gq) * K + k;
    float acc = src[0];
    for (int j = 1; j < gq; ++j) acc = FADD(acc, src[(size_t)j * K]);
    const float m = __half2float(__float2half_rn(FDIV(acc, gw)));
    a.cv[(pix0 + pl) * a.cv_stride + rem] = m;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" {

int m4d_pscv_fused_fwd_ex(const float* c1, const float* c2, const float* para_prev_t, const float* para_prev_l,
                          const float* rot, int rot_dim, const float* trans, const float* cam_f, const float* cam_c,
                          int b, int h, int w, int c, int cuts, int search_range,
                          float* cv, int cv_pix_stride, float* prev_disp, int pd_pix_stride,
                          float* centre_log, int centre_log_pix_stride, float centre_log_scale,
                          int32_t* idx_dbg, int interp, void* stream) {
  M4D_REQUIRE(c1 && c2 && para_prev_l && rot && trans && cam_f && cam_c && cv, "m4d_pscv_fused_fwd: null pointer");
  M4D_REQUIRE(para_prev_t || (!prev_disp && !centre_log), "m4d_pscv_fused_fwd: para_prev_t is required for prev_disp / centre_log");
  M4D_REQUIRE(b > 0 && h > 0 && w > 0 && c > 0 && cuts > 0, "m4d_pscv_fused_fwd: non-positive size");
  M4D_REQUIRE(rot_dim == 3 || rot_dim == 4, "m4d_pscv_fused_fwd: rot_dim must be 3 or 4");
  M4D_REQUIRE(search_range >= 0 && search_range <= 8, "m4d_pscv_fused_fwd: search_range must be in [0,8] (got %d)", search_range);
  M4D_REQUIRE(c % cuts == 0 && (c / cuts) % 4 == 0, "m4d_pscv_fused_fwd: group width c/cuts must be a multiple of 4 (c=%d cuts=%d)", c, cuts);
  M4D_REQUIRE(interp >= 0 && interp <= 2, "m4d_pscv_fused_fwd: bad interp mode %d", interp);
  M4D_REQUIRE(interp != kGather || (h >= 2 && w >= 2), "m4d_pscv_fused_fwd: the gather convention needs h,w >= 2");
  M4D_REQUIRE(aligned16(c1) && aligned16(c2), "m4d_pscv_fused_fwd: feature maps must be 16-byte aligned");
  M4D_REQUIRE(!idx_dbg || aligned16(idx_dbg), "m4d_pscv_fused_fwd: idx_dbg must be 16-byte aligned");
  const int K = 2 * search_range + 1;
  M4D_REQUIRE(cv_pix_stride >= cuts * K, "m4d_pscv_fused_fwd: cv_pix_stride %d < cuts*K %d", cv_pix_stride, cuts * K);
  M4D_REQUIRE(!prev_disp || pd_pix_stride >= K, "m4d_pscv_fused_fwd: pd_pix_stride too small");
  M4D_REQUIRE(!centre_log || centre_log_pix_stride >= 1, "m4d_pscv_fused_fwd: centre_log_pix_stride must be >= 1");
  const int64_t npix = (int64_t)b * h * w;
  M4D_REQUIRE(npix * (c / 4) < (int64_t)0xFFFFFFFFll, "m4d_pscv_fused_fwd: tensor too large for 32-bit tap indices");

  PscvArgs a;
  a.c1 = c1; a.c2 = c2; a.para_t = para_prev_t; a.para_l = para_prev_l; a.rot = rot; a.trans = trans;
  a.cam_f = cam_f; a.cam_c = cam_c; a.cv = cv; a.prev_disp = prev_disp; a.centre_log = centre_log; a.idx_dbg = idx_dbg;
  a.rot_dim = rot_dim; a.b = b; a.h = h; a.w = w; a.c = c; a.cuts = cuts; a.r = search_range; a.K = K;
  a.Q = c / 4;
  a.TP = 256 / a.Q > 0 ? 256 / a.Q : 1;
  if (a.TP > 64) a.TP = 64;
  a.cv_stride = cv_pix_stride; a.pd_stride = pd_pix_stride; a.cl_stride = centre_log_pix_stride;
  a.cl_scale = centre_log_scale; a.npix = npix;
  a.one = 1.0f; a.neg_one = -1.0f; a.neg_zero = -0.0f;
  const size_t smem = (size_t)a.TP * K * sizeof(TapRec) + (size_t)a.TP * a.Q * K * sizeof(float) + (size_t)a.TP * sizeof(PixRec);
  M4D_REQUIRE(smem <= 200 * 1024, "m4d_pscv_fused_fwd: c=%d needs %zu bytes of shared memory", c, smem);
  const int grid = (int)cdiv64(npix, a.TP);
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
#define M4D_PSCV_LAUNCH(MODE)                                                                                   \
  do {                                                                                                          \
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(pscv_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) pscv_kernel<MODE><<<grid, 256, smem, st>>>(a);                                        \
  } while (0)
  if (interp == kGather) M4D_PSCV_LAUNCH(kGather);
  else if (interp == kBP) M4D_PSCV_LAUNCH(kBP);
  else M4D_PSCV_LAUNCH(kBPFma);
#undef M4D_PSCV_LAUNCH
  if (e != cudaSuccess) {
    m4d_set_error("m4d_pscv_fused_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
    return M4D_ECUDA;
  }
  M4D_CHECK_LAUNCH("m4d_pscv_fused_fwd");
  return M4D_OK;
}

int m4d_pscv_fused_fwd(const float* c1, const float* c2, const float* para_prev_t, const float* para_prev_l,
                       const float* rot, int rot_dim, const float* trans, const float* cam_f, const float* cam_c,
                       int b, int h, int w, int c, int cuts, int search_range,
                       float* cv, int cv_pix_stride, float* prev_disp, int pd_pix_stride,
                       float* centre_log, int centre_log_pix_stride, float centre_log_scale,
                       int32_t* idx_dbg, void* stream) {
  return m4d_pscv_fused_fwd_ex(c1, c2, para_prev_t, para_prev_l, rot, rot_dim, trans, cam_f, cam_c, b, h, w, c, cuts,
                               search_range, cv, cv_pix_stride, prev_disp, pd_pix_stride, centre_log,
                               centre_log_pix_stride, centre_log_scale, idx_dbg, M4D_INTERP_GATHER, stream);
}

}  // extern "C"
