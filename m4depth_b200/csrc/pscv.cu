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
#include "pscv_common.cuh"

namespace {


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
            va = fma2(a11, w11, fma2(a10, w10, fma2(a00, w00, fma2(a01, w01, NZ2))));      // contraction of the reference binary
            vb = fma2(b11, w11, fma2(b10, w10, fma2(b00, w00, fma2(b01, w01, NZ2))));
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


// =====================================================================================================================
// Fast path: search_range 4 (K = 9, the only value the network uses, m4depth_network.py:220) and the pyramid's channel
// counts.  Same arithmetic as the generic kernel above, restructured around what the first ncu capture showed
// (profiles/r1a_pscv_l2_full.md: issue-bound, 82 M warp instructions, L1 hit rate 83 %):
//   * everything that was a run-time division (K, Q, cuts*K) is a compile-time constant or a reciprocal multiply;
//   * 2-D pixel tiles (8x8 at level 2) instead of 1-D runs: the taps of a tile overlap in both directions, so fewer
//     c2 lines are fetched from L2 per SM;
//   * the four taps of the gather convention are one base address plus constant / uniform offsets (x1 = x0+1, y1 = y0+1
//     always, dense_image_warp.py:146), one 16-byte record per (pixel, k);
//   * the k loop is fully unrolled, partial sums stay in registers; lerps are FADD2 / FFMA2 on packed f32x2 and the
//     fp32 sum of the four fp16 products is a chain of mixed-precision FHADD (add.rn.f32.f16), 4 instead of 7 ops;
//   * the tap loads of hypothesis k+1 are issued before the arithmetic of hypothesis k (the second capture,
//     profiles/r1b_pscv9_l2_full.md, showed 68 % of the stall samples on the first use of the taps), and the CTA keeps its
//     shared memory small (15 KB) so that most of the 228 KB stays L1 for the gathers.
template <int C>
struct FastCfg {
  static constexpr int Q = C / 4;
  static constexpr int NT = (Q == 24 || Q == 48) ? 144 : 128;
  static constexpr int PPT = 1;
  static constexpr int TP = PPT * NT / Q;                 // 32, 16, 8, 6, 4, 3 pixels per CTA
  static constexpr int TW = TP == 32 ? 8 : TP == 16 ? 4 : TP == 8 ? 4 : TP == 6 ? 3 : TP == 4 ? 2 : 3;
  static constexpr int TH = TP / TW;
  static_assert(TW * TH == TP, "tile");
};

struct FastPix {
  Epi e;
  float para_l;
  int x, y;          // x < 0: outside the image (partial tile)
  uint32_t p;        // linear pixel index (b*H + y)*W + x
};

struct FastArgs {
  PscvArgs a;
  int tiles_x, tiles_y;
  uint32_t WQ;       // W * Q: float4 distance between image rows
  float inv_per_pix; // 1 / (cuts * 9)
  float inv_gw;      // 1 / group width when it is a power of two, else 0
};

// fast-path record -> the generic TapRec that sample_scalar understands
template <int MODE, int Q>
__device__ __forceinline__ void rec_to_tap(const uint4& ri, const float4& rw, uint32_t WQ, TapRec& t) {
  if (MODE == kGather) {
    const bool valid = ri.w != 0u;
    t.i[0] = valid ? ri.x : kOutside; t.i[1] = ri.x + Q; t.i[2] = ri.x + WQ; t.i[3] = ri.x + WQ + Q;
    t.w[0] = __uint_as_float(ri.y); t.w[1] = __uint_as_float(ri.z); t.w[2] = t.w[3] = 0.f;
  } else {
    const bool valid = ri.w != kOutside;
    t.i[0] = valid ? ri.x : kOutside; t.i[1] = ri.y; t.i[2] = ri.z; t.i[3] = ri.w;
    t.w[0] = rw.x; t.w[1] = rw.y; t.w[2] = rw.z; t.w[3] = rw.w;
  }
}

template <int C, int MODE>
__global__ void __launch_bounds__(FastCfg<C>::NT, 8) pscv9_kernel(FastArgs fa) {
  typedef FastCfg<C> Cfg;
  constexpr int K = 9, R = 4, Q = Cfg::Q, NT = Cfg::NT, TP = Cfg::TP, TW = Cfg::TW, TH = Cfg::TH, PPT = Cfg::PPT;
  constexpr int RECW = (MODE == kGather) ? 1 : 2;            // uint4 words per (pixel, k) record
  const PscvArgs& a = fa.a;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint4* recs = reinterpret_cast<uint4*>(smem_raw);                              // [TP*K*RECW]
  float* part = reinterpret_cast<float*>(recs + TP * K * RECW);                  // [NT*PPT][K]
  FastPix* pix = reinterpret_cast<FastPix*>(part + NT * PPT * K);                // [TP]

  const int tid = threadIdx.x;
  const int H = a.h, W = a.w;
  const int bi = blockIdx.z;
  const int x_base = blockIdx.x * TW, y_base = blockIdx.y * TH;

  // ---- phase 0a: per-pixel epipolar terms
  if (tid < TP) {
    FastPix pr;
    const int x = x_base + tid % TW, y = y_base + tid / TW;
    if (x < W && y < H) {
      pr.x = x; pr.y = y;
      pr.p = (uint32_t)((bi * H + y) * W + x);
      Pose P;
      load_pose(a.rot, a.rot_dim, a.trans, a.cam_f, a.cam_c, bi, P);
      pr.e = epipolar(P, x, y);
      pr.para_l = __ldg(a.para_l + pr.p);
    } else {
      pr.x = -1; pr.y = 0; pr.p = 0; pr.para_l = 1.f;
      pr.e = Epi();
    }
    pix[tid] = pr;
  }
  __syncthreads();

  // ---- phase 0b: query point and taps per (pixel, hypothesis)
  // record: word 0 = {float4 index of tap (y0,x0) for q = 0, ax | idx01, ay | idx10, valid | idx11}; BP modes add the 4 weights.
  // Records of samples that contribute 0 (outside the image / NaN query) point at the image's first pixel so that the
  // loads of phase 1 need no branch; .w == 0 marks them.
  for (int it = tid; it < TP * K; it += NT) {
    const int pl = it / K, k = it - pl * K;
    const FastPix& pr = pix[pl];
    const uint32_t img = (uint32_t)bi * (uint32_t)(H * W);
    uint4 ri = (MODE == kGather) ? make_uint4(img * Q, 0u, 0u, 0u) : make_uint4(img * Q, img * Q, img * Q, kOutside);
    float4 rw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pr.x >= 0) {
      float rho = FADD(pr.para_l, (float)(k - R));
      rho = (rho != rho) ? rho : fminf(fmaxf(rho, 1e-6f), 1e6f);
      const float div = FDIV(pr.e.s, rho);
      const float ex = FDIV(pr.e.dx, div), ey = FDIV(pr.e.dy, div);
      const float flx = FSUB(FADD(pr.e.px, ex), pr.e.sx);
      const float fly = FSUB(FADD(pr.e.py, ey), pr.e.sy);
      const float qy = FADD((float)pr.y, fly), qx = FADD((float)pr.x, flx);
      Tap bt;
      if (MODE != kGather || a.idx_dbg) {
        const float cqx = clip_keep_nan(qx, (float)(W - 1)), cqy = clip_keep_nan(qy, (float)(H - 1));
        bt = make_tap(cqx, cqy, W, H);
        if (a.idx_dbg) {
          int4 v = bt.inside ? make_int4(bt.x0, bt.x0 + bt.dxo, bt.y0, bt.y0 + bt.dyo) : make_int4(-1, -1, -1, -1);
          reinterpret_cast<int4*>(a.idx_dbg)[(size_t)pr.p * K + k] = v;
        }
      }
      if (MODE == kGather) {
        if (qx == qx && qy == qy) {
          const float fx0 = fminf(fmaxf(0.f, floorf(qx)), (float)(W - 2));
          const float fy0 = fminf(fmaxf(0.f, floorf(qy)), (float)(H - 2));
          const float ax = fminf(fmaxf(FSUB(qx, fx0), 0.f), 1.f);
          const float ay = fminf(fmaxf(FSUB(qy, fy0), 0.f), 1.f);
          ri = make_uint4((img + (uint32_t)((int)fy0 * W + (int)fx0)) * (uint32_t)Q, __float_as_uint(ax), __float_as_uint(ay), 1u);
        }
      } else if (bt.inside) {
        tap_weights(bt.wx, bt.wy, rw.x, rw.y, rw.z, rw.w);
        const uint32_t base = (img + (uint32_t)(bt.y0 * W + bt.x0)) * (uint32_t)Q;
        ri.x = base;
        ri.y = base + (uint32_t)bt.dxo * Q;
        ri.z = base + (uint32_t)bt.dyo * fa.WQ;
        ri.w = ri.z + (uint32_t)bt.dxo * Q;
      }
      if (a.prev_disp != nullptr) {                                     // function-level API: all K warped parallaxes (:280)
        TapRec t;
        rec_to_tap<MODE, Q>(ri, rw, fa.WQ, t);
        a.prev_disp[(size_t)pr.p * a.pd_stride + k] = sample_scalar<MODE>(a.para_t, t, Q);
      }
    }
    if (MODE == kGather) {
      recs[it] = ri;
    } else {
      recs[2 * it] = ri;
      recs[2 * it + 1] = make_uint4(__float_as_uint(rw.x), __float_as_uint(rw.y), __float_as_uint(rw.z), __float_as_uint(rw.w));
    }
  }
  __syncthreads();

  // ---- phase 1: one (pixel, channel quad) item per thread and pass; all 9 hypotheses unrolled
  const u64 NZ2 = pk(a.neg_zero, a.neg_zero);
  const float4* __restrict__ c1v = reinterpret_cast<const float4*>(a.c1);
  const unsigned char* __restrict__ c2b = reinterpret_cast<const unsigned char*>(a.c2);
  const uint32_t row_bytes = fa.WQ * 16u;
#pragma unroll 1
  for (int pp = 0; pp < PPT; ++pp) {
    const int item = tid + pp * NT;
    const int pl = item / Q, q = item - pl * Q;
    if (pix[pl].x < 0) continue;
    const float4 cc = __ldg(c1v + (size_t)pix[pl].p * Q + q);
    // thread's 64-bit base pointer kept opaque so that each tap address is ONE IMAD.WIDE (idx * 16 + base)
    unsigned long long c2q_;
    asm volatile("add.u64 %0, %1, %2;" : "=l"(c2q_) : "l"(reinterpret_cast<unsigned long long>(c2b)), "l"((unsigned long long)(q * 16)));
    const unsigned char* __restrict__ c2q = reinterpret_cast<const unsigned char*>(c2q_);
    const uint32_t rp = (uint32_t)__cvta_generic_to_shared(recs + pl * K * RECW);
    float* my_part = part + item * K;

    uint4 iv[2];
    float4 T[2][4];
    // issue the record read and the four tap loads of hypothesis k into buffer k & 1
#define M4D_LOAD_TAPS(k)                                                                                       \
    do {                                                                                                       \
      uint4& v_ = iv[(k) & 1];                                                                                 \
      float4* t_ = T[(k) & 1];                                                                                 \
      v_ = lds128(rp + (k) * RECW * 16);                                                                       \
      if (MODE == kGather) {                                                                                   \
        const unsigned char* p0 = c2q + (size_t)v_.x * 16u;                                                    \
        const unsigned char* p1 = p0 + row_bytes;                                                              \
        t_[0] = __ldg(reinterpret_cast<const float4*>(p0)); t_[1] = __ldg(reinterpret_cast<const float4*>(p0 + C * 4)); \
        t_[2] = __ldg(reinterpret_cast<const float4*>(p1)); t_[3] = __ldg(reinterpret_cast<const float4*>(p1 + C * 4)); \
      } else {                                                                                                 \
        const uint32_t i11 = v_.w != kOutside ? v_.w : v_.x;                                                   \
        t_[0] = __ldg(reinterpret_cast<const float4*>(c2q + (size_t)v_.x * 16u));                              \
        t_[1] = __ldg(reinterpret_cast<const float4*>(c2q + (size_t)v_.y * 16u));                              \
        t_[2] = __ldg(reinterpret_cast<const float4*>(c2q + (size_t)v_.z * 16u));                              \
        t_[3] = __ldg(reinterpret_cast<const float4*>(c2q + (size_t)i11 * 16u));                               \
      }                                                                                                        \
    } while (0)

    M4D_LOAD_TAPS(0);
    const __half2 h01 = __floats2half2_rn(cc.x, cc.y), h23 = __floats2half2_rn(cc.z, cc.w);     // :276 tf.cast(c1, fp16)
#pragma unroll
    for (int k = 0; k < K; ++k) {
      if (k + 1 < K) M4D_LOAD_TAPS(k + 1);
      const uint4 v = iv[k & 1];
      const float4 t00 = T[k & 1][0], t01 = T[k & 1][1], t10 = T[k & 1][2], t11 = T[k & 1][3];
      const bool valid = (MODE == kGather) ? (v.w != 0u) : (v.w != kOutside);
      const u64 a00 = pk(t00.x, t00.y), b00 = pk(t00.z, t00.w), a01 = pk(t01.x, t01.y), b01 = pk(t01.z, t01.w);
      const u64 a10 = pk(t10.x, t10.y), b10 = pk(t10.z, t10.w), a11 = pk(t11.x, t11.y), b11 = pk(t11.z, t11.w);
      u64 va, vb;
      if (MODE == kGather) {
        const float axf = __uint_as_float(v.y), ayf = __uint_as_float(v.z);
        const u64 ax = pk(axf, axf), ay = pk(ayf, ayf);
        // top = ax*(TR-TL)+TL ; bot = ax*(BR-BL)+BL ; out = ay*(bot-top)+top, every op rounded (dense_image_warp.py:188-190).
        // The multiply is fma(x, y, -0) with -0 read from the parameter bank: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
        // into FFMA2 (observed in SASS, even with -fmad=false), which would drop a rounding the reference performs.
        const u64 topa = add2(fma2(ax, sub2(a01, a00), NZ2), a00);
        const u64 bota = add2(fma2(ax, sub2(a11, a10), NZ2), a10);
        va = add2(fma2(ay, sub2(bota, topa), NZ2), topa);
        const u64 topb = add2(fma2(ax, sub2(b01, b00), NZ2), b00);
        const u64 botb = add2(fma2(ax, sub2(b11, b10), NZ2), b10);
        vb = add2(fma2(ay, sub2(botb, topb), NZ2), topb);
      } else {
        const uint4 wv = lds128(rp + (k * RECW + 1) * 16);
        const float w0 = __uint_as_float(wv.x), w1 = __uint_as_float(wv.y), w2 = __uint_as_float(wv.z), w3 = __uint_as_float(wv.w);
        const u64 w00 = pk(w0, w0), w01 = pk(w1, w1), w10 = pk(w2, w2), w11 = pk(w3, w3);
        if (MODE == kBP) {
          va = add2(add2(add2(fma2(a00, w00, NZ2), fma2(a01, w01, NZ2)), fma2(a10, w10, NZ2)), fma2(a11, w11, NZ2));
          vb = add2(add2(add2(fma2(b00, w00, NZ2), fma2(b01, w01, NZ2)), fma2(b10, w10, NZ2)), fma2(b11, w11, NZ2));
        } else {
          va = fma2(a11, w11, fma2(a10, w10, fma2(a00, w00, fma2(a01, w01, NZ2))));      // contraction of the reference binary
          vb = fma2(b11, w11, fma2(b10, w10, fma2(b00, w00, fma2(b01, w01, NZ2))));
        }
      }
      float v0, v1, v2, v3;
      upk(va, v0, v1);
      upk(vb, v2, v3);
      const __half2 p01 = __hmul2(h01, __floats2half2_rn(v0, v1));      // :276 fp16 operands, fp16 product
      const __half2 p23 = __hmul2(h23, __floats2half2_rn(v2, v3));
      const float sk = sum4_h(h2_bits(p01), h2_bits(p23));
      my_part[k] = valid ? sk : 0.f;
    }
#undef M4D_LOAD_TAPS
  }
  __syncthreads();

  // ---- phase 2: group means -> cv, cut-major channel = cut*K + k
  const int gq = Q / a.cuts;
  const float gw = (float)(4 * gq);
  const int per_pix = a.cuts * K;
  for (int it = tid; it < TP * per_pix; it += NT) {
    const int pl = (int)(((float)it + 0.5f) * fa.inv_per_pix);
    const int rem = it - pl * per_pix;
    if (pix[pl].x < 0) continue;
    const int cut = rem / K, k = rem - cut * K;
    const float* src = part + (pl * Q + cut * gq) * K + k;
    float acc = src[0];
    for (int j = 1; j < gq; ++j) acc = FADD(acc, src[j * K]);
    const float mean = fa.inv_gw != 0.f ? FMUL(acc, fa.inv_gw) : FDIV(acc, gw);
    a.cv[(size_t)pix[pl].p * a.cv_stride + rem] = __half2float(__float2half_rn(mean));
  }
  // ---- log of the centre hypothesis' warped previous parallax (m4depth_network.py:238), one thread per pixel
  if (a.centre_log != nullptr && tid < TP && pix[tid].x >= 0) {
    const uint4 ri = recs[(tid * K + R) * RECW];
    float4 rw = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE != kGather) {
      const uint4 wv = recs[(tid * K + R) * RECW + 1];
      rw = make_float4(__uint_as_float(wv.x), __uint_as_float(wv.y), __uint_as_float(wv.z), __uint_as_float(wv.w));
    }
    TapRec t;
    rec_to_tap<MODE, Q>(ri, rw, fa.WQ, t);
    const float pd = sample_scalar<MODE>(a.para_t, t, Q);
    a.centre_log[(size_t)pix[tid].p * a.cl_stride] = logf(FMUL(pd, a.cl_scale));
  }
}

template <int C>
static cudaError_t launch_fast(const FastArgs& fa, int interp, cudaStream_t st) {
  typedef FastCfg<C> Cfg;
  const PscvArgs& a = fa.a;
  const dim3 grid(fa.tiles_x, fa.tiles_y, a.b);
  const size_t rec_bytes = (size_t)Cfg::TP * 9 * 16;
  const size_t rest = (size_t)Cfg::NT * Cfg::PPT * 9 * sizeof(float) + (size_t)Cfg::TP * sizeof(FastPix);
  cudaError_t e = cudaSuccess;
#define M4D_PSCV9_LAUNCH(MODE, RECW)                                                                            \
  do {                                                                                                          \
    const size_t smem = rec_bytes * (RECW) + rest;                                                              \
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(pscv9_kernel<C, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) pscv9_kernel<C, MODE><<<grid, Cfg::NT, smem, st>>>(fa);                               \
  } while (0)
  if (interp == kGather) M4D_PSCV9_LAUNCH(kGather, 1);
  else if (interp == kBP) M4D_PSCV9_LAUNCH(kBP, 2);
  else M4D_PSCV9_LAUNCH(kBPFma, 2);
#undef M4D_PSCV9_LAUNCH
  return e;
}

// =====================================================================================================================
// Warp-autonomous kernel: same arithmetic, no CTA-wide barriers.  What the captures of pscv9_kernel showed
// (profiles/r1b_pscv9_l2_full.md): 37 % of the warp instructions were geometry executed by mostly-empty warps (16 of
// 128 threads active in phase 0a, a half-empty second pass in 0b), 18 % the group means with run-time index math, and
// 26 % of the stall samples sat on the three __syncthreads.  Here one WARP owns an 8x4-pixel tile end to end:
//   phase 0   lane = pixel: epipolar terms, then the 9 query points / tap records (all lanes busy), records -> the warp's
//             own shared-memory slice; centre_log / prev_disp / idx_dbg written from here
//   phase 1   lane = (pixel, channel quad) items, Q passes of 32 items; per hypothesis one broadcast LDS.128 of the
//             record, four LDG.E.128 taps (a pass reads whole pixel rows: 128 B per quarter warp at c = 32), the un-fused
//             lerps in packed f32x2, fp16 products, FHADD sum; taps are prefetched TWO hypotheses ahead through three
//             register buffers (9 = 3 x 3, so the rotation is static across passes)
//   phase 2   every PH pixels: ordered group sums from the warp's partial-sum slice, mean, fp16 round, coalesced stores
// Only __syncwarp() separates the phases.  The grid is persistent (a multiple of the SM count); a CTA walks 16x8-pixel
// super tiles (its four warps take the four 8x4 quadrants) so that concurrently gathered c2 rows share L1 lines.
template <int C, int VAR>
struct WCfg {
  static constexpr int Q = C / 4;
  static constexpr int DIST = (VAR == 1 || VAR == 2) ? 1 : 2;                  // tap prefetch distance (hypotheses)
  static constexpr int MINB = VAR == 0 ? 3 : VAR == 1 ? 5 : VAR == 2 ? 6 : 4;  // resident CTAs per SM the registers allow
  static constexpr int PH0 = Q <= 4 ? 32 : Q == 8 ? 16 : Q == 16 ? 8 : Q == 24 ? 4 : Q == 32 ? 4 : 2;
  static constexpr int PH = (VAR != 0 && Q <= 8) ? PH0 / 2 : PH0;              // pixels per phase-2 round
  static constexpr int ROUND_PASSES = PH * Q / 32;
  static_assert(PH * Q % 32 == 0 && 32 % PH == 0, "round");
  static constexpr int part_bytes = ROUND_PASSES * 32 * 9 * 4;
  __host__ __device__ static constexpr int warp_bytes(int recw) { return 32 * 9 * 16 * recw + part_bytes; }
};

struct WArgs {
  PscvArgs a;
  int tiles_x, tiles_y;   // 8x4-pixel warp tiles
  int sup_x, sup_y;       // super tiles (2x2 warp tiles) per image
  int n_sup;              // sup_x * sup_y * b
  uint32_t WQ;
};

template <int C, int CUTS, int MODE, int VAR>
__global__ void __launch_bounds__(128, WCfg<C, VAR>::MINB) pscv9w_kernel(WArgs wa) {
  typedef WCfg<C, VAR> Cfg;
  constexpr int DIST = Cfg::DIST;
  constexpr int K = 9, R = 4, Q = Cfg::Q, GQ = Q / CUTS, PH = Cfg::PH, RP = Cfg::ROUND_PASSES;
  constexpr int RECW = (MODE == kGather) ? 1 : 2;
  constexpr int REC_BYTES = 32 * K * 16 * RECW;
  static_assert(Q % CUTS == 0, "groups");
  const PscvArgs& a = wa.a;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned char* wbase = smem_raw + warp * Cfg::warp_bytes(RECW);
  uint4* recs = reinterpret_cast<uint4*>(wbase);                      // [32 pixels][K][RECW]
  float* part = reinterpret_cast<float*>(wbase + REC_BYTES);          // [RP * 32 items][K]
  const uint32_t recs_s = (uint32_t)__cvta_generic_to_shared(recs);
  const int H = a.h, W = a.w;
  const u64 NZ2 = pk(a.neg_zero, a.neg_zero);
  const float4* __restrict__ c1v = reinterpret_cast<const float4*>(a.c1);
  const unsigned char* __restrict__ c2b = reinterpret_cast<const unsigned char*>(a.c2);
  const uint32_t row_bytes = wa.WQ * 16u;
  const int sup_per_img = wa.sup_x * wa.sup_y;

#pragma unroll 1
  for (int sup = blockIdx.x; sup < wa.n_sup; sup += gridDim.x) {
    const int bi = sup / sup_per_img;
    const int r0 = sup - bi * sup_per_img;
    const int sy = r0 / wa.sup_x, sx = r0 - sy * wa.sup_x;
    const int tx = sx * 2 + (warp & 1), ty = sy * 2 + (warp >> 1);
    if (tx >= wa.tiles_x || ty >= wa.tiles_y) continue;              // warp-uniform
    const int x_base = tx * 8, y_base = ty * 4;
    const uint32_t img = (uint32_t)bi * (uint32_t)(H * W);

    // ---- phase 0: lane = pixel (depth_operations.py:229-265, dense_image_warp.py:135-149 / 238-253)
    {
      const int x = x_base + (lane & 7), y = y_base + (lane >> 3);
      const bool inimg = x < W && y < H;
      const uint32_t p = img + (uint32_t)(y * W + x);
      Pose P;
      load_pose(a.rot, a.rot_dim, a.trans, a.cam_f, a.cam_c, bi, P);
      const Epi e = epipolar(P, x, y);
      const float para_l = inimg ? __ldg(a.para_l + p) : 1.f;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        uint4 ri = (MODE == kGather) ? make_uint4(img * Q, 0u, 0u, 0u) : make_uint4(img * Q, img * Q, img * Q, kOutside);
        float4 rw = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inimg) {
          float rho = FADD(para_l, (float)(k - R));
          rho = (rho != rho) ? rho : fminf(fmaxf(rho, 1e-6f), 1e6f);
          const float div = FDIV(e.s, rho);
          const float ex = FDIV(e.dx, div), ey = FDIV(e.dy, div);
          const float flx = FSUB(FADD(e.px, ex), e.sx);
          const float fly = FSUB(FADD(e.py, ey), e.sy);
          const float qy = FADD((float)y, fly), qx = FADD((float)x, flx);
          Tap bt;
          if (MODE != kGather || a.idx_dbg) {
            const float cqx = clip_keep_nan(qx, (float)(W - 1)), cqy = clip_keep_nan(qy, (float)(H - 1));
            bt = make_tap(cqx, cqy, W, H);
            if (a.idx_dbg) {
              int4 v = bt.inside ? make_int4(bt.x0, bt.x0 + bt.dxo, bt.y0, bt.y0 + bt.dyo) : make_int4(-1, -1, -1, -1);
              reinterpret_cast<int4*>(a.idx_dbg)[(size_t)p * K + k] = v;
            }
          }
          if (MODE == kGather) {
            if (qx == qx && qy == qy) {
              const float fx0 = fminf(fmaxf(0.f, floorf(qx)), (float)(W - 2));
              const float fy0 = fminf(fmaxf(0.f, floorf(qy)), (float)(H - 2));
              const float ax = fminf(fmaxf(FSUB(qx, fx0), 0.f), 1.f);
              const float ay = fminf(fmaxf(FSUB(qy, fy0), 0.f), 1.f);
              ri = make_uint4((img + (uint32_t)((int)fy0 * W + (int)fx0)) * (uint32_t)Q, __float_as_uint(ax), __float_as_uint(ay), 1u);
            }
          } else if (bt.inside) {
            tap_weights(bt.wx, bt.wy, rw.x, rw.y, rw.z, rw.w);
            const uint32_t base = (img + (uint32_t)(bt.y0 * W + bt.x0)) * (uint32_t)Q;
            ri.x = base;
            ri.y = base + (uint32_t)bt.dxo * Q;
            ri.z = base + (uint32_t)bt.dyo * wa.WQ;
            ri.w = ri.z + (uint32_t)bt.dxo * Q;
          }
          const bool want_pd = a.prev_disp != nullptr;
          const bool want_cl = a.centre_log != nullptr && k == R;
          if (want_pd || want_cl) {                                    // warped previous parallax (:268, :280)
            TapRec t;
            rec_to_tap<MODE, Q>(ri, rw, wa.WQ, t);
            const float pd = sample_scalar<MODE>(a.para_t, t, Q);
            if (want_pd) a.prev_disp[(size_t)p * a.pd_stride + k] = pd;
            if (want_cl) a.centre_log[(size_t)p * a.cl_stride] = logf(FMUL(pd, a.cl_scale));     // m4depth_network.py:238
          }
        }
        recs[(lane * K + k) * RECW] = ri;
        if (MODE != kGather)
          recs[(lane * K + k) * RECW + 1] = make_uint4(__float_as_uint(rw.x), __float_as_uint(rw.y), __float_as_uint(rw.z), __float_as_uint(rw.w));
      }
    }
    __syncwarp();

    // ---- phases 1 and 2
    uint4 iv[3];
    float4 T[3][4];
    // (record, taps) of hypothesis k_ of the item whose record slice starts at rp_ and whose quad pointer is cq_ -> buffer b_
#define M4D_W_LOAD(b_, rp_, cq_, k_)                                                                           \
    do {                                                                                                       \
      uint4& v_ = iv[b_];                                                                                      \
      float4* t_ = T[b_];                                                                                      \
      v_ = lds128((rp_) + (k_) * RECW * 16);                                                                   \
      if (MODE == kGather) {                                                                                   \
        const unsigned char* p0 = (cq_) + (size_t)v_.x * 16u;                                                  \
        const unsigned char* p1 = p0 + row_bytes;                                                              \
        t_[0] = __ldg(reinterpret_cast<const float4*>(p0)); t_[1] = __ldg(reinterpret_cast<const float4*>(p0 + C * 4)); \
        t_[2] = __ldg(reinterpret_cast<const float4*>(p1)); t_[3] = __ldg(reinterpret_cast<const float4*>(p1 + C * 4)); \
      } else {                                                                                                 \
        const uint32_t i11 = v_.w != kOutside ? v_.w : v_.x;                                                   \
        t_[0] = __ldg(reinterpret_cast<const float4*>((cq_) + (size_t)v_.x * 16u));                            \
        t_[1] = __ldg(reinterpret_cast<const float4*>((cq_) + (size_t)v_.y * 16u));                            \
        t_[2] = __ldg(reinterpret_cast<const float4*>((cq_) + (size_t)v_.z * 16u));                            \
        t_[3] = __ldg(reinterpret_cast<const float4*>((cq_) + (size_t)i11 * 16u));                             \
      }                                                                                                        \
    } while (0)

    // item of pass gp: (pixel pl, quad q); its c1 quad, record slice and c2 quad pointer
    int pl = lane / Q, q = lane - pl * Q;
    uint32_t rp = recs_s + (uint32_t)(pl * K * RECW * 16);
    const unsigned char* cq = c2b + q * 16;
    auto c1_of = [&](int pl_, int q_) -> float4 {
      const int xx = min(x_base + (pl_ & 7), W - 1), yy = min(y_base + (pl_ >> 3), H - 1);
      return __ldg(c1v + (size_t)(img + (uint32_t)(yy * W + xx)) * Q + q_);
    };
    float4 cc = c1_of(pl, q);
    M4D_W_LOAD(0, rp, cq, 0);
    if (DIST == 2) M4D_W_LOAD(1, rp, cq, 1);
#pragma unroll 1
    for (int gp = 0; gp < Q; ++gp) {
      // next pass' item (prefetched from k = 7 on)
      const int item_n = (gp + 1) * 32 + lane;
      const int pl_n = item_n / Q, q_n = item_n - pl_n * Q;
      const uint32_t rp_n = recs_s + (uint32_t)(pl_n * K * RECW * 16);
      const unsigned char* cq_n = c2b + q_n * 16;
      const bool more = gp + 1 < Q;
      const __half2 h01 = __floats2half2_rn(cc.x, cc.y), h23 = __floats2half2_rn(cc.z, cc.w);     // :276 tf.cast(c1, fp16)
      if (more) cc = c1_of(pl_n, q_n);
      float* my_part = part + ((gp % RP) * 32 + lane) * K;
#pragma unroll
      for (int k = 0; k < K; ++k) {
        if (k + DIST < K) M4D_W_LOAD((k + DIST) % 3, rp, cq, k + DIST);
        else if (more) M4D_W_LOAD((k + DIST) % 3, rp_n, cq_n, k + DIST - K);
        const uint4 v = iv[k % 3];
        const float4 t00 = T[k % 3][0], t01 = T[k % 3][1], t10 = T[k % 3][2], t11 = T[k % 3][3];
        const bool valid = (MODE == kGather) ? (v.w != 0u) : (v.w != kOutside);
        const u64 a00 = pk(t00.x, t00.y), b00 = pk(t00.z, t00.w), a01 = pk(t01.x, t01.y), b01 = pk(t01.z, t01.w);
        const u64 a10 = pk(t10.x, t10.y), b10 = pk(t10.z, t10.w), a11 = pk(t11.x, t11.y), b11 = pk(t11.z, t11.w);
        u64 va, vb;
        if (MODE == kGather) {
          const float axf = __uint_as_float(v.y), ayf = __uint_as_float(v.z);
          const u64 ax = pk(axf, axf), ay = pk(ayf, ayf);
          // top = ax*(TR-TL)+TL ; bot = ax*(BR-BL)+BL ; out = ay*(bot-top)+top, every op rounded (dense_image_warp.py:188-190);
          // the multiply is fma(x, y, -0) with an opaque -0 so that ptxas cannot contract it with the following add.
          const u64 topa = add2(fma2(ax, sub2(a01, a00), NZ2), a00);
          const u64 bota = add2(fma2(ax, sub2(a11, a10), NZ2), a10);
          va = add2(fma2(ay, sub2(bota, topa), NZ2), topa);
          const u64 topb = add2(fma2(ax, sub2(b01, b00), NZ2), b00);
          const u64 botb = add2(fma2(ax, sub2(b11, b10), NZ2), b10);
          vb = add2(fma2(ay, sub2(botb, topb), NZ2), topb);
        } else {
          const uint4 wv = lds128(rp + (k * RECW + 1) * 16);
          const float w0 = __uint_as_float(wv.x), w1 = __uint_as_float(wv.y), w2 = __uint_as_float(wv.z), w3 = __uint_as_float(wv.w);
          const u64 w00 = pk(w0, w0), w01 = pk(w1, w1), w10 = pk(w2, w2), w11 = pk(w3, w3);
          if (MODE == kBP) {
            va = add2(add2(add2(fma2(a00, w00, NZ2), fma2(a01, w01, NZ2)), fma2(a10, w10, NZ2)), fma2(a11, w11, NZ2));
            vb = add2(add2(add2(fma2(b00, w00, NZ2), fma2(b01, w01, NZ2)), fma2(b10, w10, NZ2)), fma2(b11, w11, NZ2));
          } else {
            va = fma2(a11, w11, fma2(a10, w10, fma2(a00, w00, fma2(a01, w01, NZ2))));      // contraction of the reference binary
            vb = fma2(b11, w11, fma2(b10, w10, fma2(b00, w00, fma2(b01, w01, NZ2))));
          }
        }
        float v0, v1, v2, v3;
        upk(va, v0, v1);
        upk(vb, v2, v3);
        const __half2 p01 = __hmul2(h01, __floats2half2_rn(v0, v1));      // :276 fp16 operands, fp16 product
        const __half2 p23 = __hmul2(h23, __floats2half2_rn(v2, v3));
        const float sk = sum4_h(h2_bits(p01), h2_bits(p23));
        my_part[k] = valid ? sk : 0.f;
      }
      rp = rp_n; cq = cq_n;
      if ((gp + 1) % RP == 0) {
        // ---- phase 2: group means of the PH pixels just finished -> cv, cut-major channel = cut*K + k (:277-278)
        __syncwarp();
        constexpr int OUTS = PH * CUTS * K;
        constexpr int GW = C / CUTS;
        const int pl0 = (gp / RP) * PH;
#pragma unroll
        for (int j = 0; j < (OUTS + 31) / 32; ++j) {
          const int it = j * 32 + lane;
          if (OUTS % 32 == 0 || it < OUTS) {
            const int plr = it / (CUTS * K), rem = it - plr * (CUTS * K);
            const int cut = rem / K, k = rem - cut * K;
            const float* src = part + (plr * Q + cut * GQ) * K + k;
            float acc = src[0];
#pragma unroll
            for (int g = 1; g < GQ; ++g) acc = FADD(acc, src[g * K]);
            const float mean = (GW & (GW - 1)) == 0 ? FMUL(acc, 1.0f / (float)GW) : FDIV(acc, (float)GW);
            const int px = x_base + ((pl0 + plr) & 7), py = y_base + ((pl0 + plr) >> 3);
            if (px < W && py < H)
              a.cv[(size_t)(img + (uint32_t)(py * W + px)) * a.cv_stride + rem] = __half2float(__float2half_rn(mean));
          }
        }
        __syncwarp();
      }
    }
#undef M4D_W_LOAD
    __syncwarp();
  }
}

template <int C, int CUTS, int VAR>
static cudaError_t launch_warp(WArgs& wa, int interp, int ctas_per_sm, cudaStream_t st) {
  typedef WCfg<C, VAR> Cfg;
  if (ctas_per_sm <= 0) ctas_per_sm = Cfg::MINB;
  const PscvArgs& a = wa.a;
  wa.tiles_x = (a.w + 7) / 8;
  wa.tiles_y = (a.h + 3) / 4;
  wa.sup_x = (wa.tiles_x + 1) / 2;
  wa.sup_y = (wa.tiles_y + 1) / 2;
  wa.n_sup = wa.sup_x * wa.sup_y * a.b;
  cudaError_t e = cudaSuccess;
#define M4D_PSCVW_LAUNCH(MODE, RECW)                                                                            \
  do {                                                                                                          \
    const size_t smem = 4 * (size_t)Cfg::warp_bytes(RECW);                                                      \
    int grid = m4d_sm_count() * ctas_per_sm;                                                                    \
    if (grid > wa.n_sup) grid = wa.n_sup;                                                                       \
    e = cudaFuncSetAttribute(pscv9w_kernel<C, CUTS, MODE, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e == cudaSuccess) pscv9w_kernel<C, CUTS, MODE, VAR><<<grid, 128, smem, st>>>(wa);                       \
  } while (0)
  if (interp == kGather) M4D_PSCVW_LAUNCH(kGather, 1);
  else if (interp == kBP) M4D_PSCVW_LAUNCH(kBP, 2);
  else M4D_PSCVW_LAUNCH(kBPFma, 2);
#undef M4D_PSCVW_LAUNCH
  return e;
}

// =====================================================================================================================
// Backward of the fused PSCV (gather convention), for train_step (m4depth_network.py:371-399): what TensorFlow's autodiff
// makes of utils/depth_operations.py:223-281 + utils/dense_image_warp.py:127-190.  One thread per pixel, serial over the
// hypotheses and channels (a parity implementation, not a tuned one): d_c1 and d_para_prev_l are thread-private sums
// (deterministic); d_c2 and d_para_prev_t are scattered onto the four taps with atomics, as BackProjectGrad does.
// Gradient arithmetic follows the forward's types: cv = fp32(fp16(sum / n)) -> the incoming gradient is rounded to fp16,
// divided by n in fp32, rounded to fp16 again and multiplied in fp16 with the fp16 operands (:276-277); the warp part is
// fp32.  floor() and the integer taps carry no gradient; clip ops pass it inside their range (boundaries included).
struct PscvBwdArgs {
  PscvArgs a;
  const float *d_cv, *d_pd;
  float *d_c1, *d_c2, *d_pt, *d_pl;
  int dcv_stride, dpd_stride;
};

__global__ void __launch_bounds__(128) pscv_bwd_kernel(PscvBwdArgs g) {
  const PscvArgs& a = g.a;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= a.npix) return;
  const int H = a.h, W = a.w, C = a.c, K = a.K;
  const int x = (int)(p % W), y = (int)((p / W) % H), bi = (int)(p / ((int64_t)W * H));
  Pose P;
  load_pose(a.rot, a.rot_dim, a.trans, a.cam_f, a.cam_c, bi, P);
  const Epi e = epipolar(P, x, y);
  const float para_l = __ldg(a.para_l + p);
  const int gw = C / a.cuts;
  const int64_t img = (int64_t)bi * H * W;
  constexpr int KMAX = 17;
  int64_t base[KMAX];
  float ax[KMAX], ay[KMAX], div_[KMAX], rho_[KMAX], dax[KMAX], day[KMAX];
  bool ok[KMAX], passx[KMAX], passy[KMAX], passr[KMAX];
  for (int k = 0; k < K; ++k) {
    const float raw = FADD(para_l, (float)(k - a.r));
    float rho = (raw != raw) ? raw : fminf(fmaxf(raw, 1e-6f), 1e6f);
    passr[k] = raw >= 1e-6f && raw <= 1e6f;
    const float div = FDIV(e.s, rho);
    const float ex = FDIV(e.dx, div), ey = FDIV(e.dy, div);
    const float qx = FADD((float)x, FSUB(FADD(e.px, ex), e.sx)), qy = FADD((float)y, FSUB(FADD(e.py, ey), e.sy));
    ok[k] = qx == qx && qy == qy;
    const float fx0 = fminf(fmaxf(0.f, floorf(qx)), (float)(W - 2)), fy0 = fminf(fmaxf(0.f, floorf(qy)), (float)(H - 2));
    const float rx = FSUB(qx, fx0), ry = FSUB(qy, fy0);
    ax[k] = fminf(fmaxf(rx, 0.f), 1.f);
    ay[k] = fminf(fmaxf(ry, 0.f), 1.f);
    passx[k] = rx >= 0.f && rx <= 1.f;
    passy[k] = ry >= 0.f && ry <= 1.f;
    base[k] = ok[k] ? img + (int64_t)fy0 * W + (int64_t)fx0 : img;
    div_[k] = div; rho_[k] = rho;
    dax[k] = day[k] = 0.f;
  }
  // feature channels: d_c1 (private), d_c2 (scatter), d alpha
  for (int j = 0; j < C; ++j) {
    const float c1j = __ldg(a.c1 + p * C + j);
    const __half c1h = __float2half_rn(c1j);
    const int cut = j / gw;
    float acc = 0.f;
    for (int k = 0; k < K; ++k) {
      if (!ok[k]) continue;
      const float gin = __ldg(g.d_cv + p * g.dcv_stride + cut * K + k);
      const __half gp = __float2half_rn(FDIV(__half2float(__float2half_rn(gin)), (float)gw));
      const int64_t i00 = base[k] * C + j;
      const float tl = __ldg(a.c2 + i00), tr = __ldg(a.c2 + i00 + C), bl = __ldg(a.c2 + i00 + (int64_t)W * C), br = __ldg(a.c2 + i00 + (int64_t)(W + 1) * C);
      const float top = FADD(FMUL(ax[k], FSUB(tr, tl)), tl), bot = FADD(FMUL(ax[k], FSUB(br, bl)), bl);
      const float sj = FADD(FMUL(ay[k], FSUB(bot, top)), top);
      acc += __half2float(__hmul(gp, __float2half_rn(sj)));
      const float ds = __half2float(__hmul(gp, c1h));
      const float dbot = ds * ay[k], dtop = ds - dbot;
      const float dtr = dtop * ax[k], dbr = dbot * ax[k];
      atomicAdd(g.d_c2 + i00, dtop - dtr);
      atomicAdd(g.d_c2 + i00 + C, dtr);
      atomicAdd(g.d_c2 + i00 + (int64_t)W * C, dbot - dbr);
      atomicAdd(g.d_c2 + i00 + (int64_t)(W + 1) * C, dbr);
      dax[k] += dtop * (tr - tl) + dbot * (br - bl);
      day[k] += ds * (bot - top);
    }
    g.d_c1[p * C + j] = acc;
  }
  // the warped previous parallax (prev_disp, :280) and the chain back to the parallax of this level
  float dpl = 0.f;
  for (int k = 0; k < K; ++k) {
    if (!ok[k]) continue;
    if (g.d_pd != nullptr) {
      const float ds = __ldg(g.d_pd + p * g.dpd_stride + k);
      const int64_t i00 = base[k];
      const float tl = __ldg(a.para_t + i00), tr = __ldg(a.para_t + i00 + 1), bl = __ldg(a.para_t + i00 + W), br = __ldg(a.para_t + i00 + W + 1);
      const float top = FADD(FMUL(ax[k], FSUB(tr, tl)), tl), bot = FADD(FMUL(ax[k], FSUB(br, bl)), bl);
      const float dbot = ds * ay[k], dtop = ds - dbot;
      const float dtr = dtop * ax[k], dbr = dbot * ax[k];
      atomicAdd(g.d_pt + i00, dtop - dtr);
      atomicAdd(g.d_pt + i00 + 1, dtr);
      atomicAdd(g.d_pt + i00 + W, dbot - dbr);
      atomicAdd(g.d_pt + i00 + W + 1, dbr);
      dax[k] += dtop * (tr - tl) + dbot * (br - bl);
      day[k] += ds * (bot - top);
    }
    const float dqx = passx[k] ? dax[k] : 0.f, dqy = passy[k] ? day[k] : 0.f;
    // q = pix + (p_rot + d / div) - start, div = s / rho, rho = clip(para_l + k - r)
    const float ddiv = -(dqx * e.dx + dqy * e.dy) / (div_[k] * div_[k]);
    const float drho = -ddiv * e.s / (rho_[k] * rho_[k]);
    if (passr[k]) dpl += drho;
  }
  g.d_pl[p] = dpl;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// pscv_smem.cu: shared-memory staged kernel (gather convention, the shapes it is built for); 1 = launched, 0 = not handled
int m4d_pscv_smem_try_launch(const PscvArgs& a, int ctas_per_sm, cudaStream_t st, cudaError_t* err);

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
  const bool force_generic = (interp & M4D_INTERP_FLAG_GENERIC) != 0;
  const bool force_tile = (interp & M4D_INTERP_FLAG_TILE) != 0;
  const bool force_warp = (interp & M4D_INTERP_FLAG_WARP) != 0;
  const int ctas_override = (interp >> 12) & 0xF;                  // tuning experiments only (tools/microbench.py)
  interp &= 0xFF;
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
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaSuccess;
  const bool fast_c = c == 16 || c == 32 || c == 64 || c == 96 || c == 128 || c == 192;
  // gather convention (the default), level-2 shape: window of c2 staged in shared memory by bulk copies (pscv_smem.cu)
  if (search_range == 4 && interp == kGather && !force_generic && !force_tile && !force_warp) {
    const int hit = m4d_pscv_smem_try_launch(a, ctas_override, st, &e);
    if (hit < 0) {
      m4d_set_error("m4d_pscv_fused_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return M4D_ECUDA;
    }
    if (hit > 0) {
      M4D_CHECK_LAUNCH("m4d_pscv_fused_fwd");
      return M4D_OK;
    }
  }
  // network (c, cuts) pairs (m4depth_network.py:59,174): the warp-autonomous kernel
  if (search_range == 4 && !force_generic && !force_tile) {
    WArgs wa;
    wa.a = a;
    wa.WQ = (uint32_t)w * (uint32_t)a.Q;
    const int cps = ctas_override;
    bool hit = true;
    // measured on B200 (tools/microbench.py pscv, b=8): 175 vs 192 us at level 1, 86 vs 92 us at level 2; from c = 64 on a
    // pixel is a whole pass or more, a warp task becomes a long serial chain and the CTA-tile kernel is faster.
    switch (c * 16 + cuts) {
      case 16 * 16 + 1: e = launch_warp<16, 1, 1>(wa, interp, cps, st); break;
      case 32 * 16 + 2: e = launch_warp<32, 2, 1>(wa, interp, cps, st); break;
      default: hit = false;
    }
    if (hit) {
      if (e != cudaSuccess) {
        m4d_set_error("m4d_pscv_fused_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return M4D_ECUDA;
      }
      M4D_CHECK_LAUNCH("m4d_pscv_fused_fwd");
      return M4D_OK;
    }
  }
  if (search_range == 4 && fast_c && b <= 65535 && !force_generic) {
    FastArgs fa;
    fa.a = a;
    fa.WQ = (uint32_t)w * (uint32_t)a.Q;
    fa.inv_per_pix = 1.0f / (float)(cuts * K);
    const int gwi = c / cuts;
    fa.inv_gw = (gwi & (gwi - 1)) == 0 ? 1.0f / (float)gwi : 0.f;
#define M4D_FAST_CASE(CC)                                                \
  case CC:                                                               \
    fa.tiles_x = (w + FastCfg<CC>::TW - 1) / FastCfg<CC>::TW;            \
    fa.tiles_y = (h + FastCfg<CC>::TH - 1) / FastCfg<CC>::TH;            \
    M4D_REQUIRE(fa.tiles_y <= 65535, "m4d_pscv_fused_fwd: image too tall for the grid"); \
    e = launch_fast<CC>(fa, interp, st);                                 \
    break;
    switch (c) {
      M4D_FAST_CASE(16) M4D_FAST_CASE(32) M4D_FAST_CASE(64) M4D_FAST_CASE(96) M4D_FAST_CASE(128) M4D_FAST_CASE(192)
    }
#undef M4D_FAST_CASE
    if (e != cudaSuccess) {
      m4d_set_error("m4d_pscv_fused_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return M4D_ECUDA;
    }
    M4D_CHECK_LAUNCH("m4d_pscv_fused_fwd");
    return M4D_OK;
  }
  const size_t smem = (size_t)a.TP * K * sizeof(TapRec) + (size_t)a.TP * a.Q * K * sizeof(float) + (size_t)a.TP * sizeof(PixRec);
  M4D_REQUIRE(smem <= 200 * 1024, "m4d_pscv_fused_fwd: c=%d needs %zu bytes of shared memory", c, smem);
  const int grid = (int)cdiv64(npix, a.TP);
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

int m4d_pscv_fused_bwd(const float* c1, const float* c2, const float* para_prev_t, const float* para_prev_l,
                       const float* rot, int rot_dim, const float* trans, const float* cam_f, const float* cam_c,
                       int b, int h, int w, int c, int cuts, int search_range,
                       const float* d_cv, int d_cv_pix_stride, const float* d_prev_disp, int d_pd_pix_stride,
                       float* d_c1, float* d_c2, float* d_para_prev_t, float* d_para_prev_l, void* stream) {
  M4D_REQUIRE(c1 && c2 && para_prev_t && para_prev_l && rot && trans && cam_f && cam_c && d_cv && d_c1 && d_c2 && d_para_prev_t &&
              d_para_prev_l, "m4d_pscv_fused_bwd: null pointer");
  M4D_REQUIRE(b > 0 && h >= 2 && w >= 2 && c > 0 && cuts > 0, "m4d_pscv_fused_bwd: bad sizes (the gather convention needs h,w >= 2)");
  M4D_REQUIRE(rot_dim == 3 || rot_dim == 4, "m4d_pscv_fused_bwd: rot_dim must be 3 or 4");
  M4D_REQUIRE(search_range >= 0 && search_range <= 8, "m4d_pscv_fused_bwd: search_range must be in [0,8] (got %d)", search_range);
  M4D_REQUIRE(c % cuts == 0, "m4d_pscv_fused_bwd: c must be a multiple of cuts");
  const int K = 2 * search_range + 1;
  M4D_REQUIRE(d_cv_pix_stride >= cuts * K && (!d_prev_disp || d_pd_pix_stride >= K), "m4d_pscv_fused_bwd: gradient pixel stride too small");
  PscvBwdArgs g;
  PscvArgs& a = g.a;
  a.c1 = c1; a.c2 = c2; a.para_t = para_prev_t; a.para_l = para_prev_l; a.rot = rot; a.trans = trans; a.cam_f = cam_f; a.cam_c = cam_c;
  a.cv = nullptr; a.prev_disp = nullptr; a.centre_log = nullptr; a.idx_dbg = nullptr;
  a.rot_dim = rot_dim; a.b = b; a.h = h; a.w = w; a.c = c; a.cuts = cuts; a.r = search_range; a.K = K; a.Q = 0; a.TP = 0;
  a.cv_stride = a.pd_stride = a.cl_stride = 0; a.cl_scale = 1.f; a.npix = (int64_t)b * h * w;
  a.one = 1.0f; a.neg_one = -1.0f; a.neg_zero = -0.0f;
  g.d_cv = d_cv; g.d_pd = d_prev_disp; g.d_c1 = d_c1; g.d_c2 = d_c2; g.d_pt = d_para_prev_t; g.d_pl = d_para_prev_l;
  g.dcv_stride = d_cv_pix_stride; g.dpd_stride = d_pd_pix_stride;
  cudaStream_t st = (cudaStream_t)stream;
  cudaError_t e = cudaMemsetAsync(d_c2, 0, (size_t)a.npix * c * sizeof(float), st);
  if (e == cudaSuccess) e = cudaMemsetAsync(d_para_prev_t, 0, (size_t)a.npix * sizeof(float), st);
  if (e != cudaSuccess) {
    m4d_set_error("m4d_pscv_fused_bwd: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return M4D_ECUDA;
  }
  pscv_bwd_kernel<<<(int)cdiv64(a.npix, 128), 128, 0, st>>>(g);
  M4D_CHECK_LAUNCH("m4d_pscv_fused_bwd");
  return M4D_OK;
}

}  // extern "C"
