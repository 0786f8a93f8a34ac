// Shared pieces of the fused backproject + PSCV kernels (pscv.cu, pscv_smem.cu): packed fp32x2 arithmetic, tap records,
// the argument block.  Everything lives in an anonymous namespace: each translation unit gets its own copy.
#pragma once
#include "common.cuh"

// argument block shared by every PSCV kernel (one definition, external linkage: it crosses translation units)
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
    return __fmaf_rn(v11, t.w[3], __fmaf_rn(v10, t.w[2], __fmaf_rn(v00, t.w[0], FMUL(v01, t.w[1]))));
  }
}

__device__ __forceinline__ u64 add2(u64 a, u64 b) {
  u64 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ u64 sub2(u64 a, u64 b) {
  u64 d;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
// fp32 sum, in order, of the four fp16 values of two half2 registers: ((p0 + p1) + p2) + p3, each add rounded to fp32
__device__ __forceinline__ float sum4_h(uint32_t p01, uint32_t p23) {
  float r;
  asm("{\n\t.reg .b16 a, b, c, d;\n\t.reg .f32 t;\n\t"
      "mov.b32 {a, b}, %1;\n\tmov.b32 {c, d}, %2;\n\t"
      "cvt.f32.f16 t, a;\n\tadd.rn.f32.f16 t, b, t;\n\tadd.rn.f32.f16 t, c, t;\n\tadd.rn.f32.f16 %0, d, t;\n\t}"
      : "=f"(r) : "r"(p01), "r"(p23));
  return r;
}
__device__ __forceinline__ uint32_t h2_bits(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }

__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

}  // namespace
