// Shared device/host helpers for libm4d (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>

#include "../../include/m4d.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libm4d is written for sm_100a (B200) only"
#endif

// ---------------------------------------------------------------------------------------- host side
void m4d_set_error(const char* fmt, ...);
extern std::atomic<uint64_t> g_m4d_launches;

#define M4D_REQUIRE(cond, ...)                 \
  do {                                         \
    if (!(cond)) {                             \
      m4d_set_error(__VA_ARGS__);              \
      return M4D_EINVAL;                       \
    }                                          \
  } while (0)

// Called right after a <<<>>> launch: counts it and converts launch errors into a return code.
#define M4D_CHECK_LAUNCH(name)                                                   \
  do {                                                                           \
    g_m4d_launches.fetch_add(1, std::memory_order_relaxed);                      \
    cudaError_t e__ = cudaGetLastError();                                        \
    if (e__ != cudaSuccess) {                                                    \
      m4d_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));     \
      return M4D_ECUDA;                                                          \
    }                                                                            \
  } while (0)

// SM count of the CURRENT device, cached per device id (a process may drive several GPUs; relaxed atomics: every thread
// that races on the first call stores the same value).
static inline int m4d_sm_count() {
  static std::atomic<int> cache[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  const int slot = dev & 63;
  int n = cache[slot].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;  // B200
    cache[slot].store(n, std::memory_order_relaxed);
  }
  return n;
}

static inline int64_t cdiv64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// -------------------------------------------------------------------------------------- device side
// Geometry follows SURVEY.md Appendix A / utils/depth_operations.py op for op.  Every operation is an
// explicitly rounded intrinsic so that nvcc cannot contract a*b+c into an FMA: the integer tap grids
// derived from these values must be bit-identical to the oracle's (which rounds after every TF op).
#define FMUL(a, b) __fmul_rn((a), (b))
#define FADD(a, b) __fadd_rn((a), (b))
#define FSUB(a, b) __fsub_rn((a), (b))
#define FDIV(a, b) __fdiv_rn((a), (b))
#define FSQRT(a) __fsqrt_rn((a))

struct Pose {
  float R[9];
  float fx, fy, cx, cy;
  float stx, sty, stz;  // scaled_t = t * (fx, fy, 1)   (depth_operations.py:157,185,249)
  float tx, ty, tz;
};

// get_rot_mat (utils/depth_operations.py:18-53)
__device__ __forceinline__ void rot_to_mat(const float* __restrict__ rot, int rot_dim, float* R) {
  if (rot_dim == 3) {
    float x = rot[0], y = rot[1], z = rot[2];
    R[0] = 1.f; R[1] = -z;  R[2] = y;
    R[3] = z;   R[4] = 1.f; R[5] = -x;
    R[6] = -y;  R[7] = x;   R[8] = 1.f;
    return;
  }
  float w = rot[0], x = rot[1], y = rot[2], z = rot[3];
  float tx = FMUL(2.f, x), ty = FMUL(2.f, y), tz = FMUL(2.f, z);
  float twx = FMUL(tx, w), twy = FMUL(ty, w), twz = FMUL(tz, w);
  float txx = FMUL(tx, x), txy = FMUL(ty, x), txz = FMUL(tz, x);
  float tyy = FMUL(ty, y), tyz = FMUL(tz, y), tzz = FMUL(tz, z);
  R[0] = FSUB(1.f, FADD(tyy, tzz)); R[1] = FSUB(txy, twz);            R[2] = FADD(txz, twy);
  R[3] = FADD(txy, twz);            R[4] = FSUB(1.f, FADD(txx, tzz)); R[5] = FSUB(tyz, twx);
  R[6] = FSUB(txz, twy);            R[7] = FADD(tyz, twx);            R[8] = FSUB(1.f, FADD(txx, tyy));
}

__device__ __forceinline__ void load_pose(const float* __restrict__ rot, int rot_dim, const float* __restrict__ trans,
                                          const float* __restrict__ cam_f, const float* __restrict__ cam_c, int b,
                                          Pose& P) {
  rot_to_mat(rot + (size_t)b * rot_dim, rot_dim, P.R);
  P.fx = cam_f[2 * b]; P.fy = cam_f[2 * b + 1];
  P.cx = cam_c[2 * b]; P.cy = cam_c[2 * b + 1];
  P.tx = trans[3 * b]; P.ty = trans[3 * b + 1]; P.tz = trans[3 * b + 2];
  P.stx = FMUL(P.tx, P.fx); P.sty = FMUL(P.ty, P.fy); P.stz = FMUL(P.tz, 1.f);
}

// Per-pixel epipolar terms shared by parallax2depth / depth2parallax / the PSCV
// (utils/depth_operations.py:146-163, 174-191, 239-259).
struct Epi {
  float alpha, px, py, dx, dy, s, sx, sy;
};

__device__ __forceinline__ void start_coords(const Pose& P, int x, int y, float& nx, float& ny, float& sx, float& sy) {
  float mx = FSUB(FADD((float)x, 0.5f), P.cx);   // get_coords_2d :60-64
  float my = FSUB(FADD((float)y, 0.5f), P.cy);
  nx = FDIV(mx, P.fx);
  ny = FDIV(my, P.fy);
  sx = FMUL(nx, P.fx);                            // (mesh/f)*f, NOT mesh (:256)
  sy = FMUL(ny, P.fy);
}

__device__ __forceinline__ Epi epipolar(const Pose& P, int x, int y) {
  Epi e;
  float nx, ny;
  start_coords(P, x, y, nx, ny, e.sx, e.sy);
  // r = R @ (nx, ny, 1): ((R0*nx + R1*ny) + R2*1)
  float rx = FADD(FADD(FMUL(P.R[0], nx), FMUL(P.R[1], ny)), FMUL(P.R[2], 1.f));
  float ry = FADD(FADD(FMUL(P.R[3], nx), FMUL(P.R[4], ny)), FMUL(P.R[5], 1.f));
  float rz = FADD(FADD(FMUL(P.R[6], nx), FMUL(P.R[7], ny)), FMUL(P.R[8], 1.f));
  e.alpha = rz;
  e.px = FDIV(FMUL(rx, P.fx), rz);
  e.py = FDIV(FMUL(ry, P.fy), rz);
  e.dx = FSUB(P.stx, FMUL(P.stz, e.px));
  e.dy = FSUB(P.sty, FMUL(P.stz, e.py));
  e.s = FSQRT(FADD(FMUL(e.dx, e.dx), FMUL(e.dy, e.dy)));
  return e;
}

// prev_d2para (:196-215): rotation ignored
__device__ __forceinline__ float prev_d2para_px(const Pose& P, int x, int y, float prev_d) {
  float nx, ny, sx, sy;
  start_coords(P, x, y, nx, ny, sx, sy);
  float den = FSUB(prev_d, P.tz);
  float vx = FDIV(FSUB(P.stx, FMUL(P.tz, sx)), den);
  float vy = FDIV(FSUB(P.sty, FMUL(P.tz, sy)), den);
  return FSQRT(FADD(FMUL(vx, vx), FMUL(vy, vy)));
}

__device__ __forceinline__ float parallax2depth_px(const Epi& e, const Pose& P, float para) {
  return FDIV(FSUB(FDIV(e.s, para), P.stz), e.alpha);
}
__device__ __forceinline__ float depth2parallax_px(const Epi& e, const Pose& P, float depth) {
  return FDIV(e.s, FADD(FMUL(depth, e.alpha), P.stz));
}

// One bilinear sample of the BackProject convention (backproject_op_gpu.cu.cc:44-59).
struct Tap {
  int x0, y0;      // floor
  int dxo, dyo;    // x1-x0, y1-y0 in {0,1} (ceil == floor on integral coordinates)
  float wx, wy;    // fractional parts
  bool inside;     // coordinate inside [0,W-1]x[0,H-1] and not NaN
};

__device__ __forceinline__ Tap make_tap(float qx, float qy, int W, int H) {
  Tap t;
  t.inside = (qx >= 0.f) && (qy >= 0.f) && (qx <= (float)(W - 1)) && (qy <= (float)(H - 1));
  float x = t.inside ? qx : 0.f, y = t.inside ? qy : 0.f;
  float fx0 = floorf(x), fy0 = floorf(y);
  t.x0 = (int)fx0; t.y0 = (int)fy0;
  t.dxo = (int)ceilf(x) - t.x0;
  t.dyo = (int)ceilf(y) - t.y0;
  t.wx = FSUB(x, fx0);
  t.wy = FSUB(y, fy0);
  return t;
}

// tf.clip_by_value(q, 0, size-1) with NaN propagation (dense_image_warp.py:248)
__device__ __forceinline__ float clip_keep_nan(float q, float hi) {
  return (q != q) ? q : fminf(fmaxf(q, 0.f), hi);
}

__device__ __forceinline__ void tap_weights(float wx, float wy, float& w00, float& w01, float& w10, float& w11) {
  float ox = FSUB(1.f, wx), oy = FSUB(1.f, wy);
  w00 = FMUL(oy, ox); w01 = FMUL(oy, wx); w10 = FMUL(wy, ox); w11 = FMUL(wy, wx);
}

__device__ __forceinline__ float leaky(float v, float alpha) { return v >= 0.f ? v : v * alpha; }
