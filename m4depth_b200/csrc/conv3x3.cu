// libm4d: Keras Conv2D(3x3, padding='same') + bias + leaky_relu on NHWC fp32 tensors
// (m4depth_network.py:63-72 encoder, :104-114 refiner).  fp32 in, fp32 accumulate, fp32 out.
//
// algo 1 (this file): direct convolution on the FP32 pipe with packed FFMA2.
//   CTA = 256 threads -> 8x16 output pixels x 64 output channels; thread = 4 pixels (along x) x 8 channels,
//   accumulators as 16 f32x2 pairs.  Input channels are consumed in chunks of 8: the (8s+2)x(16s+2) input halo is
//   transposed into channel-major planes in shared memory (zero filled = TF 'SAME' padding, pad_before =
//   floor(total/2): stride 2 on an even size pads bottom/right only), the 9x8x64 weight slab next to it.  Per
//   (ky, ci) a thread reads its 6 (stride 1) or 9 (stride 2) input values with 2-3 LDS and per kx 8 weights with
//   2 LDS.128, then issues 16 FFMA2 whose x operand is the broadcast form.
#include "common.cuh"

namespace {

typedef unsigned long long u64;
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

constexpr int TH = 8, TWD = 16, TN = 64, CK = 8;

struct ConvArgs {
  const float *x, *wgt, *bias;
  float* y;
  int b, h, w, cin, cout, oh, ow, xs, ys, pad_t, pad_l, tiles_x;
  float alpha;
};

template <int S>
__global__ void __launch_bounds__(256) conv3x3_ffma_kernel(ConvArgs a) {
  constexpr int IH = (TH - 1) * S + 3;
  constexpr int IW = (TWD - 1) * S + 3;
  constexpr int IWP = (IW + 3) & ~3;
  constexpr int NX = 3 * S + 3;                       // input values per thread row: 6 (s=1) or 9 (s=2)
  __shared__ __align__(16) float s_in[CK][IH][IWP];
  __shared__ __align__(16) float s_w[9][CK][TN];

  const int tid = threadIdx.x;
  const int tile_x = blockIdx.x % a.tiles_x, tile_y = blockIdx.x / a.tiles_x;
  const int n0 = blockIdx.y * TN, bi = blockIdx.z;
  const int ox0 = tile_x * TWD, oy0 = tile_y * TH;
  const int ix0 = ox0 * S - a.pad_l, iy0 = oy0 * S - a.pad_t;

  const int cg = tid & 7;                  // channel group: couts n0 + cg*4 + {0..3} and n0 + 32 + cg*4 + {0..3}
  const int pg = tid >> 3;                 // pixel group 0..31
  const int ry = pg >> 2, rx = (pg & 3) * 4;

  u64 acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = pk(0.f, 0.f);

  const float* xb = a.x + (size_t)bi * a.h * a.w * a.xs;
  const bool vec_in = (a.xs % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.x) & 15u) == 0);

  for (int c0 = 0; c0 < a.cin; c0 += CK) {
    __syncthreads();
    // ---- input halo -> channel-major planes
    if (vec_in && c0 + CK <= a.cin) {
      for (int e = tid; e < IH * IW * 2; e += 256) {
        const int p = e % (IH * IW), j = e / (IH * IW);
        const int lx = p % IW, ly = p / IW;
        const int gx = ix0 + lx, gy = iy0 + ly;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < a.w && gy >= 0 && gy < a.h)
          v = __ldg(reinterpret_cast<const float4*>(xb + ((size_t)gy * a.w + gx) * a.xs + c0) + j);
        s_in[j * 4 + 0][ly][lx] = v.x; s_in[j * 4 + 1][ly][lx] = v.y;
        s_in[j * 4 + 2][ly][lx] = v.z; s_in[j * 4 + 3][ly][lx] = v.w;
      }
    } else {
      for (int e = tid; e < IH * IW * CK; e += 256) {
        const int ci = e % CK, p = e / CK;
        const int lx = p % IW, ly = p / IW;
        const int gx = ix0 + lx, gy = iy0 + ly;
        float v = 0.f;
        if (c0 + ci < a.cin && gx >= 0 && gx < a.w && gy >= 0 && gy < a.h)
          v = __ldg(xb + ((size_t)gy * a.w + gx) * a.xs + c0 + ci);
        s_in[ci][ly][lx] = v;
      }
    }
    // ---- weight slab [tap][ci][co]
    for (int e = tid; e < 9 * CK * TN; e += 256) {
      const int co = e % TN, ci = (e / TN) % CK, t = e / (TN * CK);
      float v = 0.f;
      if (c0 + ci < a.cin && n0 + co < a.cout) v = __ldg(a.wgt + ((size_t)t * a.cin + c0 + ci) * a.cout + n0 + co);
      s_w[t][ci][co] = v;
    }
    __syncthreads();

#pragma unroll 2
    for (int ci = 0; ci < CK; ++ci) {
#pragma unroll
      for (int ky = 0; ky < 3; ++ky) {
        float xv[NX + 3];
        const float* src = &s_in[ci][ry * S + ky][rx * S];
        {
          const float4 v0 = *reinterpret_cast<const float4*>(src);
          xv[0] = v0.x; xv[1] = v0.y; xv[2] = v0.z; xv[3] = v0.w;
          if (S == 1) {
            const float2 v1 = *reinterpret_cast<const float2*>(src + 4);
            xv[4] = v1.x; xv[5] = v1.y;
          } else {
            const float4 v1 = *reinterpret_cast<const float4*>(src + 4);
            xv[4] = v1.x; xv[5] = v1.y; xv[6] = v1.z; xv[7] = v1.w;
            xv[8] = src[8];
          }
        }
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
          const float* wp = &s_w[ky * 3 + kx][ci][cg * 4];
          const float4 w0 = *reinterpret_cast<const float4*>(wp);
          const float4 w1 = *reinterpret_cast<const float4*>(wp + 32);
          const u64 wa = pk(w0.x, w0.y), wb = pk(w0.z, w0.w), wc = pk(w1.x, w1.y), wd = pk(w1.z, w1.w);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float xs_ = xv[i * S + kx];
            const u64 xx = pk(xs_, xs_);
            acc[i][0] = fma2(xx, wa, acc[i][0]);
            acc[i][1] = fma2(xx, wb, acc[i][1]);
            acc[i][2] = fma2(xx, wc, acc[i][2]);
            acc[i][3] = fma2(xx, wd, acc[i][3]);
          }
        }
      }
    }
  }

  // ---- epilogue: bias, leaky_relu, store
  const int oy = oy0 + ry;
  if (oy >= a.oh) return;
  float bv[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int co = n0 + (j < 4 ? cg * 4 + j : 32 + cg * 4 + (j - 4));
    bv[j] = co < a.cout ? __ldg(a.bias + co) : 0.f;
  }
  const bool vec_out = (a.ys % 4 == 0) && ((reinterpret_cast<uintptr_t>(a.y) & 15u) == 0) && (a.cout % 4 == 0);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ox = ox0 + rx + i;
    if (ox >= a.ow) continue;
    float v[8];
    upk(acc[i][0], v[0], v[1]); upk(acc[i][1], v[2], v[3]);
    upk(acc[i][2], v[4], v[5]); upk(acc[i][3], v[6], v[7]);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = leaky(v[j] + bv[j], a.alpha);
    float* o = a.y + (((size_t)bi * a.oh + oy) * a.ow + ox) * a.ys + n0;
    if (vec_out) {
      if (n0 + cg * 4 < a.cout) *reinterpret_cast<float4*>(o + cg * 4) = make_float4(v[0], v[1], v[2], v[3]);
      if (n0 + 32 + cg * 4 < a.cout) *reinterpret_cast<float4*>(o + 32 + cg * 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (n0 + cg * 4 + j < a.cout) o[cg * 4 + j] = v[j];
        if (n0 + 32 + cg * 4 + j < a.cout) o[32 + cg * 4 + j] = v[4 + j];
      }
    }
  }
}

// ---- thin-input layer (the RGB conv, m4depth_network.py:63 with cin = 3): one output pixel x all COUT channels per thread.
// The FFMA2 tile kernel above wastes 5/8 of its input-channel chunk and 3/4 of its channel tile on this shape; here the
// 27 x 16 weights sit in shared memory (broadcast reads), the 9 x cin inputs come from L1, and the layer becomes what it
// should be: bound by the 252 MB it writes.
template <int COUT>
__global__ void __launch_bounds__(256) conv3x3_thin_kernel(ConvArgs a) {
  __shared__ __align__(16) float s_w[9 * 4 * COUT];
  __shared__ float s_b[COUT];
  for (int e = threadIdx.x; e < 9 * a.cin * COUT; e += 256) s_w[e] = __ldg(a.wgt + e);       // HWIO is already [tap][ci][co]
  if (threadIdx.x < COUT) s_b[threadIdx.x] = __ldg(a.bias + threadIdx.x);
  __syncthreads();
  const int64_t npix = (int64_t)a.b * a.h * a.w;
  for (int64_t p = (int64_t)blockIdx.x * 256 + threadIdx.x; p < npix; p += (int64_t)gridDim.x * 256) {
    const int x = (int)(p % a.w), y = (int)((p / a.w) % a.h);
    float acc[COUT];
#pragma unroll
    for (int co = 0; co < COUT; ++co) acc[co] = 0.f;
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      const int gy = y + ky - 1;
      if (gy < 0 || gy >= a.h) continue;
#pragma unroll
      for (int kx = 0; kx < 3; ++kx) {
        const int gx = x + kx - 1;
        if (gx < 0 || gx >= a.w) continue;
        const float* xp = a.x + (p + (int64_t)(ky - 1) * a.w + (kx - 1)) * a.xs;
        for (int ci = 0; ci < a.cin; ++ci) {
          const float v = __ldg(xp + ci);
          const float4* wp = reinterpret_cast<const float4*>(s_w + ((ky * 3 + kx) * a.cin + ci) * COUT);
#pragma unroll
          for (int q = 0; q < COUT / 4; ++q) {
            const float4 w4 = wp[q];
            acc[4 * q + 0] = fmaf(v, w4.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(v, w4.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(v, w4.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(v, w4.w, acc[4 * q + 3]);
          }
        }
      }
    }
    float* o = a.y + p * a.ys;
#pragma unroll
    for (int q = 0; q < COUT / 4; ++q) {
      float4 v;
      v.x = leaky(acc[4 * q + 0] + s_b[4 * q + 0], a.alpha);
      v.y = leaky(acc[4 * q + 1] + s_b[4 * q + 1], a.alpha);
      v.z = leaky(acc[4 * q + 2] + s_b[4 * q + 2], a.alpha);
      v.w = leaky(acc[4 * q + 3] + s_b[4 * q + 3], a.alpha);
      reinterpret_cast<float4*>(o)[q] = v;
    }
  }
}

}  // namespace

// TF 'SAME' padding: out = ceil(in/s); total = max((out-1)*s + 3 - in, 0); before = total / 2
static inline void same_pad(int in, int s, int& out, int& before) {
  out = (in + s - 1) / s;
  int total = (out - 1) * s + 3 - in;
  if (total < 0) total = 0;
  before = total / 2;
}

int m4d_conv3x3_tc(const float* x, int xs, const float* wgt, const float* bias, int b, int h, int w, int cin, int cout,
                   int stride, float alpha, float* y, int ys, cudaStream_t st);   // conv3x3_tc.cu; M4D_ENOTSUP if unsupported

extern "C" int m4d_conv3x3_nhwc(const float* x, int x_pix_stride, const float* kernel_hwio, const float* bias,
                                int b, int h, int w, int cin, int cout, int stride, float leaky_alpha,
                                float* y, int y_pix_stride, int algo, void* stream) {
  M4D_REQUIRE(x && kernel_hwio && bias && y, "m4d_conv3x3_nhwc: null pointer");
  M4D_REQUIRE(b > 0 && h > 0 && w > 0 && cin > 0 && cout > 0, "m4d_conv3x3_nhwc: non-positive size");
  M4D_REQUIRE(stride == 1 || stride == 2, "m4d_conv3x3_nhwc: stride must be 1 or 2 (got %d)", stride);
  M4D_REQUIRE(x_pix_stride >= cin && y_pix_stride >= cout, "m4d_conv3x3_nhwc: pixel stride smaller than channel count");
  M4D_REQUIRE(algo >= 0 && algo <= 2, "m4d_conv3x3_nhwc: algo must be 0 (auto), 1 (FFMA2) or 2 (tcgen05)");
  M4D_REQUIRE(b <= 65535, "m4d_conv3x3_nhwc: batch too large for the grid");
  cudaStream_t st = (cudaStream_t)stream;
  if (algo == 2 || algo == 0) {
    int rc = m4d_conv3x3_tc(x, x_pix_stride, kernel_hwio, bias, b, h, w, cin, cout, stride, leaky_alpha, y, y_pix_stride, st);
    if (rc != M4D_ENOTSUP || algo == 2) return rc;
  }
  ConvArgs a;
  a.x = x; a.wgt = kernel_hwio; a.bias = bias; a.y = y;
  a.b = b; a.h = h; a.w = w; a.cin = cin; a.cout = cout; a.xs = x_pix_stride; a.ys = y_pix_stride; a.alpha = leaky_alpha;
  same_pad(h, stride, a.oh, a.pad_t);
  same_pad(w, stride, a.ow, a.pad_l);
  if (stride == 1 && cin <= 4 && cout == 16 && y_pix_stride % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15u) == 0) {
    const int64_t npix = (int64_t)b * h * w;
    const int64_t want = (npix + 255) / 256;
    const int grid = (int)(want < (int64_t)m4d_sm_count() * 16 ? want : (int64_t)m4d_sm_count() * 16);
    conv3x3_thin_kernel<16><<<grid, 256, 0, st>>>(a);
    M4D_CHECK_LAUNCH("m4d_conv3x3_nhwc");
    return M4D_OK;
  }
  a.tiles_x = (a.ow + TWD - 1) / TWD;
  const int tiles_y = (a.oh + TH - 1) / TH;
  dim3 grid(a.tiles_x * tiles_y, (cout + TN - 1) / TN, b);
  if (stride == 1) conv3x3_ffma_kernel<1><<<grid, 256, 0, st>>>(a);
  else conv3x3_ffma_kernel<2><<<grid, 256, 0, st>>>(a);
  M4D_CHECK_LAUNCH("m4d_conv3x3_nhwc");
  return M4D_OK;
}
