// libm4d: spatial neighbourhood cost volume (SNCV), replaces cost_volume (utils/depth_operations.py:283-313).
//
//   out[b,y,x,(dy*n+dx)*cuts + k] = leaky_0.1( mean_{j in group k} c1[b,y,x,j] * c2pad[b,y+dy-r,x+dx-r,j] ),  n = 2r+1
//
// The reference emits n*n*cuts separate slice -> multiply -> reduce_mean ops on NCHW transposes (about 25x the
// algorithmic traffic); here each CTA stages ONE zero-padded (TH+2r) x (TW+2r) halo tile of ONE feature group in
// shared memory and every thread produces n*n outputs for 4 horizontally adjacent pixels out of registers:
// per (dy, 4-channel chunk) it issues 4 + (4+2r) LDS.128 for 4*n*4 multiply-adds (packed FFMA2, even/odd channel
// partial sums).  Shared-memory pixel stride is gw+4 floats and every lane pair starts from a different chunk so
// that the 8 lanes of a quarter-warp hit 8 distinct 16-byte bank groups.
// HBM traffic: c read once (+halo re-reads served by L2), out written once.
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

constexpr int TW = 32;      // tile width (pixels); 8 threads x 4 pixels per row
constexpr int PX = 4;       // pixels per thread

struct SncvArgs {
  const float *c1, *c2;
  float* out;
  int b, h, w, c, cuts, gw, out_stride, tiles_x, tiles_y, TH;
};

// R = search range (compile time: the network uses 3).
template <int R>
__global__ void __launch_bounds__(128) sncv_kernel(SncvArgs a) {
  constexpr int N = 2 * R + 1;
  constexpr int HW_ = TW + 2 * R;          // halo tile width
  extern __shared__ __align__(16) float smem[];
  const int gw = a.gw, S = gw + 4, nch = gw / 4;
  const int TH = a.TH, HH = TH + 2 * R;
  float* halo = smem;                                  // [HH][HW_][S]   (c2, zero padded)
  float* ctr = smem + (size_t)HH * HW_ * S;            // [TH][TW][S]    (c1)

  const int tile = blockIdx.x;
  const int tx0 = (tile % a.tiles_x) * TW, ty0 = (tile / a.tiles_x) * TH;
  const int cut = blockIdx.y, bi = blockIdx.z;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const size_t img = (size_t)bi * a.h * a.w;
  const int ch0 = cut * gw;

  // ---- stage: halo of c2 (zero outside the image = tf.pad, :293) and the centre tile of c1
  for (int i = tid; i < HH * HW_ * nch; i += nthr) {
    const int j = i % nch, p = i / nch;
    const int hx = p % HW_, hy = p / HW_;
    const int gx = tx0 + hx - R, gy = ty0 + hy - R;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gx >= 0 && gx < a.w && gy >= 0 && gy < a.h)
      v = __ldg(reinterpret_cast<const float4*>(a.c2 + (img + (size_t)gy * a.w + gx) * a.c + ch0) + j);
    *reinterpret_cast<float4*>(halo + (size_t)p * S + j * 4) = v;
  }
  for (int i = tid; i < TH * TW * nch; i += nthr) {
    const int j = i % nch, p = i / nch;
    const int lx = p % TW, ly = p / TW;
    const int gx = tx0 + lx, gy = ty0 + ly;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gx < a.w && gy < a.h)
      v = __ldg(reinterpret_cast<const float4*>(a.c1 + (img + (size_t)gy * a.w + gx) * a.c + ch0) + j);
    *reinterpret_cast<float4*>(ctr + (size_t)p * S + j * 4) = v;
  }
  __syncthreads();

  const int g = tid & 7, row = tid >> 3;             // 8 threads per tile row
  if (row >= TH) return;
  const int lx0 = g * PX;
  const int rot = (g >> 1) & 3;                       // chunk rotation -> conflict-free LDS.128
  const float inv_gw = 1.0f / (float)gw;
  const int gy = ty0 + row;
  const bool row_ok = gy < a.h;
  const u64 Z2 = pk(0.f, 0.f);

  for (int dy = 0; dy < N; ++dy) {
    u64 acc[PX][N];
#pragma unroll
    for (int i = 0; i < PX; ++i)
#pragma unroll
      for (int d = 0; d < N; ++d) acc[i][d] = Z2;
    const float* hrow = halo + ((size_t)(row + dy) * HW_ + lx0) * S;
    const float* crow = ctr + ((size_t)row * TW + lx0) * S;
    for (int jj = 0; jj < nch; ++jj) {
      int j = jj + rot;
      if (j >= nch) j -= nch;
      u64 c_lo[PX], c_hi[PX];
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        const float4 v = *reinterpret_cast<const float4*>(crow + i * S + j * 4);
        c_lo[i] = pk(v.x, v.y);
        c_hi[i] = pk(v.z, v.w);
      }
#pragma unroll
      for (int col = 0; col < PX + 2 * R; ++col) {
        const float4 v = *reinterpret_cast<const float4*>(hrow + col * S + j * 4);
        const u64 n_lo = pk(v.x, v.y), n_hi = pk(v.z, v.w);
#pragma unroll
        for (int i = 0; i < PX; ++i) {
          const int d = col - i;               // dx index of pixel i for this neighbour column
          if (d >= 0 && d < N) {
            acc[i][d] = fma2(c_lo[i], n_lo, acc[i][d]);
            acc[i][d] = fma2(c_hi[i], n_hi, acc[i][d]);
          }
        }
      }
    }
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < PX; ++i) {
        const int gx = tx0 + lx0 + i;
        if (gx < a.w) {
          float* o = a.out + (img + (size_t)gy * a.w + gx) * a.out_stride + (size_t)(dy * N) * a.cuts + cut;
#pragma unroll
          for (int d = 0; d < N; ++d) {
            float lo, hi;
            upk(acc[i][d], lo, hi);
            const float m = (lo + hi) * inv_gw;
            o[(size_t)d * a.cuts] = leaky(m, 0.1f);          // tf.nn.leaky_relu(alpha=0.1) :311
          }
        }
      }
    }
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" int m4d_sncv_fwd(const float* c1, const float* c2, int b, int h, int w, int c, int cuts, int search_range,
                            float* out, int out_pix_stride, void* stream) {
  M4D_REQUIRE(c1 && c2 && out, "m4d_sncv_fwd: null pointer");
  M4D_REQUIRE(b > 0 && h > 0 && w > 0 && c > 0 && cuts > 0, "m4d_sncv_fwd: non-positive size");
  M4D_REQUIRE(search_range == 3, "m4d_sncv_fwd: only search_range 3 is built (m4depth_network.py:232), got %d", search_range);
  M4D_REQUIRE(c % cuts == 0 && (c / cuts) % 4 == 0, "m4d_sncv_fwd: group width c/cuts must be a multiple of 4 (c=%d cuts=%d)", c, cuts);
  M4D_REQUIRE(aligned16(c1) && aligned16(c2), "m4d_sncv_fwd: feature maps must be 16-byte aligned");
  M4D_REQUIRE(b <= 65535 && cuts <= 65535, "m4d_sncv_fwd: batch / cuts too large for the grid");
  const int n = 2 * search_range + 1;
  M4D_REQUIRE(out_pix_stride >= n * n * cuts, "m4d_sncv_fwd: out_pix_stride %d < %d", out_pix_stride, n * n * cuts);
  SncvArgs a;
  a.c1 = c1; a.c2 = c2; a.out = out; a.b = b; a.h = h; a.w = w; a.c = c; a.cuts = cuts; a.gw = c / cuts;
  a.out_stride = out_pix_stride;
  a.TH = 16;
  const int S = a.gw + 4;
  auto smem_for = [&](int th) { return (size_t)((th + 2 * search_range) * (TW + 2 * search_range) + th * TW) * S * sizeof(float); };
  while (a.TH > 4 && (smem_for(a.TH) > 100 * 1024 || a.TH / 2 >= h)) a.TH /= 2;
  const size_t smem = smem_for(a.TH);
  M4D_REQUIRE(smem <= 200 * 1024, "m4d_sncv_fwd: group width %d needs %zu bytes of shared memory", a.gw, smem);
  a.tiles_x = (w + TW - 1) / TW;
  a.tiles_y = (h + a.TH - 1) / a.TH;
  cudaStream_t st = (cudaStream_t)stream;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sncv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      m4d_set_error("m4d_sncv_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return M4D_ECUDA;
    }
    attr_set = true;
  }
  dim3 grid(a.tiles_x * a.tiles_y, cuts, b);
  sncv_kernel<3><<<grid, a.TH * 8, smem, st>>>(a);
  M4D_CHECK_LAUNCH("m4d_sncv_fwd");
  return M4D_OK;
}
