// libm4d: spatial neighbourhood cost volume (SNCV), replaces cost_volume (utils/depth_operations.py:283-313).
//
//   out[b,y,x,(dy*n+dx)*cuts + k] = leaky_0.1( mean_{j in group k} c1[b,y,x,j] * c2pad[b,y+dy-r,x+dx-r,j] ),  n = 2r+1
//
// The reference emits n*n*cuts separate slice -> multiply -> reduce_mean ops on NCHW transposes (about 25x the
// algorithmic traffic); here a CTA stages zero-padded halo tiles in shared memory, keeps the correlations in registers
// (packed FFMA2, even/odd channel partial sums) and writes each pixel's n*n*cuts outputs as one contiguous run.
// HBM traffic: c read once (+halo re-reads served by L2), out written once.
#include "common.cuh"
#include <stdlib.h>

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

// 16-byte global -> shared copy that bypasses registers; !valid writes zeros (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async16(float* dst_smem, const float* src, bool valid) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(n) : "memory");
}

struct SncvArgs {
  const float *c1, *c2;
  float* out;
  int b, h, w, c, cuts, gw, out_stride, tiles_x, tiles_y, TH, TW;
  unsigned long long m_nch, m_hw, m_tw, m_oc, m_tp;    // 2^32/d + 1: exact quotients for the small indices used here
};

// i / d for i < 2^20 or so, d >= 1, with m = 2^32/d + 1 (run-time divisors, no division instruction sequence)
__device__ __forceinline__ int fdiv(int i, unsigned long long m) { return (int)(((unsigned long long)(unsigned)i * m) >> 32); }

// R = search range (compile time: the network uses 3).
//
// CTA = one TH x TW pixel tile, ALL feature groups; thread = (pixel, dy) with the N dx-correlations of that row of the
// window in registers.  Per group the zero-padded (TH+2R) x (TW+2R) halo of c2 and the centre tile of c1 are staged in
// shared memory (pixel stride gw+4 floats: the 8 lanes of a quarter warp hit 8 distinct 16-byte bank groups); per
// 4-channel chunk a thread issues 1 + N LDS.128 and 2N FFMA2.  Results go to a shared [pixel][N*N*cuts] tile in the
// reference's channel order (offset-major, cut-minor, :309-311) and leave as whole contiguous pixel rows: the first
// version of this kernel wrote every value with its own 4-byte store at stride cuts (one sector per lane) and was bound
// by that (profiles/r1a_launches.md: 182 us at level 2 for 128 MB).
template <int R>
__global__ void __launch_bounds__(896) sncv_kernel(SncvArgs a) {
  constexpr int N = 2 * R + 1;
  extern __shared__ __align__(16) float smem[];
  const int gw = a.gw, S = gw + 4, nch = gw / 4;
  const int TH = a.TH, TW = a.TW, HH = TH + 2 * R, HW_ = TW + 2 * R, TP = TH * TW;
  const int OC = N * N * a.cuts, OS = OC | 1;          // odd pixel stride of the output tile: conflict-free column writes
  float* halo = smem;                                  // [HH][HW_][S]   (c2, zero padded)
  float* ctr = halo + (size_t)HH * HW_ * S;            // [TP][S]        (c1)
  float* otile = ctr + (size_t)TP * S;                 // [TP][OS]

  const int tx0 = (blockIdx.x % a.tiles_x) * TW, ty0 = (blockIdx.x / a.tiles_x) * TH;
  const int bi = blockIdx.z;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const size_t img = (size_t)bi * a.h * a.w;
  const int dy = fdiv(tid, a.m_tp), px = tid - dy * TP;  // blockDim = TP * N
  const int ly = fdiv(px, a.m_tw), lx = px - ly * TW;
  const float inv_gw = 1.0f / (float)gw;
  const u64 Z2 = pk(0.f, 0.f);
  const int ppr = nthr / nch;                          // pixels staged per pass
  const int sp = fdiv(tid, a.m_nch), sj = tid - sp * nch;

  for (int cut = 0; cut < a.cuts; ++cut) {
    const int ch0 = cut * gw;
    if (cut) __syncthreads();                          // everyone is done reading the previous group's tiles
    // ---- stage: halo of c2 (zero outside the image = tf.pad, :293) and the centre tile of c1.
    // nch consecutive threads per pixel (one float4 each); the (chunk, first pixel) split of tid is hoisted out of the loops
    if (tid < ppr * nch) {
      const float4* c2q = reinterpret_cast<const float4*>(a.c2 + ch0) + sj;
      for (int p = sp; p < HH * HW_; p += ppr) {
        const int hy = fdiv(p, a.m_hw), hx = p - hy * HW_;
        const int gx = tx0 + hx - R, gy = ty0 + hy - R;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx >= 0 && gx < a.w && gy >= 0 && gy < a.h) v = __ldg(c2q + (img + (size_t)gy * a.w + gx) * (a.c / 4));
        *reinterpret_cast<float4*>(halo + (size_t)p * S + sj * 4) = v;
      }
      const float4* c1q = reinterpret_cast<const float4*>(a.c1 + ch0) + sj;
      for (int p = sp; p < TP; p += ppr) {
        const int py = fdiv(p, a.m_tw);
        const int gx = tx0 + p - py * TW, gy = ty0 + py;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (gx < a.w && gy < a.h) v = __ldg(c1q + (img + (size_t)gy * a.w + gx) * (a.c / 4));
        *reinterpret_cast<float4*>(ctr + (size_t)p * S + sj * 4) = v;
      }
    }
    __syncthreads();

    u64 acc[N];
#pragma unroll
    for (int d = 0; d < N; ++d) acc[d] = Z2;
    const float* hrow = halo + ((size_t)(ly + dy) * HW_ + lx) * S;
    const float* crow = ctr + (size_t)px * S;
    for (int j = 0; j < nch; ++j) {
      const float4 cv = *reinterpret_cast<const float4*>(crow + j * 4);
      const u64 c_lo = pk(cv.x, cv.y), c_hi = pk(cv.z, cv.w);
#pragma unroll
      for (int d = 0; d < N; ++d) {
        const float4 v = *reinterpret_cast<const float4*>(hrow + d * S + j * 4);
        acc[d] = fma2(c_lo, pk(v.x, v.y), acc[d]);
        acc[d] = fma2(c_hi, pk(v.z, v.w), acc[d]);
      }
    }
    float* o = otile + (size_t)px * OS + (size_t)(dy * N) * a.cuts + cut;
#pragma unroll
    for (int d = 0; d < N; ++d) {
      float lo, hi;
      upk(acc[d], lo, hi);
      o[d * a.cuts] = leaky((lo + hi) * inv_gw, 0.1f);            // tf.nn.leaky_relu(alpha=0.1) :311
    }
  }
  __syncthreads();
  // ---- contiguous pixel rows out: one warp per pixel, lanes along the channels
  const int warp = tid >> 5, lane = tid & 31, nwarps = nthr >> 5;      // full warps only (blockDim = 7 * TP >= 112)
  for (int p = warp < nwarps ? warp : TP; p < TP; p += nwarps) {
    const int py = fdiv(p, a.m_tw);
    const int gx = tx0 + p - py * TW, gy = ty0 + py;
    if (gx >= a.w || gy >= a.h) continue;
    float* dst = a.out + (img + (size_t)gy * a.w + gx) * a.out_stride;
    const float* src = otile + (size_t)p * OS;
    for (int ch = lane; ch < OC; ch += 32) dst[ch] = src[ch];
  }
}

// Column-strip kernel for the network's (group width, cuts) pairs (16,1) (16,2) (32,2) (24,4) (32,4).  What bounded sncv_kernel above
// (profiles/r1e_sncv_full.md): with one (pixel, dy) per thread every 4-channel chunk costs 1 + 7 LDS.128 for 14 FFMA2 - the
// shared-memory pipe, not the FMA pipe, sets the pace - and a 4x8 tile stages a 4.4x larger halo.  Here
//   * CTA = 4 x 32 pixel tile (halo 10 x 38 = 3.0x), 7 warps; warp = window column dx, lane = tile column x;
//   * a thread owns the 4 pixels of its column and all 7 window rows dy: per chunk it reads 4 centre quads and the 10 halo
//     rows y-3 .. y+6 of column x+dx ONCE and feeds each into the (up to 4) accumulators it belongs to:
//     14 LDS.128 for 56 FFMA2, 2.3x fewer shared-memory bytes per FMA;
//   * lanes walk consecutive pixels at an odd 16-byte-unit pixel stride, so every LDS.128 / STS.32 is conflict-free;
//   * same FMA order per accumulator as sncv_kernel (even / odd channel pairs, chunk by chunk): bit-identical results.
template <int GW, int CUTS>
__global__ void __launch_bounds__(224, 3) sncv4_kernel(SncvArgs a) {
  constexpr int R = 3, N = 7, TH = 4, TW = 32, HH = TH + 2 * R, HW_ = TW + 2 * R, TP = TH * TW;
  constexpr int NCH = GW / 4, S = GW + 4;
  static_assert((S / 4) % 2 == 1, "odd pixel stride in 16-byte units");
  extern __shared__ __align__(16) float smem[];
  constexpr int OC = N * N * CUTS, OS = OC | 1;
  float* halo = smem;                                  // [HH][HW_][S]   (c2, zero padded)
  float* ctr = halo + HH * HW_ * S;                    // [TP][S]        (c1)
  float* otile = ctr + TP * S;                         // [TP][OS]
  const int tx0 = (blockIdx.x % a.tiles_x) * TW, ty0 = (blockIdx.x / a.tiles_x) * TH;
  const int bi = blockIdx.z;
  const int tid = threadIdx.x, lane = tid & 31, dx = tid >> 5;
  const size_t img = (size_t)bi * a.h * a.w;
  const float inv_gw = 1.0f / (float)GW;
  const u64 Z2 = pk(0.f, 0.f);
#pragma unroll 1
  for (int cut = 0; cut < CUTS; ++cut) {
    const int ch0 = cut * GW;
    if (cut) __syncthreads();                          // everyone is done reading the previous group's tiles
    // ---- stage the zero-padded halo of c2 (tf.pad, :293) and the centre tile of c1 with cp.async (zero fill outside the
    // image): nothing waits on an individual load - the first version's load -> store loop serialised ten DRAM latencies
    // per warp and set the whole kernel's pace.  A warp per halo row, lanes along (pixel, chunk).
    {
      // 10 halo rows + 4 centre rows = 14 rows, two per warp
#pragma unroll 1
      for (int row = dx; row < HH + TH; row += N) {
        const bool is_halo = row < HH;
        const int gy = is_halo ? ty0 + row - R : ty0 + row - HH;
        const bool rowin = gy >= 0 && gy < a.h;
        const int gyc = min(max(gy, 0), a.h - 1);
        const float* src_row = (is_halo ? a.c2 : a.c1) + ch0 + (img + (size_t)gyc * a.w) * a.c;
        float* dst_row = is_halo ? halo + row * HW_ * S : ctr + (row - HH) * TW * S;
        const int x_off = is_halo ? tx0 - R : tx0;
        const int n_items = (is_halo ? HW_ : TW) * NCH;
#pragma unroll 1
        for (int i = lane; i < n_items; i += 32) {
          const int hx = i / NCH, j = i - hx * NCH;
          const int gx = x_off + hx;
          const bool in = rowin && gx >= 0 && gx < a.w;
          const int gxc = min(max(gx, 0), a.w - 1);
          cp_async16(dst_row + hx * S + j * 4, src_row + (size_t)gxc * a.c + j * 4, in);
        }
      }
      asm volatile("cp.async.wait_all;" ::: "memory");
    }
    __syncthreads();

    u64 acc[TH][N];
#pragma unroll
    for (int i = 0; i < TH; ++i)
#pragma unroll
      for (int d = 0; d < N; ++d) acc[i][d] = Z2;
    const float* hcol = halo + (lane + dx) * S;        // halo column x + dx, row 0
    const float* ccol = ctr + lane * S;
#pragma unroll 1
    for (int j = 0; j < NCH; ++j) {
      u64 c_lo[TH], c_hi[TH];
#pragma unroll
      for (int i = 0; i < TH; ++i) {
        const float4 cv = *reinterpret_cast<const float4*>(ccol + i * TW * S + j * 4);
        c_lo[i] = pk(cv.x, cv.y); c_hi[i] = pk(cv.z, cv.w);
      }
#pragma unroll
      for (int rr = 0; rr < HH; ++rr) {
        const float4 v = *reinterpret_cast<const float4*>(hcol + rr * HW_ * S + j * 4);
        const u64 v_lo = pk(v.x, v.y), v_hi = pk(v.z, v.w);
#pragma unroll
        for (int i = 0; i < TH; ++i) {
          const int dy = rr - i;                       // window row of halo row rr seen from tile row i
          if (dy >= 0 && dy < N) {
            acc[i][dy] = fma2(c_lo[i], v_lo, acc[i][dy]);
            acc[i][dy] = fma2(c_hi[i], v_hi, acc[i][dy]);
          }
        }
      }
    }
#pragma unroll
    for (int i = 0; i < TH; ++i) {
      float* o = otile + (i * TW + lane) * OS + dx * CUTS + cut;
#pragma unroll
      for (int d = 0; d < N; ++d) {
        float lo, hi;
        upk(acc[i][d], lo, hi);
        o[d * N * CUTS] = leaky((lo + hi) * inv_gw, 0.1f);         // tf.nn.leaky_relu(alpha=0.1) :311
      }
    }
  }
  __syncthreads();
  // ---- out: the tile's [pixel][channel] values flattened over the lanes (every lane busy, compile-time divisor), so that
  // a warp writes whole runs of a pixel's channels
  for (int py = 0; py < TH; ++py) {
    const int gy = ty0 + py;
    if (gy >= a.h) break;
    float* drow = a.out + (img + (size_t)gy * a.w + tx0) * a.out_stride;
    const float* srow = otile + py * TW * OS;
    const int npx = min(TW, a.w - tx0);
#pragma unroll 2
    for (int idx = tid; idx < npx * OC; idx += 224) {
      const int px = idx / OC, ch = idx - px * OC;
      drow[(size_t)px * a.out_stride + ch] = srow[px * OS + ch];
    }
  }
}

template <int GW, int CUTS>
static cudaError_t launch_sncv4(SncvArgs& a, cudaStream_t st) {
  constexpr int S = GW + 4;
  const int OS = (49 * CUTS) | 1;
  const size_t smem = ((size_t)10 * 38 * S + (size_t)128 * S + (size_t)128 * OS) * sizeof(float);
  a.tiles_x = (a.w + 31) / 32;
  a.tiles_y = (a.h + 3) / 4;
  cudaError_t e = cudaFuncSetAttribute(sncv4_kernel<GW, CUTS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  dim3 grid(a.tiles_x * a.tiles_y, 1, a.b);
  sncv4_kernel<GW, CUTS><<<grid, 224, smem, st>>>(a);
  return cudaSuccess;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

extern "C" int m4d_sncv_fwd_ex(const float* c1, const float* c2, int b, int h, int w, int c, int cuts, int search_range,
                               float* out, int out_pix_stride, int variant, void* stream) {
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
  {
    const int gw = c / cuts;
    // the network's (group width, cuts) pairs below the 192-channel level (m4depth_network.py:59,174)
    // (below ~100 column-strip tiles the smaller tiles of the (pixel, dy) kernel fill the SMs better: level 5 of config 3)
    const int64_t strip_tiles = (int64_t)((w + 31) / 32) * ((h + 3) / 4) * b;
    const int key = (variant == M4D_SNCV_PIXEL_DY || (variant == M4D_SNCV_AUTO && strip_tiles < 100)) ? 0 : gw * 16 + cuts;
    if (key == 16 * 16 + 1 || key == 16 * 16 + 2 || key == 32 * 16 + 2 || key == 24 * 16 + 4 || key == 32 * 16 + 4) {
      cudaStream_t st4 = (cudaStream_t)stream;
      cudaError_t e = key == 16 * 16 + 1   ? launch_sncv4<16, 1>(a, st4)
                      : key == 16 * 16 + 2 ? launch_sncv4<16, 2>(a, st4)
                      : key == 32 * 16 + 2 ? launch_sncv4<32, 2>(a, st4)
                      : key == 24 * 16 + 4 ? launch_sncv4<24, 4>(a, st4)
                                           : launch_sncv4<32, 4>(a, st4);
      if (e != cudaSuccess) {
        m4d_set_error("m4d_sncv_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return M4D_ECUDA;
      }
      M4D_CHECK_LAUNCH("m4d_sncv_fwd");
      return M4D_OK;
    }
  }
  // small tiles (32 / 16 pixels, 7 threads per pixel): several independent CTAs per SM hide each other's staging latency
  if (cuts <= 2) { a.TH = 4; a.TW = 8; }
  else if (cuts <= 4) { a.TH = 4; a.TW = 8; }
  else { a.TH = 4; a.TW = 4; }
  {
    static const char* e = getenv("M4D_SNCV_TILE");          // tuning experiments: "TH,TW"
    int th, tw;
    if (e && sscanf(e, "%d,%d", &th, &tw) == 2 && th * tw * n <= 896 && ((size_t)th * tw * ((n * n * cuts) | 1)) * 4 < 150 * 1024) { a.TH = th; a.TW = tw; }
  }
  const int S = a.gw + 4, TP = a.TH * a.TW, OS = (n * n * cuts) | 1;
  const size_t smem = ((size_t)(a.TH + 2 * search_range) * (a.TW + 2 * search_range) * S + (size_t)TP * S + (size_t)TP * OS) * sizeof(float);
  M4D_REQUIRE(smem <= 200 * 1024, "m4d_sncv_fwd: c=%d cuts=%d needs %zu bytes of shared memory", c, cuts, smem);
  a.tiles_x = (w + a.TW - 1) / a.TW;
  a.tiles_y = (h + a.TH - 1) / a.TH;
  auto magic = [](int d) { return (0x100000000ull / (unsigned long long)d) + 1ull; };
  a.m_nch = magic(a.gw / 4); a.m_hw = magic(a.TW + 2 * search_range); a.m_tw = magic(a.TW); a.m_oc = magic(n * n * cuts); a.m_tp = magic(TP);
  cudaStream_t st = (cudaStream_t)stream;
  // function attributes are per device: remember which devices of this process have them
  static bool attr_set_dev[64] = {};
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  bool& attr_set = attr_set_dev[dev_id & 63];
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(sncv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) {
      m4d_set_error("m4d_sncv_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return M4D_ECUDA;
    }
    attr_set = true;
  }
  dim3 grid(a.tiles_x * a.tiles_y, 1, b);
  sncv_kernel<3><<<grid, TP * n, smem, st>>>(a);
  M4D_CHECK_LAUNCH("m4d_sncv_fwd");
  return M4D_OK;
}

extern "C" int m4d_sncv_fwd(const float* c1, const float* c2, int b, int h, int w, int c, int cuts, int search_range,
                            float* out, int out_pix_stride, void* stream) {
  return m4d_sncv_fwd_ex(c1, c2, b, h, w, c, cuts, search_range, out, out_pix_stride, M4D_SNCV_AUTO, stream);
}
