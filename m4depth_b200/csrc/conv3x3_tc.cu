// libm4d: tcgen05 (5th-gen tensor core) path of the stride-1 3x3 convolution, fp32-faithful through 3xTF32.
//
// Keras Conv2D(3x3, 'same') + bias + leaky_relu on NHWC fp32 (m4depth_network.py:104-114, the DispRefiner convs that
// hold 95 % of a frame's FLOPs) as an implicit GEMM per 16x8-pixel output tile:
//
//     D[128 pixels, Cout] = sum over (k-block of 32 input channels, tap ky,kx)  A_tap[128, 32] * W_tap[32, Cout]
//
//   * operands are fp32 in HBM; each is split x = hi + lo with hi = tf32(x) (round to nearest) and the product is
//     evaluated as hi*hi + lo*hi + hi*lo on the tensor cores (kind::tf32, fp32 accumulation in TMEM): the dropped lo*lo
//     term is 2^-22 relative, i.e. the result is as close to the exact sum as an fp32 FMA chain is;
//   * A: ONE TMA load per k-block brings the (16+2)x(8+2) pixel halo x 32 channels into shared memory (SWIZZLE_128B, one
//     128-byte row per pixel; out-of-image pixels and channels >= Cin are zero-filled by the TMA unit = 'same' padding).
//     The nine taps are nine UMMA descriptors into that one halo: tile row g of tap (ky,kx) is the 8-pixel run starting
//     at halo pixel (g+ky, kx), so the descriptor's start address is (ky*10+kx)*128 B and its 8-row-group stride (SBO)
//     is one halo row = 1280 B.  Four "splitter" warps turn the landed fp32 halo into its hi / lo planes in place;
//   * B: weights are pre-split and pre-packed once per layer (m4d_conv3x3_tc_pack) as K-major [k-block][tap][hi|lo][Cout][32]
//     so that one 2-D TMA load per (k-block, tap) brings both planes;
//   * warp roles: 0 = A producer, 1 = B producer, 2 = MMA issuer (one thread), 3 = TMEM allocator, 4-7 = splitter, then
//     epilogue (tcgen05.ld -> bias + leaky_relu -> NHWC stores).  mbarrier pipelines: A full/ready/empty (2 stages),
//     B full/empty (4 stages), accumulator full.
#include "common.cuh"

#include <cuda.h>
#include <stdlib.h>

namespace {

constexpr int TILE_W = 8, TILE_H = 16;
constexpr int HALO_W = TILE_W + 2, HALO_H = TILE_H + 2;
constexpr int KC = 32;                                         // channels per k-block (128-byte rows)
constexpr int A_BYTES = HALO_W * HALO_H * KC * 4;              // 23040: one halo plane as landed by TMA
constexpr int A_SLOT = (A_BYTES + 1023) / 1024 * 1024;         // 23552
constexpr int A_STAGES = 2;
constexpr int B_STAGES = 4;
constexpr int NTHREADS = 256;

struct TcArgs {
  const float* bias;
  float* y;
  int h, w, cout, ys, kblocks;      // cout = MMA N = output channels rounded up to 16
  int cout_real;                    // channels actually stored
  float alpha;
  uint32_t tmem_cols;
  int vec_out;       // float4 stores allowed (cout_real % 4 == 0, ys % 4 == 0, y 16-byte aligned)
  int dbg;           // timing experiments only (M4D_TC_DEBUG): 1 = reuse resident weight slabs, 2 = skip halo reload + split, 4 = no stores
  int cs;            // thread-block cluster size along x (1 or 2): the CTAs of a cluster share every weight slab
};

// ---------------------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// one rank's share of a weight slab, delivered to the same shared-memory offset (and mbarrier) of every CTA in the cluster
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar), "h"(mask)
               : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start address and the byte
// stride between 8-row groups in 16-byte units, LBO = 1 (unused for swizzled K-major), version 1, layout type 2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, "
      "%19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
        "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld16(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ void sts128(uint32_t a, uint4 v) {
  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
// x = hi + lo with hi = x rounded to TF32 (10-bit mantissa, ties away); lo = x - hi is exact in fp32
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(__uint_as_float(x)));
  lo = __float_as_uint(__fsub_rn(__uint_as_float(x), __uint_as_float(hi)));
}

// ------------------------------------------------------------------------------------------------- the kernel
__global__ void __launch_bounds__(NTHREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w, TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;               // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t b_stage_bytes = 2u * (uint32_t)a.cout * 128u;                // hi + lo planes of [Cout][32]
  const uint32_t sA = smem0;                                                  // [A_STAGES][hi, lo][A_SLOT]
  const uint32_t sB = sA + A_STAGES * 2 * A_SLOT;                             // [B_STAGES][b_stage_bytes]
  const uint32_t sBar = sB + B_STAGES * b_stage_bytes;
  const uint32_t a_full = sBar, a_ready = sBar + 8 * A_STAGES, a_empty = sBar + 16 * A_STAGES;
  const uint32_t b_full = sBar + 24 * A_STAGES, b_empty = b_full + 8 * B_STAGES;
  const uint32_t acc_full = b_empty + 8 * B_STAGES;
  const uint32_t tmem_slot = acc_full + 8;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ox0 = blockIdx.x * TILE_W, oy0 = blockIdx.y * TILE_H, bi = blockIdx.z;
  const int KB = a.kblocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < A_STAGES; ++s) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_ready + 8 * s, 128);
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < B_STAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, a.cs);          // a slot is free once EVERY CTA of the cluster has consumed it
    }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(a.tmem_cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  if (a.cs > 1) cluster_sync_all();            // peers' barriers must be initialised before multicast data / commits reach them
  else __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  if (warp == 0) {
    // ===== A producer: one halo load per k-block
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
      for (int kb = 0; kb < KB; ++kb) {
        const int s = kb % A_STAGES;
        mbar_wait(a_empty + 8 * s, ((kb / A_STAGES) & 1) ^ 1);
        if ((a.dbg & 2) && kb >= A_STAGES) { mbar_arrive(a_full + 8 * s); continue; }
        mbar_expect_tx(a_full + 8 * s, A_BYTES);
        tma_load_4d(sA + s * 2 * A_SLOT, &tmap_x, a_full + 8 * s, kb * KC, ox0 - 1, oy0 - 1, bi);
      }
    }
  } else if (warp == 1) {
    // ===== B producer: both planes of one (k-block, tap) weight slab per load
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
      const int rank = a.cs > 1 ? (int)cluster_rank() : 0;
      const int rows = 2 * a.cout / a.cs;                       // this CTA's share of the [hi|lo][Cout] rows
      const uint16_t mask = (uint16_t)((1u << a.cs) - 1u);
      int it = 0;
      for (int kb = 0; kb < KB; ++kb)
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int s = it % B_STAGES;
          mbar_wait(b_empty + 8 * s, ((it / B_STAGES) & 1) ^ 1);
          if ((a.dbg & 1) && it >= B_STAGES) { mbar_arrive(b_full + 8 * s); continue; }
          mbar_expect_tx(b_full + 8 * s, b_stage_bytes);        // whole slab: own share + the peers' multicast shares
          const uint32_t dst = sB + s * b_stage_bytes + (uint32_t)(rank * rows) * 128u;
          const int row = (kb * 9 + tap) * 2 * a.cout + rank * rows;
          if (a.cs > 1) tma_load_2d_mc(dst, &tmap_w, b_full + 8 * s, 0, row, mask);
          else tma_load_2d(dst, &tmap_w, b_full + 8 * s, 0, row);
        }
    }
  } else if (warp == 2) {
    // ===== MMA issuer
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(a.cout >> 3) << 17) | ((128u >> 4) << 24);
      const uint32_t lo_off = (uint32_t)a.cout * 128u;
      // Four accumulators of Cout columns each.  The tensor core adds every K=8 partial product to the fp32 accumulator
      // with truncation, so the error grows with the number of additions made at full magnitude: the hi*hi products
      // are therefore spread over three accumulators (by tap column) and the two small cross terms, whose truncation
      // errors are 2^-11 smaller, go to a fourth; the epilogue adds the four in fp32.
      const uint32_t ncol = (uint32_t)a.cout;
      int it = 0;
      uint32_t started = 0;                                      // bit j: accumulator j has been written
      for (int kb = 0; kb < KB; ++kb) {
        const int sa = kb % A_STAGES;
        mbar_wait(a_ready + 8 * sa, (kb / A_STAGES) & 1);
        tc_fence_after();
        const uint32_t a_hi = sA + sa * 2 * A_SLOT, a_lo = a_hi + A_SLOT;
        for (int tap = 0; tap < 9; ++tap, ++it) {
          const int sb = it % B_STAGES;
          mbar_wait(b_full + 8 * sb, (it / B_STAGES) & 1);
          tc_fence_after();
          const uint32_t tap_off = (uint32_t)((tap / 3) * HALO_W + (tap % 3)) * 128u;
          const uint32_t b_hi = sB + sb * b_stage_bytes;
#pragma unroll
          for (int ks = 0; ks < KC / 8; ++ks) {
            const uint64_t da_hi = umma_desc(a_hi + tap_off + ks * 32, HALO_W * 128);
            const uint64_t da_lo = umma_desc(a_lo + tap_off + ks * 32, HALO_W * 128);
            const uint64_t db_hi = umma_desc(b_hi + ks * 32, 1024);
            const uint64_t db_lo = umma_desc(b_hi + lo_off + ks * 32, 1024);
            const uint32_t jm = (uint32_t)(tap % 3);
            tc_mma_tf32(tmem_base + jm * ncol, da_hi, db_hi, idesc, (started >> jm) & 1u);
            tc_mma_tf32(tmem_base + 3 * ncol, da_lo, db_hi, idesc, (started >> 3) & 1u);
            tc_mma_tf32(tmem_base + 3 * ncol, da_hi, db_lo, idesc, 1);
            started |= (1u << jm) | 8u;
          }
          if (a.cs > 1) tc_commit_mc(b_empty + 8 * sb, (uint16_t)((1u << a.cs) - 1u));
          else tc_commit(b_empty + 8 * sb);
        }
        tc_commit(a_empty + 8 * sa);
      }
      tc_commit(acc_full);
    }
  } else if (warp >= 4) {
    // ===== splitter: fp32 halo -> tf32 hi (in place) + lo plane, element-wise on the swizzled bytes
    const int t = threadIdx.x - 128;
    for (int kb = 0; kb < KB; ++kb) {
      const int s = kb % A_STAGES;
      mbar_wait(a_full + 8 * s, (kb / A_STAGES) & 1);
      const uint32_t hi_p = sA + s * 2 * A_SLOT, lo_p = hi_p + A_SLOT;
#pragma unroll 4
      for (int i = ((a.dbg & 2) && kb >= A_STAGES) ? A_BYTES : t; i < A_BYTES / 16; i += 128) {
        const uint4 v = lds128(hi_p + i * 16);
        uint4 hi, lo;
        split_tf32(v.x, hi.x, lo.x);
        split_tf32(v.y, hi.y, lo.y);
        split_tf32(v.z, hi.z, lo.z);
        split_tf32(v.w, hi.w, lo.w);
        sts128(hi_p + i * 16, hi);
        sts128(lo_p + i * 16, lo);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
      mbar_arrive(a_ready + 8 * s);
    }
    // ===== epilogue: accumulator rows (pixels) of this warp's TMEM lane quadrant
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int q = warp - 4;
    const int m = q * 32 + lane;
    const int oy = oy0 + m / TILE_W, ox = ox0 + m % TILE_W;
    const bool valid = oy < a.h && ox < a.w;
    float* yp = a.y + (((size_t)bi * a.h + (valid ? oy : 0)) * a.w + (valid ? ox : 0)) * a.ys;
    const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
    for (int c0 = 0; c0 < a.cout; c0 += 32) {
      uint32_t v[32];
      float sum[32];
      const int nc = a.cout - c0 >= 32 ? 32 : 16;
#pragma unroll
      for (int j = 0; j < 4; ++j) {                              // ((main0 + main1) + main2) + cross terms
        if (nc == 32) tc_ld32(trow + j * a.cout + c0, v);
        else tc_ld16(trow + j * a.cout + c0, v);
        tc_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nc) sum[i] = j == 0 ? __uint_as_float(v[i]) : sum[i] + __uint_as_float(v[i]);
      }
      if (valid && !(a.dbg & 4)) {
        if (a.vec_out) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < nc && c0 + j < a.cout_real) {
              const float4 bv = __ldg(reinterpret_cast<const float4*>(a.bias + c0 + j));
              float4 o;
              o.x = leaky(sum[j + 0] + bv.x, a.alpha);
              o.y = leaky(sum[j + 1] + bv.y, a.alpha);
              o.z = leaky(sum[j + 2] + bv.z, a.alpha);
              o.w = leaky(sum[j + 3] + bv.w, a.alpha);
              *reinterpret_cast<float4*>(yp + c0 + j) = o;
            }
          }
        } else {                                                 // cout not a multiple of 4 (the 5-channel output layer)
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < nc && c0 + j < a.cout_real) yp[c0 + j] = leaky(sum[j] + __ldg(a.bias + c0 + j), a.alpha);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(a.tmem_cols) : "memory");
  }
  if (a.cs > 1) cluster_sync_all();            // a peer's commit may still be in flight towards this CTA's barriers
}

// ------------------------------------------------------------------------------------------ weight packing
// HWIO [3,3,cin,cout] -> [kb][tap][hi|lo][cout][32]: row ((kb*9+tap)*2+hl)*cout+co, column ci - 32*kb (zero beyond cin)
__global__ void conv3x3_tc_pack_kernel(const float* __restrict__ w, int cin, int cout_real, int cout, int kblocks, float* __restrict__ out) {
  const int64_t n = (int64_t)kblocks * 9 * cout * KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % KC);
    const int co = (int)((i / KC) % cout);
    const int tap = (int)((i / ((int64_t)KC * cout)) % 9);
    const int kb = (int)(i / ((int64_t)KC * cout * 9));
    const int ci = kb * KC + c;
    const float v = (ci < cin && co < cout_real) ? w[((size_t)tap * cin + ci) * cout_real + co] : 0.f;
    uint32_t hi, lo, lo_r;
    split_tf32(__float_as_uint(v), hi, lo);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo_r) : "f"(__uint_as_float(lo)));
    const size_t row_hi = ((size_t)(kb * 9 + tap) * 2 + 0) * cout + co;
    const size_t row_lo = ((size_t)(kb * 9 + tap) * 2 + 1) * cout + co;
    out[row_hi * KC + c] = __uint_as_float(hi);
    out[row_lo * KC + c] = __uint_as_float(lo_r);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
      p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}

inline bool tc_shape_ok(int cin, int cout) { return cin >= 1 && cout >= 1 && cout <= 128; }
inline int tc_cout_pad(int cout) { return (cout + 15) / 16 * 16; }
inline int tc_kblocks(int cin) { return (cin + KC - 1) / KC; }

}  // namespace

extern "C" {

int64_t m4d_conv3x3_tc_packed_floats(int cin, int cout) {
  if (!tc_shape_ok(cin, cout)) return 0;
  return (int64_t)tc_kblocks(cin) * 9 * 2 * tc_cout_pad(cout) * KC;
}

int m4d_conv3x3_tc_pack(const float* kernel_hwio, int cin, int cout, float* packed, void* stream) {
  M4D_REQUIRE(kernel_hwio && packed, "m4d_conv3x3_tc_pack: null pointer");
  M4D_REQUIRE(tc_shape_ok(cin, cout), "m4d_conv3x3_tc_pack: unsupported shape cin=%d cout=%d (cout must be in [1,128])", cin, cout);
  const int kb = tc_kblocks(cin);
  const int cp = tc_cout_pad(cout);
  const int64_t n = (int64_t)kb * 9 * cp * KC;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  conv3x3_tc_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(kernel_hwio, cin, cout, cp, kb, packed);
  M4D_CHECK_LAUNCH("m4d_conv3x3_tc_pack");
  return M4D_OK;
}

int m4d_conv3x3_tc_fwd(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w, int cin,
                       int cout, float leaky_alpha, float* y, int y_pix_stride, void* stream) {
  M4D_REQUIRE(x && packed && bias && y, "m4d_conv3x3_tc_fwd: null pointer");
  M4D_REQUIRE(b > 0 && h > 0 && w > 0, "m4d_conv3x3_tc_fwd: non-positive size");
  if (!tc_shape_ok(cin, cout) || x_pix_stride % 4 != 0 || x_pix_stride < cin || y_pix_stride < cout ||
      (reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(packed) & 15u) || b > 65535 ||
      (h + TILE_H - 1) / TILE_H > 65535) {
    m4d_set_error("m4d_conv3x3_tc_fwd: shape / alignment outside the tcgen05 path (cin=%d cout=%d xs=%d ys=%d)", cin, cout, x_pix_stride, y_pix_stride);
    return M4D_ENOTSUP;
  }
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    m4d_set_error("m4d_conv3x3_tc_fwd: cuTensorMapEncodeTiled is not available from this driver");
    return M4D_ECUDA;
  }
  const int kb = tc_kblocks(cin);
  const int cout_real = cout;
  const bool vec_out = cout % 4 == 0 && y_pix_stride % 4 == 0 && !(reinterpret_cast<uintptr_t>(y) & 15u) && !(reinterpret_cast<uintptr_t>(bias) & 15u);
  cout = tc_cout_pad(cout);                    // from here on: the MMA N / packed row count
  const int tiles_x = (w + TILE_W - 1) / TILE_W;
  static const int cs_env = [] { const char* e = getenv("M4D_TC_CS"); return e ? atoi(e) : 0; }();
  // CTA pairs along x can share each weight slab through TMA multicast (M4D_TC_CS=2).  Measured on B200 it is ~4 % slower than
  // independent CTAs: the kernel is bound by the tensor pipe and the per-tile prologue / epilogue, not by L2 -> SM weight traffic
  // (profiles/r1c_conv_tc_full.md and DESIGN.md), so 1 is the default.
  const int cs = (cs_env == 2 && tiles_x >= 2) ? 2 : 1;
  CUtensorMap mx, mw;
  {
    const cuuint64_t dims[4] = {(cuuint64_t)cin, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)b};
    const cuuint64_t strides[3] = {(cuuint64_t)x_pix_stride * 4, (cuuint64_t)w * x_pix_stride * 4, (cuuint64_t)h * w * x_pix_stride * 4};
    const cuuint32_t box[4] = {KC, HALO_W, HALO_H, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      m4d_set_error("m4d_conv3x3_tc_fwd: cuTensorMapEncodeTiled(x) failed with %d", (int)r);
      return M4D_ECUDA;
    }
  }
  {
    const cuuint64_t dims[2] = {KC, (cuuint64_t)kb * 9 * 2 * cout};
    const cuuint64_t strides[1] = {KC * 4};
    const cuuint32_t box[2] = {KC, (cuuint32_t)(2 * cout / cs)};
    const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mw, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(packed), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      m4d_set_error("m4d_conv3x3_tc_fwd: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
      return M4D_ECUDA;
    }
  }
  TcArgs a;
  a.bias = bias; a.y = y; a.h = h; a.w = w; a.cout = cout; a.cout_real = cout_real; a.vec_out = vec_out ? 1 : 0;
  a.ys = y_pix_stride; a.kblocks = kb; a.alpha = leaky_alpha;
  a.tmem_cols = cout <= 16 ? 64 : cout <= 32 ? 128 : cout <= 64 ? 256 : 512;      // 4 accumulators of cout columns, power of two
  const size_t smem = 1024 + (size_t)A_STAGES * 2 * A_SLOT + (size_t)B_STAGES * 2 * cout * 128 + 256;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) {
      m4d_set_error("m4d_conv3x3_tc_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
      return M4D_ECUDA;
    }
    attr_set = true;
  }
  a.cs = cs;
  {
    static const int dbg = [] { const char* e = getenv("M4D_TC_DEBUG"); return e ? atoi(e) : 0; }();
    a.dbg = dbg;
    const char* e2 = getenv("M4D_TC_CS");
    (void)e2;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((tiles_x + cs - 1) / cs * cs, (h + TILE_H - 1) / TILE_H, b);     // columns past the image compute on zero fill, store nothing
  cfg.blockDim = dim3(NTHREADS, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t le = cudaLaunchKernelEx(&cfg, conv3x3_tc_kernel, mx, mw, a);
  if (le != cudaSuccess) {
    m4d_set_error("m4d_conv3x3_tc_fwd: launch failed: %s", cudaGetErrorString(le));
    (void)cudaGetLastError();
    return M4D_ECUDA;
  }
  M4D_CHECK_LAUNCH("m4d_conv3x3_tc_fwd");
  return M4D_OK;
}

}  // extern "C"

// conv3x3.cu's algo = 2: the tensor-core path needs the pre-packed weights of m4d_conv3x3_tc_pack
int m4d_conv3x3_tc(const float*, int, const float*, const float*, int, int, int, int, int, int, float, float*, int, cudaStream_t) {
  m4d_set_error("m4d_conv3x3_nhwc: algo 2 needs pre-packed weights: use m4d_conv3x3_tc_pack + m4d_conv3x3_tc_fwd");
  return M4D_ENOTSUP;
}
