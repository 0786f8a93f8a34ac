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
//   * persistent, warp-specialised CTA (one per SM): see the kernel's header comment; mbarrier rings for the halo (2 stages),
//     the weight slabs (3-6 stages) and two TMEM accumulator sets, epilogue through swizzled staging tiles and TMA stores.
#include "common.cuh"

#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int TILE_W = 8, TILE_H = 16;
constexpr int HALO_W = TILE_W + 2, HALO_H = TILE_H + 2;
constexpr int KC = 32;                                         // channels per k-block (128-byte rows)
constexpr int A_BYTES = HALO_W * HALO_H * KC * 4;              // 23040: one halo plane as landed by TMA
constexpr int A_SLOT = (A_BYTES + 1023) / 1024 * 1024;         // 23552
constexpr int MAX_A_STAGES = 3;                                // halo stages: 3 where shared memory allows (thin layers), else 2
// 3xFP16 mode: the fp32 halo lands in the first A_SLOT bytes of a stage, its two fp16 planes (64-byte rows) follow
constexpr int AH_PLANE = (HALO_W * HALO_H * KC * 2 + 1023) / 1024 * 1024;   // 12288
constexpr int A_STAGE_BYTES_F32 = 2 * A_SLOT;
constexpr int A_STAGE_BYTES_F16 = A_SLOT + 2 * AH_PLANE;
constexpr int MAX_B_STAGES = 32;
constexpr int OUT_SLOT = 128 * 128;                            // one [128 px][32 ch] fp32 staging tile
constexpr int NTHREADS = 512;

// Role timing (tools/conv_phases.py): build with -DM4D_TC_PROFILE; every role sums the clocks it spends in each of its waits
// and in its work and lane 0 writes them to a.prof[blockIdx.x * 16 + slot] at the end.
#ifdef M4D_TC_PROFILE
#define PROF_DECL(n) long long prof_##n = 0
#define PROF_BEGIN(n) const long long prof_t0_##n = clock64()
#define PROF_END(n) prof_##n += clock64() - prof_t0_##n
#define PROF_WRITE(slot, n) do { if (a.prof && lane == 0) a.prof[(size_t)blockIdx.x * 16 + (slot)] = prof_##n; } while (0)
#else
#define PROF_DECL(n)
#define PROF_BEGIN(n)
#define PROF_END(n)
#define PROF_WRITE(slot, n)
#endif

struct TcArgs {
  const float* bias;
  float* y;
  int h, w, cout, ys, kblocks;      // cout = MMA N = output channels rounded up to 16
  int cout_real;                    // channels actually stored
  int cin;                          // real input channels: k-steps of the last k-block that are all zero fill are skipped
  int tiles_x, tiles_y, ntiles;
  int na;                           // halo stages in flight (2 or 3)
  int b_resident;                   // all of the layer's weight slabs fit in the ring: loaded once per CTA, never released
  int nslices, nitems;              // output-channel slices of width cout per tile; work items = ntiles * nslices
  int ctot;                         // rows per weight plane in the packed buffer = all output channels rounded up to 16
  int nb;                           // weight-slab stages that fit in shared memory (3 at Cout = 128 ... 6)
  int nacc, nsets;                  // TMEM accumulators per tile (2 or 4) and accumulator sets (2 = epilogue overlaps the next tile)
  int tma_out;                      // epilogue through TMA stores (ys % 4 == 0 and y 16-byte aligned), else scalar stores
  int concat;                       // Cout <= 64: hi*hi and hi*lo in ONE MMA of N = 2*Cout over the adjacent [hi|lo] weight planes
  int s2d_c;                        // 0: stride 1.  C > 0: stride-2 conv over C input channels as a 2x2-cell conv (see below)
  int s2d_chunks;                   // 32-channel chunks of one input row pair's (px, c) range = 2C / 32
  int thin_mode;                    // cout <= 64: 1 = epilogue groups take alternate tiles, 4 splitter warps; 2 = one epilogue group, 8 splitter warps
  int pdl;                          // launched with programmatic stream serialization: 1 = wait for the preceding grid before
                                    // anything is read, 2 = the packed weights are older than the preceding grid: only x waits
  long long* prof;                  // M4D_TC_PROFILE builds: [grid][16] clock64 sums per role (else unused)
  const float* w_scale;             // 3xFP16 mode: 1 / (power-of-two scale the packed weights were multiplied by), device scalar
  float alpha;
};

// Stride 2 (FeaturePyramid's down-sampling convs, m4depth_network.py:66-72, even input sizes: TF SAME pads bottom / right
// only).  out[oy,ox] = sum_{ky,kx} x[2oy+ky, 2ox+kx] * w[ky,kx]  is a stride-1 conv over 2x2-pixel CELLS: input row
// 2(oy+dy)+py, column 2(ox+dx)+px with ky = 2dy+py, kx = 2dx+px, cell offsets dy,dx in {0,1}.  A 5-D tensor map over the
// NHWC tensor viewed as [b][H/2][py][W/2][(px,c)] lets ONE TMA load per k-block bring the same [18][10][32] halo as the
// stride-1 kernel: k-block kb = (py, 32-channel chunk of the contiguous (px, c) range).  Cell offset (dy,dx) is tap
// (dy+1, dx+1) of the 3x3 machinery; taps with a -1 offset, (dy=1, py=1) and (dx=1, px=1) carry no weights and are skipped
// whole (k-block, tap) slabs or k-steps at a time, so the tensor cores do exactly the 9*C MACs per output of the conv.
// (C is a multiple of 16 and a k-step is 8 channels, so the px = 0 channels of a k-block are a prefix of whole k-steps.)
// The taps of k-block kb that carry weights: ky in [ky_lo, ky_hi], kx in [kx_lo, 2]; a tap uses the first nk(kx) k-steps.
struct KbTaps {
  int ky_lo, ky_hi, kx_lo;
  int nks;      // k-steps (8 channels each) per tap; for the kx = 2 taps of the stride-2 formulation: nks2
  int nks2;
  __device__ __forceinline__ int nk(int kx) const { return kx == 2 ? nks2 : nks; }
  __device__ __forceinline__ int kx_hi() const { return nks2 > 0 ? 2 : 1; }
};
// KS = channels per MMA k-step: 8 (kind::tf32) or 16 (kind::f16)
template <bool S2D, int KS>
__device__ __forceinline__ KbTaps kb_taps(const TcArgs& a, int kb) {
  KbTaps t;
  if (!S2D) {
    const int rem = a.cin - kb * KC;
    t.nks = t.nks2 = rem >= KC ? KC / KS : (rem + KS - 1) / KS;    // 16-channel inputs: half of the k-steps are zero fill
    t.ky_lo = 0; t.ky_hi = 2; t.kx_lo = 0;
  } else {
    const int py = kb / a.s2d_chunks;
    const int ch0 = (kb - py * a.s2d_chunks) * KC;                 // first (px, c) index of this k-block; px = index / C
    const int left = a.s2d_c - ch0;                                // channels of this k-block that belong to px = 0
    t.nks = KC / KS;
    t.nks2 = left <= 0 ? 0 : (left >= KC ? KC / KS : (left + KS - 1) / KS);
    t.ky_lo = 1; t.ky_hi = py ? 1 : 2; t.kx_lo = 1;
  }
  return t;
}

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
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
// one lane of a converged warp (cute::elect_one_sync): the same lane every time, so its tcgen05.commit covers its MMAs
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xFFFFFFFF;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred));
  return pred != 0;
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
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// The same MMAs with each shared-memory descriptor passed as its two 32-bit words.  The high word (stride between 8-row
// groups, version, swizzle mode) is a constant of the kernel and only the 14-bit address field of the low word moves, so the
// stride-1 issue path keeps one running low word per operand and adds compile-time tap / k-step offsets to it: one uniform
// 32-bit add per descriptor instead of a 64-bit add in vector registers plus two R2UR moves (the issuer's instruction
// stream, not the tensor pipe, paced every layer with N <= 64: profiles/r1n_conv_role_timers.md).
template <bool HALF_, bool ACC>
__device__ __forceinline__ void tc_mma_w(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo, uint32_t b_hi, uint32_t idesc) {
  // (enable-input-d is a predicate operand: a compile-time constant here, folded by ptxas)
#define M4D_MMA_W(KIND)                                                                                              \
  asm volatile("{\n\t.reg .b64 da, db;\n\t.reg .b32 t;\n\t.reg .pred p;\n\t"                                      \
               "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\tmov.b32 t, %6;\n\tsetp.ne.b32 p, t, 0;\n\t"        \
               "tcgen05.mma.cta_group::1.kind::" KIND " [%0], da, db, %5, p;\n\t}" ::"r"(d_tmem),                      \
               "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "n"(ACC ? 1 : 0)                               \
               : "memory")
  if (HALF_) M4D_MMA_W("f16");
  else M4D_MMA_W("tf32");
#undef M4D_MMA_W
}
template <bool HALF_, bool ACC>
__device__ __forceinline__ void tc_mma_d(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc) {
#define M4D_MMA_D(KIND)                                                                                              \
  asm volatile("{\n\t.reg .b32 t;\n\t.reg .pred p;\n\tmov.b32 t, %4;\n\tsetp.ne.b32 p, t, 0;\n\t"                   \
               "tcgen05.mma.cta_group::1.kind::" KIND " [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),                      \
               "l"(adesc), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0)                                                   \
               : "memory")
  if (HALF_) M4D_MMA_D("f16");
  else M4D_MMA_D("tf32");
#undef M4D_MMA_D
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): start address and the byte
// stride between 8-row groups in 16-byte units, LBO = 1 (unused for swizzled K-major), version 1, layout type 2.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (2ull << 61);
}
// the same for SWIZZLE_64B (64-byte rows: 32 fp16 channels per pixel / per output channel), layout type 4
__device__ __forceinline__ uint64_t umma_desc64(uint32_t saddr, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46) | (4ull << 61);
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
// Two values of the 3xFP16 split: xs = x * s (s a power of two: exact), h1 = fp16(xs), h2 = fp16((xs - h1) * 2^11); both
// differences and the 2^11 are exact, so the only roundings are the two conversions.  Seven instructions per pair: packed
// FMUL2, F2FP pack, two mixed-precision FHADD (h1 - xs, fp16 operand read in place), packed FMUL2 by -2^11, F2FP pack.
__device__ __forceinline__ void split_f16_pair(uint32_t x0, uint32_t x1, float s, uint32_t& h1, uint32_t& h2) {
  unsigned long long p, q, d, e;
  float a0, a1, d0, d1, e0, e1;
  asm("mov.b64 %0, {%1, %2};" : "=l"(p) : "r"(x0), "r"(x1));
  asm("{\n\t.reg .b64 s2;\n\tmov.b64 s2, {%2, %2};\n\tmul.rn.f32x2 %0, %1, s2;\n\t}" : "=l"(q) : "l"(p), "f"(s));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(q));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h1) : "f"(a1), "f"(a0));
  asm("{\n\t.reg .b16 lo, hi;\n\tmov.b32 {lo, hi}, %2;\n\tsub.rn.f32.f16 %0, lo, %3;\n\tsub.rn.f32.f16 %1, hi, %4;\n\t}"
      : "=f"(d0), "=f"(d1) : "r"(h1), "f"(a0), "f"(a1));
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(d0), "f"(d1));
  asm("{\n\t.reg .b64 k2;\n\t.reg .f32 k;\n\tmov.f32 k, 0fC5000000;\n\tmov.b64 k2, {k, k};\n\tmul.rn.f32x2 %0, %1, k2;\n\t}" : "=l"(e) : "l"(d));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(e0), "=f"(e1) : "l"(e));
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(e1), "f"(e0));
}
// x = hi + lo with hi = x rounded to TF32 (10-bit mantissa, ties away); lo = x - hi is exact in fp32
__device__ __forceinline__ void split_tf32(uint32_t x, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(__uint_as_float(x)));
  lo = __float_as_uint(__fsub_rn(__uint_as_float(x), __uint_as_float(hi)));
}

// ------------------------------------------------------------------------------------------------- the kernel
// Persistent: one CTA per SM walks the tiles blockIdx.x, blockIdx.x + gridDim.x, ...  Warp roles (512 threads):
//   0      A producer   one TMA halo load per k-block                            ring of 2 halo stages
//   1      B producer   one TMA weight-slab load per (k-block, tap)              ring of nb weight stages
//   2      MMA issuer   warp-convergent, one elected lane issues tcgen05.mma     TMEM accumulator sets alternate per k-block
//   3      TMEM allocator
//   4-7    splitter     fp32 halo -> TF32 hi (in place) + lo plane
//   8-15   epilogue     per k-block: TMEM -> register sums; per tile: bias + leaky_relu -> swizzled staging tile -> TMA store
// All rings run across tile boundaries, so the loads and the split of tile i+1 and the whole epilogue of tile i overlap
// with the MMAs (measured before this structure: 30 % of a 128-column tile was un-overlapped prologue + epilogue,
// profiles/r1f_conv_tc_tile_phases.md).
// HALF = 3xFP16 instead of 3xTF32 (same hi*hi + hi*lo + lo*hi scheme, same 2^-22 class of error, twice the tensor-core rate
// and half the operand bytes): every operand is x * s = h1 + 2^-11 * h2 with fp16 h1 = rn(x * s), h2 = rn((x * s - h1) * 2^11)
// and a power-of-two scale s that puts the largest magnitude just below 2^15 - per layer for the weights (at pack time),
// per (pixel tile, k-block) for the activations (by the splitter warps, from the landed halo) - so fp16's narrow exponent
// range costs nothing: the epilogue warps multiply each k-block's accumulators by 1 / (s_x * s_w) (and the cross terms by
// 2^-11) while adding them into their fp32 register sums.  fp16 x fp16 products are exact in the fp32 accumulator.
// WS ("wide split", layers with at most 32 output channels): 20 warps - warps 16-19 are four more splitter warps (the split
// is the longest role of those layers) while both epilogue groups keep alternating tiles; the register budget of 640 threads
// (96 per thread) holds because such a layer's epilogue owns a single 32-column chunk.
template <bool S2D, bool CONCAT, bool HALF, bool WS = false>
__global__ void __launch_bounds__(WS ? NTHREADS + 128 : NTHREADS, 1)
conv3x3_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_w,
                  const __grid_constant__ CUtensorMap tmap_y, TcArgs a) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int KS = HALF ? 16 : 8;                                           // channels per MMA k-step
  constexpr uint32_t ROWB = HALF ? 64u : 128u;                                // bytes per operand row (32 channels)
  constexpr uint32_t A_STAGE = HALF ? A_STAGE_BYTES_F16 : A_STAGE_BYTES_F32;
  const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;               // swizzled tiles need 1024-byte alignment
  const uint32_t b_stage_bytes = 2u * (uint32_t)a.cout * ROWB;                // hi + lo planes of [Cout][32]
  const uint32_t sA = smem0;                                                  // [na][A_STAGE]
  const uint32_t sOut = sA + (uint32_t)a.na * A_STAGE;                              // [2][128 px][32 ch] staging for the TMA stores
  const uint32_t sB = sOut + 2 * OUT_SLOT;                                    // [nb][b_stage_bytes]
  const uint32_t sBar = sB + (uint32_t)a.nb * b_stage_bytes;
  const uint32_t a_full = sBar, a_ready = sBar + 8 * MAX_A_STAGES, a_empty = sBar + 16 * MAX_A_STAGES;
  const uint32_t b_full = sBar + 24 * MAX_A_STAGES, b_empty = b_full + 8 * MAX_B_STAGES;
  const uint32_t acc_full = b_empty + 8 * MAX_B_STAGES, acc_empty = acc_full + 32;   // acc_full: [set][epilogue group]
  const uint32_t tmem_slot = acc_empty + 16;
  const uint32_t s_max = tmem_slot + 16;                                      // [MAX_A_STAGES (4 slots)] uint: max |x| bits of the landed halo
  const uint32_t s_scale = s_max + 16;                                        // [8] float: 1 / (s_x * s_w) of k-block ka & 7 (8 slots: the splitters run up to 3 k-blocks ahead of the MMAs, the epilogue one behind)

  const uint32_t s_bias = sBar + 1024;                                        // [256] float: the layer's bias (zero beyond cout_real)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KB = a.kblocks, NB = a.nb, NA = a.na;
  // Programmatic dependent launch (a.pdl): the next kernel in the stream may be scheduled as soon as every CTA of this grid
  // has passed this point, i.e. on SMs this grid leaves idle or frees early; its barrier set-up, TMEM allocation and weight
  // loads then overlap this grid's tail instead of following it.  What a grid reads from its predecessor (x) - and what it
  // overwrites (y: stored only after x has been read) - is ordered by griddepcontrol.wait in the A producer below.
  if (a.pdl) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  // Up to 64 output channels ("thin") an epilogue group owns both 32-column chunks of a tile and the two groups take
  // ALTERNATE tiles: with few columns the MMAs of a tile are short, and one group's chain per tile (accumulator wait, TMEM
  // read-out, bias, staging tile, TMA store: ~2800 clocks, profiles/r1n_conv_role_timers.md) was longer than the tile's MMAs
  // and set the pace of the 32->16 / 16->5 class of layers.  Every mbarrier keeps exactly one waiting role: the issuer commits
  // a tile's accumulators to the acc_full barrier of the group that owns the tile.
  // (thin_mode 2, tuning: one epilogue group for every tile and warps 12-15 as four more splitter warps)
  const bool thin = a.cout <= 64;
  const bool split8 = !WS && thin && a.thin_mode == 2;             // (tuning) one epilogue group, warps 12-15 split
  const bool alt = thin && !split8;                                // epilogue groups alternate tiles
  const uint32_t nsplit = (WS || split8) ? 256u : 128u;
  const int tiles_per_img = a.tiles_x * a.tiles_y;

  // The bias goes to shared memory once per CTA: the epilogue used to fetch its 32 values per chunk with __ldg for every tile,
  // and with the whole L1 carved out as shared memory each fetch was an L2 round trip - ~3500 clocks per 32-channel chunk of
  // every tile (role timers), which made the epilogue the pace-setter of every layer with one or two k-blocks per tile.
  if (threadIdx.x < 256) {
    if (a.pdl == 1) asm volatile("griddepcontrol.wait;" ::: "memory");      // bias possibly written by the preceding grid
    const float bv = (int)threadIdx.x < a.cout_real ? __ldg(a.bias + threadIdx.x) : 0.f;
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_bias + 4 * threadIdx.x), "f"(bv) : "memory");
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_A_STAGES; ++s) {
      mbar_init(a_full + 8 * s, 1);
      mbar_init(a_ready + 8 * s, nsplit);
      mbar_init(a_empty + 8 * s, 1);
    }
    for (int s = 0; s < MAX_B_STAGES; ++s) {
      mbar_init(b_full + 8 * s, 1);
      mbar_init(b_empty + 8 * s, 1);
    }
    for (int s = 0; s < 4; ++s) mbar_init(acc_full + 8 * s, 1);
    for (int s = 0; s < 2; ++s) mbar_init(acc_empty + 8 * s, thin ? 4 : 8);   // one lane of each warp that drains a tile's set
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int i = 0; i < 4; ++i) asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_max + 4 * i), "r"(0u) : "memory");
  }
  if (warp == 3) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  // Roles 0-2 run warp-convergent (every lane walks the loops and the barrier waits) and one elected lane issues the
  // asynchronous instructions: inside an `if (lane == 0)` region ptxas cannot keep the descriptors in uniform registers and
  // wraps every UTCHMMA / UTMALDG in an ELECT + BRA.U.ANY loop with ~10 uniform-datapath instructions each.
  if (warp == 0) {
    // ===== A producer
    if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_x) : "memory");
    if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");      // the preceding grid's writes (this layer's input) are visible
    PROF_DECL(a_wait_empty);
#ifdef M4D_TC_PROFILE
    const long long prof_kernel_t0 = clock64();
#endif
    int ka = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      const int tile = item / a.nslices;
      const int bi = tile / tiles_per_img, r = tile - bi * tiles_per_img;
      const int oy0 = (r / a.tiles_x) * TILE_H, ox0 = (r % a.tiles_x) * TILE_W;
      for (int kb = 0; kb < KB; ++kb, ++ka) {
        const int s = ka % NA;
        { PROF_BEGIN(a_wait_empty); mbar_wait(a_empty + 8 * s, ((ka / NA) & 1) ^ 1); PROF_END(a_wait_empty); }
        if (elect_one()) {
          mbar_expect_tx(a_full + 8 * s, A_BYTES);
          if (!S2D) {
            tma_load_4d(sA + s * A_STAGE, &tmap_x, a_full + 8 * s, kb * KC, ox0 - 1, oy0 - 1, bi);
          } else {
            const int py = kb / a.s2d_chunks;
            tma_load_5d(sA + s * A_STAGE, &tmap_x, a_full + 8 * s, (kb - py * a.s2d_chunks) * KC, ox0 - 1, py, oy0 - 1, bi);
          }
        }
        __syncwarp();
      }
    }
    PROF_WRITE(0, a_wait_empty);
#ifdef M4D_TC_PROFILE
    if (a.prof && lane == 0) a.prof[(size_t)blockIdx.x * 16 + 10] = clock64() - prof_kernel_t0;   // the CTA's life in SM clocks
#endif
  } else if (warp == 1) {
    // ===== B producer: both planes of one (k-block, tap) weight slab per load
    if (elect_one()) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmap_w) : "memory");
    if (a.pdl == 1) asm volatile("griddepcontrol.wait;" ::: "memory");  // weights possibly packed by the preceding grid
    int s = 0;                                                       // ring position and phase, advanced incrementally:
    uint32_t ph = 0;                                                 // (a run-time % and / per tap cost more than the tap's MMAs)
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x) {
      if (a.b_resident && item != (int)blockIdx.x) break;           // resident weights: one pass fills the ring for every tile
      const int n0 = (item % a.nslices) * a.cout;                    // first output channel of this item's slice
      for (int kb = 0; kb < KB; ++kb) {
        const KbTaps tp = kb_taps<S2D, KS>(a, kb);
        for (int ky = tp.ky_lo; ky <= tp.ky_hi; ++ky)
          for (int kx = tp.kx_lo; kx <= tp.kx_hi(); ++kx) {
            if (!a.b_resident) mbar_wait(b_empty + 8 * s, ph ^ 1);
            if (elect_one()) {
              mbar_expect_tx(b_full + 8 * s, b_stage_bytes);
              const int row = (kb * 9 + ky * 3 + kx) * 2 * a.ctot + n0;   // the slice's rows of the hi plane; lo plane: + ctot
              tma_load_2d(sB + s * b_stage_bytes, &tmap_w, b_full + 8 * s, 0, row);
              tma_load_2d(sB + s * b_stage_bytes + (uint32_t)a.cout * ROWB, &tmap_w, b_full + 8 * s, 0, row + a.ctot);
            }
            __syncwarp();
            if (++s == NB) { s = 0; ph ^= 1u; }
          }
      }
    }
  } else if (warp == 2) {
    // ===== MMA issuer
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32, A = B = TF32, both K-major, N >> 3, M >> 4
    // (operand format field: 2 = TF32, 0 = F16)
    const uint32_t fmt = HALF ? 0u : ((2u << 7) | (2u << 10));
    const uint32_t idesc = (1u << 4) | fmt | ((uint32_t)(a.cout >> 3) << 17) | ((128u >> 4) << 24);
    const uint32_t idesc2 = (1u << 4) | fmt | ((uint32_t)(a.cout >> 2) << 17) | ((128u >> 4) << 24);   // N = 2*Cout
    const uint32_t lo_off16 = ((uint32_t)a.cout * ROWB) >> 4;
    // The tensor core adds every K=8 partial product to the fp32 accumulator with TRUNCATION, so the error is biased and
    // grows with the number of additions made at full magnitude (one accumulator per tile: 1e-5 of the scale at cin = 128,
    // and the bias survives into the depth maps).  Two measures keep the kernel in the FFMA class: every k-block starts
    // fresh accumulators that the epilogue warps add in registers (round to nearest), and the small cross terms, whose
    // truncation errors are 2^-11 smaller, have their own accumulators.
    // Accumulator set (TMEM columns from d_set):  Cout > 64:  [hi*hi | lo*hi + hi*lo]           three MMAs per k-step
    //                                             concatenated: [hi*hi | hi*lo + lo*hi]         two MMAs per k-step: A_hi against
    // the adjacent [W_hi ; W_lo] planes as ONE operand of N = 2*Cout columns (which leaves hi*lo right where the cross terms
    // are collected), then A_lo x W_hi accumulated onto it.  Below 128 columns an MMA is bound by its operand reads from
    // shared memory (4 KB of A per k-step at 128 B/clk against N/2 tensor clocks), so one A pass less per k-step is 10 %
    // (N = 96) to 30 % (N <= 64) of the layer; at N = 128 it is the same tensor time in two instructions instead of three.
    const uint32_t ncol = (uint32_t)a.cout;
    PROF_DECL(i_wait_acc); PROF_DECL(i_wait_a); PROF_DECL(i_wait_b); PROF_DECL(i_issue);
    int ka = 0, sb = 0, sa = 0;
    uint32_t bph = 0, aph = 0;
    const uint32_t ring_mask = a.b_resident ? 0u : 1u;               // resident weights: every slab's barrier completed phase 0 for good
    int it = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, ++it) {
      if (a.b_resident) sb = 0;                                      // slab = position within the tile
      const uint32_t grp_off = alt ? 8u * (uint32_t)(it & 1) : 0u;   // epilogue group that owns this tile
      const bool need_wgt_wait = !a.b_resident || item == (int)blockIdx.x;   // resident slabs: landed for good after the first tile
      for (int kb = 0; kb < KB; ++kb, ++ka) {
        // every k-block accumulates into a fresh accumulator set (ping-pong): the epilogue warps add the sets in registers
        const int set = ka & 1;
        { PROF_BEGIN(i_wait_acc); mbar_wait(acc_empty + 8 * set, ((ka >> 1) & 1) ^ 1); PROF_END(i_wait_acc); }   // epilogue has drained this set
        // orders the MMAs below after the epilogue's tcgen05.ld of this set (its fence::before_thread_sync + arrive).  Once per
        // k-block: the role timers showed ~700 clocks per fence when it also sat behind every weight-slab wait, where nothing
        // needs it (TMA writes and the splitters' fence.proxy.async'ed stores reach the tensor core through the mbarriers).
        tc_fence_after();
        const uint32_t d_set = tmem_base + (uint32_t)set * 256u;
        const uint32_t d_corr = d_set + ncol;            // cross terms: right behind hi*hi, where the concatenated MMA puts hi*lo
        { PROF_BEGIN(i_wait_a); mbar_wait(a_ready + 8 * sa, aph); PROF_END(i_wait_a); }
        // descriptors of this stage's hi / lo halo planes at tap (0,0), k-step 0; taps and k-steps add 16-byte units to the low word
        const uint64_t dA_hi = HALF ? umma_desc64(sA + sa * A_STAGE + A_SLOT, HALO_W * 64) : umma_desc(sA + sa * A_STAGE, HALO_W * 128);
        const uint64_t dA_lo = HALF ? umma_desc64(sA + sa * A_STAGE + A_SLOT + AH_PLANE, HALO_W * 64)
                                    : umma_desc(sA + sa * A_STAGE + A_SLOT, HALO_W * 128);
        const KbTaps tp = kb_taps<S2D, KS>(a, kb);
        const int kx_hi = tp.kx_hi();
        const int ntap = kx_hi - (S2D ? 1 : 0) + 1;                     // taps per window row that carry weights
        if (!S2D) {
          // Stride 1: all nine taps carry weights, so the three window rows are unrolled and every tap / k-step offset is a
          // compile-time constant added to the running low word of a descriptor (tc_mma_w).  Per MMA the elected lane then
          // executes a few uniform adds instead of ~9 vector / R2UR instructions.
          constexpr uint32_t A_HI_W = HALF ? (((HALO_W * 64u) >> 4) | (1u << 14) | (4u << 29)) : (((HALO_W * 128u) >> 4) | (1u << 14) | (2u << 29));
          constexpr uint32_t B_HI_W = HALF ? ((512u >> 4) | (1u << 14) | (4u << 29)) : ((1024u >> 4) | (1u << 14) | (2u << 29));
          const uint32_t a_plane0 = sA + sa * A_STAGE + (HALF ? (uint32_t)A_SLOT : 0u);
          const uint32_t aw_hi = ((a_plane0 >> 4) & 0x3FFFu) | (1u << 16);
          const uint32_t aw_lo = aw_hi + ((HALF ? (uint32_t)AH_PLANE : (uint32_t)A_SLOT) >> 4);
          const uint32_t bw0 = ((sB >> 4) & 0x3FFFu) | (1u << 16);
          const uint32_t bslab16 = b_stage_bytes >> 4;
          const int nk = tp.nks;
          // FULLK: every k-step of the k-block holds input channels (all but the last k-block of a layer whose channel count is
          // not a multiple of 32): straight-line code, no per-k-step test inside the elected region
          // the MMAs of one tap (ky, kx) against the weight slab in ring / resident slot `slab`
          auto tap_mmas = [&](auto fullk, int ky, int kx, int slab) {
            constexpr bool FULLK = decltype(fullk)::value;
            const uint32_t tap16 = ((uint32_t)(ky * HALO_W + kx) * ROWB) >> 4;
            const uint32_t bw = bw0 + (uint32_t)slab * bslab16;
#pragma unroll
            for (int ks = 0; ks < KC / KS; ++ks) {
              if (FULLK || ks < nk) {
                const uint32_t a_hi = aw_hi + tap16 + ks * 2, a_lo = aw_lo + tap16 + ks * 2;
                const uint32_t b_hi = bw + ks * 2, b_lo = b_hi + lo_off16;
                if (ky == 0 && kx == 0 && ks == 0) {
                  tc_mma_w<HALF, false>(d_set, a_hi, A_HI_W, b_hi, B_HI_W, CONCAT ? idesc2 : idesc);
                  if (CONCAT) tc_mma_w<HALF, true>(d_corr, a_lo, A_HI_W, b_hi, B_HI_W, idesc);     // hi*lo is already there
                  else tc_mma_w<HALF, false>(d_corr, a_lo, A_HI_W, b_hi, B_HI_W, idesc);
                } else {
                  tc_mma_w<HALF, true>(d_set, a_hi, A_HI_W, b_hi, B_HI_W, CONCAT ? idesc2 : idesc);
                  tc_mma_w<HALF, true>(d_corr, a_lo, A_HI_W, b_hi, B_HI_W, idesc);
                }
                if (!CONCAT) tc_mma_w<HALF, true>(d_corr, a_hi, A_HI_W, b_lo, B_HI_W, idesc);
              }
            }
          };
          // weights through the ring (or resident but still landing: the CTA's first tile): one elected region per window
          // row, after the waits for its three slabs
          auto issue_rows = [&](auto fullk) {
#pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
              int sbs[3];
              {
                PROF_BEGIN(i_wait_b);
                int s_ = sb;
                uint32_t ph_ = bph;
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                  sbs[j] = s_;
                  if (need_wgt_wait) mbar_wait(b_full + 8 * s_, ph_ & ring_mask);
                  if (++s_ == NB) { s_ = 0; ph_ ^= 1u; }
                }
                sb = s_;
                bph = ph_;
                PROF_END(i_wait_b);
              }
              PROF_BEGIN(i_issue);
              if (elect_one()) {
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                  tap_mmas(fullk, ky, kx, sbs[kx]);
                  if (!a.b_resident) tc_commit(b_empty + 8 * sbs[kx]);    // resident slabs are never released
                }
                if (ky == 2) {
                  tc_commit(a_empty + 8 * sa);
                  tc_commit(acc_full + 16 * set + grp_off);
                }
              }
              __syncwarp();
              PROF_END(i_issue);
            }
          };
          // resident weights, landed for good: nothing to wait for, the whole k-block is ONE elected region of straight-line
          // code (slab = k-block * 9 + tap, no ring arithmetic)
          auto issue_block = [&](auto fullk) {
            PROF_BEGIN(i_issue);
            const int sb0 = sb;
            sb += 9;
            if (elect_one()) {
#pragma unroll
              for (int t = 0; t < 9; ++t) tap_mmas(fullk, t / 3, t % 3, sb0 + t);
              tc_commit(a_empty + 8 * sa);
              tc_commit(acc_full + 16 * set + grp_off);
            }
            __syncwarp();
            PROF_END(i_issue);
          };
          if (need_wgt_wait) {
            if (nk == KC / KS) issue_rows(std::true_type());
            else issue_rows(std::false_type());
          } else {
            if (nk == KC / KS) issue_block(std::true_type());
            else issue_block(std::false_type());
          }
        } else
#pragma unroll 1
        for (int ky = tp.ky_lo; ky <= tp.ky_hi; ++ky) {
          // One elected region per window ROW (up to three taps): the per-region cost - barrier polls, election, moving
          // descriptors into uniform registers - was ~70 instructions per tap and, at ~7 clocks each, longer than the six
          // fp16 MMAs of a 128-column tap (measured: the issuer, not the tensor pipe, set the pace of every layer).
          int sbs[3];
          {
            PROF_BEGIN(i_wait_b);
            int s_ = sb;
            uint32_t ph_ = bph;
#pragma unroll
            for (int j = 0; j < 3; ++j)
              if (j < ntap) {
                sbs[j] = s_;
                if (need_wgt_wait) mbar_wait(b_full + 8 * s_, ph_ & ring_mask);
                if (++s_ == NB) { s_ = 0; ph_ ^= 1u; }
              }
            sb = s_;
            bph = ph_;
            PROF_END(i_wait_b);
          }
          PROF_BEGIN(i_issue);
          if (elect_one()) {
#pragma unroll
            for (int j = 0; j < 3; ++j) {
              if (j >= ntap) break;
              const int kx = (S2D ? 1 : 0) + j;
              const int nk = tp.nk(kx);
              const uint64_t dB_hi = HALF ? umma_desc64(sB + sbs[j] * b_stage_bytes, 512) : umma_desc(sB + sbs[j] * b_stage_bytes, 1024);
              const uint32_t tap16 = ((uint32_t)(ky * HALO_W + kx) * ROWB) >> 4;
              const uint32_t later = (ky > tp.ky_lo || j > 0) ? 1u : 0u;     // past the first tap of the k-block
#pragma unroll
              for (int ks = 0; ks < KC / KS; ++ks) {              // a k-step is 32 bytes of a row in either format
                if (ks >= nk) break;
                const uint64_t a_hi = dA_hi + tap16 + ks * 2, a_lo = dA_lo + tap16 + ks * 2;
                const uint64_t b_hi = dB_hi + ks * 2, b_lo = dB_hi + lo_off16 + ks * 2;
                const uint32_t acc = ks > 0 ? 1u : later;
                if (HALF) {
                  if (CONCAT) {
                    tc_mma_f16(d_set, a_hi, b_hi, idesc2, acc);
                    tc_mma_f16(d_corr, a_lo, b_hi, idesc, 1u);
                  } else {
                    tc_mma_f16(d_set, a_hi, b_hi, idesc, acc);
                    tc_mma_f16(d_corr, a_lo, b_hi, idesc, acc);
                    tc_mma_f16(d_corr, a_hi, b_lo, idesc, 1u);
                  }
                } else if (CONCAT) {
                  tc_mma_tf32(d_set, a_hi, b_hi, idesc2, acc);
                  tc_mma_tf32(d_corr, a_lo, b_hi, idesc, 1u);
                } else {
                  tc_mma_tf32(d_set, a_hi, b_hi, idesc, acc);
                  tc_mma_tf32(d_corr, a_lo, b_hi, idesc, acc);
                  tc_mma_tf32(d_corr, a_hi, b_lo, idesc, 1u);
                }
              }
              if (!a.b_resident) tc_commit(b_empty + 8 * sbs[j]);   // resident slabs are never released (and an arrival nobody
                                                                    // waits for is what synccheck reports as a missing wait)
            }
            if (ky == tp.ky_hi) {
              tc_commit(a_empty + 8 * sa);
              tc_commit(acc_full + 16 * set + grp_off);
            }
          }
          __syncwarp();
          PROF_END(i_issue);
        }
        if (++sa == NA) { sa = 0; aph ^= 1u; }
      }
    }
    PROF_WRITE(1, i_wait_acc); PROF_WRITE(2, i_wait_a); PROF_WRITE(3, i_wait_b); PROF_WRITE(4, i_issue);
  } else if ((warp >= 4 && warp < 8) || (WS ? warp >= 16 : (split8 && warp >= 12))) {
    // ===== splitter: fp32 halo -> tf32 hi (in place) + lo plane, element-wise on the swizzled bytes / -> scaled fp16 planes
    const int t = warp < 8 ? threadIdx.x - 128 : (WS ? threadIdx.x - 512 + 128 : threadIdx.x - 384 + 128);
    const int NS = (int)nsplit;
    PROF_DECL(s_wait_full); PROF_DECL(s_split); PROF_DECL(s_bar); PROF_DECL(s_fence); PROF_DECL(s_load); PROF_DECL(s_cvt);
    // (read once: a global load per k-block in thread 0's path delayed the whole split by its latency)
    if (a.pdl == 1) asm volatile("griddepcontrol.wait;" ::: "memory");
    const float w_inv_scale = HALF ? __ldg(a.w_scale) : 1.f;
    int ka = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x)
      for (int kb = 0; kb < KB; ++kb, ++ka) {
        const int s = ka % NA;
        { PROF_BEGIN(s_wait_full); mbar_wait(a_full + 8 * s, (ka / NA) & 1); PROF_END(s_wait_full); }
        PROF_BEGIN(s_split);
        const uint32_t hi_p = sA + s * A_STAGE, lo_p = hi_p + A_SLOT;
        if (!HALF) {
#pragma unroll 4
          for (int i = t; i < A_BYTES / 16; i += NS) {
            const uint4 v = lds128(hi_p + i * 16);
            uint4 hi, lo;
            split_tf32(v.x, hi.x, lo.x);
            split_tf32(v.y, hi.y, lo.y);
            split_tf32(v.z, hi.z, lo.z);
            split_tf32(v.w, hi.w, lo.w);
            sts128(hi_p + i * 16, hi);
            sts128(lo_p + i * 16, lo);
          }
        } else {
          // a unit = 8 channels of one halo pixel: two 16-byte chunks of the landed fp32 row (SWIZZLE_128B: chunk c of
          // row r lives at c ^ (r & 7)) -> one 16-byte chunk of each fp16 plane (SWIZZLE_64B: chunk j of row r at
          // j ^ ((r >> 1) & 3); both follow from the absolute shared-memory address, the planes are 1024-byte aligned)
          constexpr int NU = HALO_W * HALO_H * 4, UPT = (NU + 127) / 128;
          // 8-channel units that hold input channels (the rest of a partial last k-block is TMA zero fill no MMA reads)
          const int rem_ch = (S2D ? 4 * a.s2d_c : a.cin) - kb * KC;
          const int jmax = rem_ch >= KC ? 4 : (rem_ch + 15) / 16 * 2;
          uint4 va[UPT], vb[UPT];
          float mxf = 0.f;
#pragma unroll
          for (int u = 0; u < UPT; ++u) {
            const int idx = t + u * NS;
            if (idx < NU && (idx & 3) < jmax) {
              const int px = idx >> 2, j = idx & 3;
              const uint32_t row = hi_p + (uint32_t)px * 128u;
              va[u] = lds128(row + (uint32_t)(((2 * j) ^ (px & 7)) * 16));
              vb[u] = lds128(row + (uint32_t)(((2 * j + 1) ^ (px & 7)) * 16));
              // |x| is an operand modifier of FMNMX / FMNMX3: five instructions per eight values (NaNs are passed over)
              const float m0 = fmaxf(fmaxf(fabsf(__uint_as_float(va[u].x)), fabsf(__uint_as_float(va[u].y))),
                                     fmaxf(fabsf(__uint_as_float(va[u].z)), fabsf(__uint_as_float(va[u].w))));
              const float m1 = fmaxf(fmaxf(fabsf(__uint_as_float(vb[u].x)), fabsf(__uint_as_float(vb[u].y))),
                                     fmaxf(fabsf(__uint_as_float(vb[u].z)), fabsf(__uint_as_float(vb[u].w))));
              mxf = fmaxf(mxf, fmaxf(m0, m1));
            }
          }
          // largest magnitude of the k-block's halo (from here on as the bits of |x|: monotonic for non-negative floats)
          uint32_t mx = __float_as_uint(mxf);
#ifdef M4D_TC_PROFILE
          prof_s_load += clock64() - prof_t0_s_split;          // loads + per-thread maximum
#endif
          mx = __reduce_max_sync(0xFFFFFFFFu, mx);           // one REDUX instead of five dependent shuffles
          // four slots in rotation: slot (ka & 3) collects this k-block's maximum, slot (ka + 2) & 3 - read two k-blocks ago, not
          // needed before two k-blocks from now - is cleared, so one barrier per k-block suffices
          const uint32_t slot = s_max + 4 * (uint32_t)(ka & 3);
          if (lane == 0) asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(slot), "r"(mx) : "memory");
          { PROF_BEGIN(s_bar); asm volatile("bar.sync 3, %0;" ::"r"(nsplit) : "memory"); PROF_END(s_bar); }
          uint32_t mbits;
          asm volatile("ld.shared.u32 %0, [%1];" : "=r"(mbits) : "r"(slot) : "memory");
          // s_x = 2^(14 - E) with E the exponent of the maximum (so max * s_x is in [2^14, 2^15)); exponents are clamped so that
          // s_x and 1 / s_x are normal floats (all-zero / denormal tiles: any scale gives zeros)
          int E = (int)(mbits >> 23) - 127;
          E = E < -100 ? -100 : (E > 100 ? 100 : E);
          const float sx = __uint_as_float((uint32_t)(127 + 14 - E) << 23);
          if (t == 0) {
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(s_max + 4 * (uint32_t)((ka + 2) & 3)), "r"(0u) : "memory");
            const float inv = __uint_as_float((uint32_t)(127 - 14 + E) << 23) * w_inv_scale;
            asm volatile("st.shared.f32 [%0], %1;" ::"r"(s_scale + 4 * (ka & 7)), "f"(inv) : "memory");
          }
          const uint32_t h1_p = hi_p + A_SLOT, h2_p = h1_p + AH_PLANE;
          PROF_BEGIN(s_cvt);
#pragma unroll
          for (int u = 0; u < UPT; ++u) {
            const int idx = t + u * NS;
            if (idx < NU && (idx & 3) < jmax) {
              const int px = idx >> 2, j = idx & 3;
              const uint32_t xin[8] = {va[u].x, va[u].y, va[u].z, va[u].w, vb[u].x, vb[u].y, vb[u].z, vb[u].w};
              uint32_t h1[4], h2[4];
#pragma unroll
              for (int i = 0; i < 4; ++i) split_f16_pair(xin[2 * i], xin[2 * i + 1], sx, h1[i], h2[i]);
              const uint32_t off = (uint32_t)px * 64u + (uint32_t)((j ^ ((px >> 1) & 3)) * 16);
              sts128(h1_p + off, make_uint4(h1[0], h1[1], h1[2], h1[3]));
              sts128(h2_p + off, make_uint4(h2[0], h2[1], h2[2], h2[3]));
            }
          }
          PROF_END(s_cvt);
        }
        { PROF_BEGIN(s_fence);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // generic-proxy writes -> visible to the tensor core
        mbar_arrive(a_ready + 8 * s);
        PROF_END(s_fence); }
        PROF_END(s_split);
      }
    if (warp == 4) { PROF_WRITE(5, s_wait_full); PROF_WRITE(6, s_split); PROF_WRITE(11, s_bar); PROF_WRITE(12, s_fence); PROF_WRITE(14, s_load); PROF_WRITE(15, s_cvt); }
  } else if (warp >= 8 && warp < 16 && (!split8 || warp < 12)) {
    // ===== epilogue: two groups of four warps; a warp owns the accumulator rows (pixels) of its TMEM lane quadrant.  More than 64
    // output channels: both groups work on every tile, group g on the 32-column chunks g and g+2.  Thin layers: group g owns
    // the CTA's tiles with local index & 1 == g, both chunks.  After every k-block the group adds that k-block's accumulators
    // into its register sums and hands the set back; after the last one it applies bias + leaky_relu and stores through its
    // staging tile.
    const int grp = (warp - 8) >> 2, q = warp & 3;
    const int m = q * 32 + lane;
    const int et = threadIdx.x - 256 - grp * 128;                  // 0..127 within the group
    // staging tiles for the TMA stores: one per group
    const int nchunks = (a.cout + 31) >> 5;
    PROF_DECL(e_wait_full); PROF_DECL(e_drain); PROF_DECL(e_final); PROF_DECL(e_store_wait);
    int ka = 0, it = 0;
    uint32_t ph_full0 = 0u, ph_full1 = 0u;                         // phase of this group's acc_full barrier of either set
    const uint32_t my_full = acc_full + (alt ? 8u * (uint32_t)grp : 0u);
    uint32_t nstore = 0;
    for (int item = blockIdx.x; item < a.nitems; item += gridDim.x, ++it) {
      if (alt && (it & 1) != grp) { ka += KB; continue; }           // the other group's tile
      const int tile = item / a.nslices;
      const int n0 = (item - tile * a.nslices) * a.cout;
      const int bi = tile / tiles_per_img, r = tile - bi * tiles_per_img;
      const int oy0 = (r / a.tiles_x) * TILE_H, ox0 = (r % a.tiles_x) * TILE_W;
      constexpr int NCI = WS ? 1 : 2;                                // 32-column chunks an epilogue thread owns
      float sum[NCI][32];
      for (int kb = 0; kb < KB; ++kb, ++ka) {
        const int set = ka & 1;
        { PROF_BEGIN(e_wait_full); mbar_wait(my_full + 16 * set, set ? ph_full1 : ph_full0); PROF_END(e_wait_full); }
        if (set) ph_full1 ^= 1u; else ph_full0 ^= 1u;
        PROF_BEGIN(e_drain);
        tc_fence_after();
        const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)set * 256u;
        float inv_main = 1.f;                                       // 3xFP16: 1 / (s_x * s_w) of this k-block, cross terms * 2^-11
        if (HALF) asm volatile("ld.shared.f32 %0, [%1];" : "=f"(inv_main) : "r"(s_scale + 4 * (ka & 7)) : "memory");
        const float inv_cross = inv_main * (1.0f / 2048.0f);
#pragma unroll
        for (int ci = 0; ci < NCI; ++ci) {
          const int c0 = (thin ? ci : grp + 2 * ci) * 32;
          if (c0 < a.cout) {
            const int nc = a.cout - c0 >= 32 ? 32 : 16;
            for (int jj = 0; jj < a.nacc; ++jj) {                   // hi*hi, then the cross-term accumulator(s)
              uint32_t v[32];
              if (nc == 32) tc_ld32(trow + jj * a.cout + c0, v);
              else tc_ld16(trow + jj * a.cout + c0, v);
              tc_ld_wait();
              const float sc = jj == 0 ? inv_main : inv_cross;
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (i < nc) {
                  if (HALF) sum[ci][i] = (kb == 0 && jj == 0) ? __uint_as_float(v[i]) * sc : fmaf(__uint_as_float(v[i]), sc, sum[ci][i]);
                  else sum[ci][i] = (kb == 0 && jj == 0) ? __uint_as_float(v[i]) : sum[ci][i] + __uint_as_float(v[i]);
                }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + 8 * set);            // every warp that works on this tile arrives, with or without chunks
        PROF_END(e_drain);
      }
      PROF_BEGIN(e_final);
      const int oy = oy0 + m / TILE_W, ox = ox0 + m % TILE_W;
      const bool valid = oy < a.h && ox < a.w;
      float* yp = a.y + (((size_t)bi * a.h + (valid ? oy : 0)) * a.w + (valid ? ox : 0)) * a.ys;
#pragma unroll
      for (int ci = 0; ci < NCI; ++ci) {
        const int c0 = (thin ? ci : grp + 2 * ci) * 32;
        if ((thin ? ci : grp + 2 * ci) >= nchunks) continue;
        const int nc = a.cout - c0 >= 32 ? 32 : 16;
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (i < nc) {
            float bv;
            asm volatile("ld.shared.f32 %0, [%1];" : "=f"(bv) : "r"(s_bias + 4u * (uint32_t)(n0 + c0 + i)));
            sum[ci][i] = leaky(sum[ci][i] + bv, a.alpha);
          }
        if (a.tma_out) {
          // staging tile [128 px][32 ch] in the SWIZZLE_128B layout the store's tensor map expects: 16-byte chunk c of
          // row m lives at chunk c ^ (m & 7).  One buffer per group: the issuing thread first waits until the previous
          // store has finished READING it.
          // one staging tile per group; a single group (split8) alternates between both and only waits for the store before
          // the previous one to have read its tile
          const uint32_t sbuf = sOut + (split8 ? (nstore & 1u) : (uint32_t)grp) * OUT_SLOT;
          ++nstore;
          { PROF_BEGIN(e_store_wait);
          if (et == 0) {
            if (split8) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          PROF_END(e_store_wait); }
          const uint32_t rowp = sbuf + (uint32_t)m * 128u;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            if (c * 4 < nc)
              sts128(rowp + (uint32_t)((c ^ (m & 7)) * 16),
                     make_uint4(__float_as_uint(sum[ci][4 * c]), __float_as_uint(sum[ci][4 * c + 1]), __float_as_uint(sum[ci][4 * c + 2]),
                                __float_as_uint(sum[ci][4 * c + 3])));
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
          asm volatile("bar.sync %0, 128;" ::"r"(1 + grp) : "memory");
          if (et == 0) {
            asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(&tmap_y), "r"(sbuf),
                         "r"(n0 + c0), "r"(ox0), "r"(oy0), "r"(bi)
                         : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
          }
        } else if (valid) {                                         // pixel stride not a multiple of 16 bytes (the 5-channel output layer)
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (i < nc && n0 + c0 + i < a.cout_real) yp[n0 + c0 + i] = sum[ci][i];
        }
      }
      PROF_END(e_final);
    }
    if (a.tma_out && et == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");   // shared memory must outlive the last store's reads
    if (warp == 8) { PROF_WRITE(7, e_wait_full); PROF_WRITE(8, e_drain); PROF_WRITE(9, e_final); PROF_WRITE(13, e_store_wait); }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 3) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ weight packing
// HWIO [3,3,cin,cout_total], output channels [co_off, co_off + cout_real) -> [kb][tap][hi|lo][cout][32]:
// row ((kb*9+tap)*2+hl)*cout+co, column = channel within the k-block (zero beyond cin / beyond cout_real).
// s2d_c > 0 (stride 2): k-block kb = (py, chunk), column -> (px, c); tap (dy+1, dx+1) holds w[2dy+py][2dx+px][c] (see kb_taps).
__global__ void conv3x3_tc_pack_kernel(const float* __restrict__ w, int cin, int cout_total, int co_off, int cout_real, int cout,
                                       int kblocks, int s2d_c, int s2d_chunks, float* __restrict__ out) {
  const int64_t n = (int64_t)kblocks * 9 * cout * KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % KC);
    const int co = (int)((i / KC) % cout);
    const int tap = (int)((i / ((int64_t)KC * cout)) % 9);
    const int kb = (int)(i / ((int64_t)KC * cout * 9));
    float v = 0.f;
    if (co < cout_real) {
      if (s2d_c == 0) {
        const int ci = kb * KC + c;
        if (ci < cin) v = w[((size_t)tap * cin + ci) * cout_total + co_off + co];
      } else {
        const int py = kb / s2d_chunks, idx = (kb - py * s2d_chunks) * KC + c;     // idx in the (px, c) range [0, 2C)
        const int px = idx / s2d_c, ci = idx - px * s2d_c;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int ky = 2 * dy + py, kx = 2 * dx + px;
        if (dy >= 0 && dx >= 0 && ky <= 2 && kx <= 2 && px < 2) v = w[((size_t)(ky * 3 + kx) * cin + ci) * cout_total + co_off + co];
      }
    }
    uint32_t hi, lo, lo_r;
    split_tf32(__float_as_uint(v), hi, lo);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(lo_r) : "f"(__uint_as_float(lo)));
    const size_t row_hi = ((size_t)(kb * 9 + tap) * 2 + 0) * cout + co;
    const size_t row_lo = ((size_t)(kb * 9 + tap) * 2 + 1) * cout + co;
    out[row_hi * KC + c] = __uint_as_float(hi);
    out[row_lo * KC + c] = __uint_as_float(lo_r);
  }
}

// 3xFP16 packing: same [kb][tap][h1|h2][cout][32] order with fp16 elements (64-byte rows), w * s_w = h1 + 2^-11 * h2, the
// power-of-two s_w putting the layer's largest |w| in [2^14, 2^15); 1 / s_w goes to scale_out for the forward kernel.
__global__ void conv3x3_tc_wmax_kernel(const float* __restrict__ w, int64_t n, unsigned int* __restrict__ out_bits) {
  unsigned int m = 0u;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    m = max(m, __float_as_uint(w[i]) & 0x7FFFFFFFu);
  for (int o = 16; o >= 1; o >>= 1) m = max(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, m);
}

__global__ void conv3x3_tc_pack_f16_kernel(const float* __restrict__ w, int cin, int cout_real, int cout, int kblocks, int s2d_c,
                                           int s2d_chunks, const unsigned int* __restrict__ wmax_bits, __half* __restrict__ out,
                                           float* __restrict__ scale_out) {
  int E = (int)(*wmax_bits >> 23) - 127;
  E = E < -100 ? -100 : (E > 100 ? 100 : E);
  const float sw = __uint_as_float((uint32_t)(127 + 14 - E) << 23);
  if (blockIdx.x == 0 && threadIdx.x == 0) *scale_out = __uint_as_float((uint32_t)(127 - 14 + E) << 23);
  const int64_t n = (int64_t)kblocks * 9 * cout * KC;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % KC);
    const int co = (int)((i / KC) % cout);
    const int tap = (int)((i / ((int64_t)KC * cout)) % 9);
    const int kb = (int)(i / ((int64_t)KC * cout * 9));
    float v = 0.f;
    if (co < cout_real) {
      if (s2d_c == 0) {
        const int ci = kb * KC + c;
        if (ci < cin) v = w[((size_t)tap * cin + ci) * cout_real + co];
      } else {
        const int py = kb / s2d_chunks, idx = (kb - py * s2d_chunks) * KC + c;
        const int px = idx / s2d_c, ci = idx - px * s2d_c;
        const int dy = tap / 3 - 1, dx = tap % 3 - 1;
        const int ky = 2 * dy + py, kx = 2 * dx + px;
        if (dy >= 0 && dx >= 0 && ky <= 2 && kx <= 2 && px < 2) v = w[((size_t)(ky * 3 + kx) * cin + ci) * cout_real + co];
      }
    }
    const float x = v * sw;
    const __half h1 = __float2half_rn(x);
    const __half h2 = __float2half_rn((x - __half2float(h1)) * 2048.f);
    const size_t row_hi = ((size_t)(kb * 9 + tap) * 2 + 0) * cout + co;
    const size_t row_lo = ((size_t)(kb * 9 + tap) * 2 + 1) * cout + co;
    out[row_hi * KC + c] = h1;
    out[row_lo * KC + c] = h2;
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

long long* g_conv_prof = nullptr;      // M4D_TC_PROFILE builds only (m4d_debug_conv_profile)

inline bool tc_shape_ok(int cin, int cout, int stride) {
  if (cin < 1 || cout < 1 || cout > 256) return false;
  if (stride == 1) return true;
  return stride == 2 && cin % 16 == 0;          // the (px, c) range of a cell row splits into whole 32-channel k-blocks
}
inline int tc_cout_pad(int cout) { return (cout + 15) / 16 * 16; }
inline int tc_kblocks(int cin, int stride) { return stride == 1 ? (cin + KC - 1) / KC : 2 * (2 * cin / KC); }

// Modelled clocks per k-step of one work item whose MMA N is n (shared-memory operand reads at 128 B/clk against the tensor
// pipe's N/2 clocks per 128x N x8 MMA): three MMAs for n > 64, two (N = 2n and n) in the concatenated mode.
inline int tc_kstep_clocks(int n) {
  auto mma = [](int nn) { const int smem = 32 + nn / 4, tensor = nn / 2; return smem > tensor ? smem : tensor; };
  return n > 64 ? 3 * mma(n) : mma(2 * n) + mma(n);
}
// Output-channel slices per tile.  One slice (N = all channels) does the least work per output, but a layer with fewer
// tiles than SMs (pyramid levels 4-6) is a long serial MMA chain on a few SMs: slicing the channels over 2 or 4 CTAs
// shortens the chain (N = 32: 88 clocks per k-step instead of 192 at N = 128) and fills the machine.  More than 128
// channels (the 192-channel encoder level) always needs slices.  Picks the split with the fewest modelled clocks.
inline int tc_pick_slices(int cout_pad, int64_t ntiles, int sms) {
  int best = 0;
  int64_t best_cost = 0;
  for (int ns = 1; ns <= 4; ns *= 2) {
    if (cout_pad % (16 * ns) != 0) continue;
    const int n = cout_pad / ns;
    if (n > 128 || (ns > 1 && n % 32 != 0)) continue;             // the epilogue's TMA stores are 32 channels wide
    const int64_t waves = (ntiles * ns + sms - 1) / sms;
    const int64_t cost = waves * (tc_kstep_clocks(n) + 12);       // + per-k-step issue / pipeline overhead of a work item
    if (best == 0 || cost < best_cost) { best = ns; best_cost = cost; }
  }
  return best;
}

inline int64_t tc_plane_floats(int cin, int cout, int stride, int prec) {       // packed weight planes, in floats
  const int64_t elems = (int64_t)tc_kblocks(cin, stride) * 9 * 2 * tc_cout_pad(cout) * KC;
  return prec == M4D_CONV_PREC_3XFP16 ? elems / 2 : elems;
}

int tc_launch(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w, int cin, int cout,
              int stride, int prec, float leaky_alpha, float* y, int y_pix_stride, int force, cudaStream_t stream) {
  const bool half = prec == M4D_CONV_PREC_3XFP16;
  const int force_slices = force & 15, force_na = (force >> 4) & 3;      // tuning: slices | halo stages << 4
  const int pdl = (force >> 8) & 3;       // M4D_CONV_PDL (1 << 8) [| M4D_CONV_PDL_WEIGHTS_STABLE (1 << 9)]: see include/m4d.h
  const float* w_scale = packed + tc_plane_floats(cin, cout, stride, prec);      // 3xFP16: 1 / s_w behind the planes
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) {
    m4d_set_error("m4d_conv3x3_tc_fwd: cuTensorMapEncodeTiled is not available from this driver");
    return M4D_ECUDA;
  }
  const int kb = tc_kblocks(cin, stride);
  const int cout_real = cout;
  const bool tma_out = y_pix_stride % 4 == 0 && !(reinterpret_cast<uintptr_t>(y) & 15u);
  const int ctot = tc_cout_pad(cout);          // rows per weight plane of the packed buffer
  const int oh = stride == 1 ? h : h / 2, ow = stride == 1 ? w : w / 2;
  const int64_t ntiles = (int64_t)((ow + TILE_W - 1) / TILE_W) * ((oh + TILE_H - 1) / TILE_H) * b;
  M4D_REQUIRE(ntiles < (1ll << 28), "m4d_conv3x3_tc_fwd: too many tiles");
  int nslices = force_slices > 0 ? force_slices : tc_pick_slices(ctot, ntiles, m4d_sm_count());
  M4D_REQUIRE(nslices >= 1 && ctot % nslices == 0 && ctot / nslices <= 128 && (nslices == 1 || (ctot / nslices) % 32 == 0),
              "m4d_conv3x3_tc_fwd: cannot slice %d output channels into %d (slices are multiples of 32 channels, at most 128)", cout_real,
              nslices);
  cout = ctot / nslices;                       // from here on: the MMA N = slice width
  CUtensorMap mx, mw, my;
  if (stride == 1) {
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
  } else {
    // [b][H/2][py][W/2][(px, c)]: x_pix_stride == cin, so the two pixels of a cell row are 2*cin contiguous floats
    const cuuint64_t ps = (cuuint64_t)cin * 4;
    const cuuint64_t dims[5] = {(cuuint64_t)2 * cin, (cuuint64_t)ow, 2, (cuuint64_t)oh, (cuuint64_t)b};
    const cuuint64_t strides[4] = {2 * ps, (cuuint64_t)w * ps, 2 * (cuuint64_t)w * ps, (cuuint64_t)h * w * ps};
    const cuuint32_t box[5] = {KC, HALO_W, 1, HALO_H, 1};
    const cuuint32_t es[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mx, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(x), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      m4d_set_error("m4d_conv3x3_tc_fwd: cuTensorMapEncodeTiled(x, stride 2) failed with %d", (int)r);
      return M4D_ECUDA;
    }
  }
  {
    const cuuint64_t dims[2] = {KC, (cuuint64_t)kb * 9 * 2 * ctot};
    const cuuint64_t strides[1] = {(cuuint64_t)KC * (half ? 2 : 4)};
    const cuuint32_t box[2] = {KC, (cuuint32_t)cout};         // one plane's rows of a slice; two loads per stage
    const cuuint32_t es[2] = {1, 1};
    CUresult r = enc(&mw, half ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(packed), dims, strides,
                     box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, half ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      m4d_set_error("m4d_conv3x3_tc_fwd: cuTensorMapEncodeTiled(weights) failed with %d", (int)r);
      return M4D_ECUDA;
    }
  }
  if (tma_out) {
    const cuuint64_t dims[4] = {(cuuint64_t)cout_real, (cuuint64_t)ow, (cuuint64_t)oh, (cuuint64_t)b};
    const cuuint64_t strides[3] = {(cuuint64_t)y_pix_stride * 4, (cuuint64_t)ow * y_pix_stride * 4, (cuuint64_t)oh * ow * y_pix_stride * 4};
    const cuuint32_t box[4] = {32, TILE_W, TILE_H, 1};
    const cuuint32_t es[4] = {1, 1, 1, 1};
    CUresult r = enc(&my, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, y, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                     CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      m4d_set_error("m4d_conv3x3_tc_fwd: cuTensorMapEncodeTiled(y) failed with %d", (int)r);
      return M4D_ECUDA;
    }
  } else {
    my = mw;                                   // unused by the kernel
  }
  TcArgs a;
  a.prof = g_conv_prof;
  a.pdl = pdl == 0 ? 0 : (pdl & 2 ? 2 : 1);
  a.thin_mode = ((force >> 11) & 1) ? 2 : 1;           // force bit 11 (tuning): one epilogue group + eight splitter warps
  a.bias = bias; a.y = y; a.h = oh; a.w = ow; a.cout = cout; a.cout_real = cout_real; a.tma_out = tma_out ? 1 : 0;
  a.ys = y_pix_stride; a.kblocks = kb; a.alpha = leaky_alpha; a.w_scale = w_scale;
  a.s2d_c = stride == 1 ? 0 : cin;
  a.s2d_chunks = stride == 1 ? 1 : 2 * cin / KC;
  a.cin = stride == 1 ? cin : 4 * cin;
  a.tiles_x = (ow + TILE_W - 1) / TILE_W;
  a.tiles_y = (oh + TILE_H - 1) / TILE_H;
  a.ntiles = (int)ntiles;
  a.nslices = nslices;
  a.nitems = a.ntiles * nslices;
  a.ctot = ctot;
  // TMEM: 512 columns = 2 accumulator sets of 256 that alternate per k-block; a set is [hi*hi | cross terms].
  // force bit 10 (tuning): three separate MMAs per k-step for N > 64 instead of the concatenated pair
  a.concat = (cout <= 64 || !((force >> 10) & 1)) ? 1 : 0;
  a.nacc = 2;
  a.nsets = 2;
  // weight slabs a tile streams through the ring: (k-block, tap) pairs that carry weights
  int slabs = 0;
  for (int k = 0; k < kb; ++k) {
    if (stride == 1) { slabs += 9; continue; }
    const int chunks = 2 * cin / KC, py = k / chunks, left = cin - (k - py * chunks) * KC;
    slabs += (py ? 1 : 2) * (left > 0 ? 2 : 1);
  }
  // halo stages: a k-block's chain TMA load -> split -> MMAs -> release is several thousand clocks of latency.  Two stages by
  // default; a third for the thinnest layers (Cout <= 32) where the whole layer's weights still stay resident beside it: measured
  // 16->16 stride 2 @384x1280 155 -> 142 us, 32->32 stride 2 85 -> 78, 16->32 95 -> 89, 32->16 97 -> 93 (wider layers lose their
  // resident weights to the third stage and get slower: 64->32 183 -> 195 us)
  const size_t a_stage = half ? A_STAGE_BYTES_F16 : A_STAGE_BYTES_F32;
  const size_t stage = (size_t)2 * cout * (half ? 64 : 128);
  auto fixed_for = [&](int n) { return (size_t)1024 + (size_t)n * a_stage + 2 * OUT_SLOT + 1024 + 1024; };   // alignment, halo stages, staging tiles, barriers, bias
  int na = force_na > 0 ? force_na : 2;
  if (force_na == 0 && cout <= 32 && nslices == 1 && MAX_A_STAGES >= 3 && 227 * 1024 >= fixed_for(3) + (size_t)slabs * stage) na = 3;
  size_t fixed = fixed_for(na);
  if (na > 2 && (227 * 1024 < fixed + 4 * stage)) {
    na = 2;
    fixed = fixed_for(na);
  }
  M4D_REQUIRE(na >= 2 && na <= MAX_A_STAGES && 227 * 1024 >= fixed + 2 * stage, "m4d_conv3x3_tc_fwd: not enough shared memory for the pipeline");
  int nb = (int)((227 * 1024 - fixed) / stage);
  if (nb > MAX_B_STAGES) nb = MAX_B_STAGES;
  a.nb = nb;
  a.na = na;
  // Thin layers issue a tap's MMAs in ~100-200 clocks, far less than the latency of the TMA load that refills its stage, so
  // the ring is deep (up to 32 slabs) and, where the whole layer fits, loaded once per CTA instead of once per tile.
  a.b_resident = (nslices == 1 && slabs <= nb) ? 1 : 0;
  const size_t smem = fixed + (size_t)nb * stage;
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, TcArgs);
  // at most 32 output channels per work item (3xFP16): the 20-warp variant with eight splitter warps (force bit 12: off)
  const bool ws = half && a.concat && cout <= 32 && !((force >> 12) & 1) && a.thin_mode != 2;
  static const KernelFn kernels_ws[2] = {conv3x3_tc_kernel<false, true, true, true>, conv3x3_tc_kernel<true, true, true, true>};
  static const KernelFn kernels[8] = {conv3x3_tc_kernel<false, false, false>, conv3x3_tc_kernel<false, true, false>,
                                      conv3x3_tc_kernel<true, false, false>,  conv3x3_tc_kernel<true, true, false>,
                                      conv3x3_tc_kernel<false, false, true>,  conv3x3_tc_kernel<false, true, true>,
                                      conv3x3_tc_kernel<true, false, true>,   conv3x3_tc_kernel<true, true, true>};
  // function attributes are per device: remember which devices of this process have them
  // (atomic flags: two host threads racing on the first call both set the attributes, which is idempotent)
  static std::atomic<bool> attr_set_dev[64];
  int dev_id = 0;
  cudaGetDevice(&dev_id);
  std::atomic<bool>& attr_set = attr_set_dev[dev_id & 63];
  if (!attr_set.load(std::memory_order_acquire)) {
    for (int i = 0; i < 10; ++i) {
      cudaError_t e = cudaFuncSetAttribute(i < 8 ? kernels[i] : kernels_ws[i - 8], cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
      if (e != cudaSuccess) {
        m4d_set_error("m4d_conv3x3_tc_fwd: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
        return M4D_ECUDA;
      }
    }
    attr_set.store(true, std::memory_order_release);
  }
  const int grid = a.nitems < m4d_sm_count() ? a.nitems : m4d_sm_count();
  KernelFn kern = ws ? kernels_ws[stride == 2 ? 1 : 0] : kernels[(half ? 4 : 0) + (stride == 2 ? 2 : 0) + (a.concat ? 1 : 0)];
  const int nthreads = ws ? NTHREADS + 128 : NTHREADS;
  if (a.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(nthreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, mx, mw, my, a);
    if (e != cudaSuccess) {
      m4d_set_error("m4d_conv3x3_tc_fwd: cudaLaunchKernelEx failed: %s", cudaGetErrorString(e));
      return M4D_ECUDA;
    }
  } else {
    kern<<<grid, nthreads, smem, stream>>>(mx, mw, my, a);
  }
  M4D_CHECK_LAUNCH("m4d_conv3x3_tc_fwd");
  return M4D_OK;
}

}  // namespace

extern "C" {

int m4d_debug_conv_profile(long long* device_buf) {
#ifdef M4D_TC_PROFILE
  g_conv_prof = device_buf;
  return 1;
#else
  (void)device_buf;
  return 0;                               // this build carries no role timers
#endif
}

int64_t m4d_conv3x3_tc_packed_floats_p(int cin, int cout, int stride, int prec) {
  if (!tc_shape_ok(cin, cout, stride) || (prec != M4D_CONV_PREC_3XTF32 && prec != M4D_CONV_PREC_3XFP16)) return 0;
  return tc_plane_floats(cin, cout, stride, prec) + 32;           // + the weight scale (3xFP16) and its scratch word
}

int m4d_conv3x3_tc_pack_p(const float* kernel_hwio, int cin, int cout, int stride, int prec, float* packed, void* stream) {
  M4D_REQUIRE(kernel_hwio && packed, "m4d_conv3x3_tc_pack: null pointer");
  M4D_REQUIRE(tc_shape_ok(cin, cout, stride),
              "m4d_conv3x3_tc_pack: unsupported shape cin=%d cout=%d stride=%d (cout <= 256; stride 2 needs cin %% 16 == 0)", cin, cout, stride);
  M4D_REQUIRE(prec == M4D_CONV_PREC_3XTF32 || prec == M4D_CONV_PREC_3XFP16, "m4d_conv3x3_tc_pack: unknown precision mode %d", prec);
  const int kb = tc_kblocks(cin, stride);
  const int cp = tc_cout_pad(cout);
  const int64_t n = (int64_t)kb * 9 * cp * KC;
  const int grid = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  cudaStream_t st = (cudaStream_t)stream;
  float* tail = packed + tc_plane_floats(cin, cout, stride, prec);
  if (prec == M4D_CONV_PREC_3XTF32) {
    conv3x3_tc_pack_kernel<<<grid, 256, 0, st>>>(kernel_hwio, cin, cout, 0, cout, cp, kb, stride == 1 ? 0 : cin, stride == 1 ? 1 : 2 * cin / KC,
                                                 packed);
    M4D_CHECK_LAUNCH("m4d_conv3x3_tc_pack");
    return M4D_OK;
  }
  cudaError_t e = cudaMemsetAsync(tail, 0, 32 * sizeof(float), st);
  if (e != cudaSuccess) {
    m4d_set_error("m4d_conv3x3_tc_pack: cudaMemsetAsync failed: %s", cudaGetErrorString(e));
    return M4D_ECUDA;
  }
  const int64_t nw = (int64_t)9 * cin * cout;
  conv3x3_tc_wmax_kernel<<<(int)((nw + 255) / 256 < 1024 ? (nw + 255) / 256 : 1024), 256, 0, st>>>(kernel_hwio, nw,
                                                                                                  reinterpret_cast<unsigned int*>(tail + 1));
  M4D_CHECK_LAUNCH("m4d_conv3x3_tc_pack(max)");
  conv3x3_tc_pack_f16_kernel<<<grid, 256, 0, st>>>(kernel_hwio, cin, cout, cp, kb, stride == 1 ? 0 : cin, stride == 1 ? 1 : 2 * cin / KC,
                                                   reinterpret_cast<const unsigned int*>(tail + 1), reinterpret_cast<__half*>(packed), tail);
  M4D_CHECK_LAUNCH("m4d_conv3x3_tc_pack");
  return M4D_OK;
}

int m4d_conv3x3_tc_fwd_p(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w, int cin,
                         int cout, int stride, int prec, float leaky_alpha, float* y, int y_pix_stride, int slices, void* stream) {
  M4D_REQUIRE(x && packed && bias && y, "m4d_conv3x3_tc_fwd: null pointer");
  M4D_REQUIRE(b > 0 && h > 0 && w > 0, "m4d_conv3x3_tc_fwd: non-positive size");
  M4D_REQUIRE(prec == M4D_CONV_PREC_3XTF32 || prec == M4D_CONV_PREC_3XFP16, "m4d_conv3x3_tc_fwd: unknown precision mode %d", prec);
  const bool s2_ok = stride != 2 || (h % 2 == 0 && w % 2 == 0 && x_pix_stride == cin);
  if (!tc_shape_ok(cin, cout, stride) || !s2_ok || x_pix_stride % 4 != 0 || x_pix_stride < cin || y_pix_stride < cout ||
      (reinterpret_cast<uintptr_t>(x) & 15u) || (reinterpret_cast<uintptr_t>(packed) & 15u)) {
    m4d_set_error("m4d_conv3x3_tc_fwd: shape / alignment outside the tcgen05 path (cin=%d cout=%d stride=%d h=%d w=%d xs=%d ys=%d)", cin, cout,
                  stride, h, w, x_pix_stride, y_pix_stride);
    return M4D_ENOTSUP;
  }
  return tc_launch(x, x_pix_stride, packed, bias, b, h, w, cin, cout, stride, prec, leaky_alpha, y, y_pix_stride, slices, (cudaStream_t)stream);
}

int64_t m4d_conv3x3_tc_packed_floats_s(int cin, int cout, int stride) {
  return m4d_conv3x3_tc_packed_floats_p(cin, cout, stride, M4D_CONV_PREC_3XTF32);
}
int m4d_conv3x3_tc_pack_s(const float* kernel_hwio, int cin, int cout, int stride, float* packed, void* stream) {
  return m4d_conv3x3_tc_pack_p(kernel_hwio, cin, cout, stride, M4D_CONV_PREC_3XTF32, packed, stream);
}
int m4d_conv3x3_tc_fwd_ex(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w, int cin,
                          int cout, int stride, float leaky_alpha, float* y, int y_pix_stride, int slices, void* stream) {
  return m4d_conv3x3_tc_fwd_p(x, x_pix_stride, packed, bias, b, h, w, cin, cout, stride, M4D_CONV_PREC_3XTF32, leaky_alpha, y, y_pix_stride,
                              slices, stream);
}
int m4d_conv3x3_tc_fwd_s(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w, int cin,
                         int cout, int stride, float leaky_alpha, float* y, int y_pix_stride, void* stream) {
  return m4d_conv3x3_tc_fwd_ex(x, x_pix_stride, packed, bias, b, h, w, cin, cout, stride, leaky_alpha, y, y_pix_stride, 0, stream);
}

int64_t m4d_conv3x3_tc_packed_floats(int cin, int cout) { return m4d_conv3x3_tc_packed_floats_s(cin, cout, 1); }
int m4d_conv3x3_tc_pack(const float* kernel_hwio, int cin, int cout, float* packed, void* stream) {
  return m4d_conv3x3_tc_pack_s(kernel_hwio, cin, cout, 1, packed, stream);
}
int m4d_conv3x3_tc_fwd(const float* x, int x_pix_stride, const float* packed, const float* bias, int b, int h, int w, int cin,
                       int cout, float leaky_alpha, float* y, int y_pix_stride, void* stream) {
  return m4d_conv3x3_tc_fwd_s(x, x_pix_stride, packed, bias, b, h, w, cin, cout, 1, leaky_alpha, y, y_pix_stride, stream);
}

}  // extern "C"

// conv3x3.cu's algo = 2: the tensor-core path needs the pre-packed weights of m4d_conv3x3_tc_pack
int m4d_conv3x3_tc(const float*, int, const float*, const float*, int, int, int, int, int, int, float, float*, int, cudaStream_t) {
  m4d_set_error("m4d_conv3x3_nhwc: algo 2 needs pre-packed weights: use m4d_conv3x3_tc_pack + m4d_conv3x3_tc_fwd");
  return M4D_ENOTSUP;
}
