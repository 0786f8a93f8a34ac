// Latency of 1-D bulk copies (cp.async.bulk global -> shared, mbarrier completion) on the GPU box, as the PSCV kernel uses them:
// R row copies of B bytes each issued by lane 0 of W warps, then all threads wait on the mbarrier.  Sources are distinct,
// never-touched global addresses (DRAM) or a small re-read buffer (L2).  Reports clocks from first issue to wake-up.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/tma_probe tools/tma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void* src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__global__ void __launch_bounds__(256) tma_kernel(const unsigned char* src, size_t stride_cta, size_t stride_iter, int iters, int rows, int row_bytes,
                                                  size_t row_pitch, long long* out) {
  extern __shared__ __align__(128) unsigned char sm[];
  __shared__ __align__(8) unsigned long long bar_store;
  const unsigned bar = (unsigned)__cvta_generic_to_shared(&bar_store), s0 = (unsigned)__cvta_generic_to_shared(sm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) { mbar_init(bar, 8); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  long long tot = 0, mx = 0;
  unsigned ph = 0;
  for (int it = 0; it < iters; ++it) {
    const unsigned char* base = src + (size_t)blockIdx.x * stride_cta + (size_t)it * stride_iter;
    __syncthreads();
    const long long t0 = clock64();
    if (lane == 0) {
      const int n = rows > warp ? (rows - warp + 7) / 8 : 0;
      mbar_expect_tx(bar, (unsigned)(n * row_bytes));
      for (int r = warp; r < rows; r += 8) bulk_g2s(s0 + (unsigned)(r * row_bytes), base + (size_t)r * row_pitch, (unsigned)row_bytes, bar);
    }
    __syncwarp();
    mbar_wait(bar, ph);
    ph ^= 1u;
    const long long t1 = clock64();
    if (tid == 0) { tot += t1 - t0; if (t1 - t0 > mx) mx = t1 - t0; }
  }
  if (tid == 0) { out[blockIdx.x * 2] = tot / iters; out[blockIdx.x * 2 + 1] = mx; }
}

int main() {
  unsigned char* d;
  const size_t bytes = (size_t)3 << 30;
  cudaMalloc(&d, bytes);
  cudaMemset(d, 1, bytes);
  long long* d_out;
  cudaMalloc(&d_out, 296 * 16);
  cudaFuncSetAttribute(tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  struct Cfg { const char* name; int grid, rows, row_bytes; size_t pitch; bool dram; };
  const Cfg cfgs[] = {{"1 row x 2 KB, DRAM", 148, 1, 2048, 40960, true},      {"8 rows x 2 KB (c1 tile), DRAM", 148, 8, 2048, 40960, true},
                      {"14 rows x 3 KB (window in situ), DRAM", 148, 14, 3072, 40960, true}, {"20 rows x 4.5 KB (window micro), DRAM", 148, 20, 4608, 40960, true},
                      {"14 rows x 3 KB, DRAM, 2 CTAs/SM", 296, 14, 3072, 40960, true},    {"14 rows x 3 KB, L2 resident", 148, 14, 3072, 40960, false},
                      {"1 row x 2 KB, L2 resident", 148, 1, 2048, 40960, false}};
  for (const Cfg& c : cfgs) {
    const int iters = 8;
    const size_t per_iter = (size_t)40960 * 24;                          // a fresh region per iteration
    const size_t stride_cta = c.dram ? per_iter * iters : per_iter;
    const size_t stride_iter = c.dram ? per_iter : 0;
    if (!c.dram) tma_kernel<<<c.grid, 256, 96 * 1024>>>(d, stride_cta, stride_iter, 2, c.rows, c.row_bytes, c.pitch, d_out);   // warm the L2
    else { cudaMemset(d + ((size_t)2 << 30), 2, (size_t)1 << 30); }     // evict
    tma_kernel<<<c.grid, 256, 96 * 1024>>>(d, stride_cta, stride_iter, iters, c.rows, c.row_bytes, c.pitch, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
    long long h[592];
    cudaMemcpy(h, d_out, c.grid * 16, cudaMemcpyDeviceToHost);
    double avg = 0, mx = 0;
    for (int i = 0; i < c.grid; ++i) { avg += h[2 * i]; mx = h[2 * i + 1] > mx ? h[2 * i + 1] : mx; }
    printf("%-44s: mean %6.0f clk from first issue to wake-up (max %6.0f), %d CTAs\n", c.name, avg / c.grid, mx, c.grid);
  }
  return 0;
}
