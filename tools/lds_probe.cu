// Shared-memory load throughput on the GPU box as a function of access width and of how a warp's lanes spread over
// 128-byte rows (design input for csrc/pscv_smem.cu: which gather mapping gets the most bytes per clock).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/lds_probe tools/lds_probe.cu
// Patterns (all bank-conflict free in the classic 32 x 4-byte-bank sense):
//   W=16 (LDS.128): G lanes share a 128-byte row, G in {8 (one row per quarter warp), 4, 2 (rotated quads), 1}
//   W=8  (LDS.64) : G in {16, 8, 4, 2}
//   W=4  (LDS.32) : G in {32 (lane = channel), 16, 8}
// Rows are pseudo-random (different per warp, per load, per lane group).
#include <cstdio>
#include <cuda_runtime.h>

template <int WIDTH, int G>
__global__ void lds_kernel(long long* out, int iters) {
  extern __shared__ __align__(128) unsigned char sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < 16384; i += blockDim.x) reinterpret_cast<unsigned*>(sm)[i] = i;
  __syncthreads();
  const unsigned base = (unsigned)__cvta_generic_to_shared(sm);
  // lane group = lanes sharing a row; inside the row the group reads G consecutive WIDTH-byte chunks starting at a slot that
  // rotates with the group index so that the lanes of one LDS phase (128 bytes) cover all 32 banks
  const int grp = lane / G, sub = lane % G;
  const int chunks_per_row = 128 / WIDTH;
  const int slot = (grp * G + sub) % chunks_per_row;
  const unsigned lane_off = (unsigned)(slot * WIDTH);
  unsigned acc0 = 0, acc1 = 0, acc2 = 0, acc3 = 0;
  unsigned rowsel = (unsigned)(warp * 37 + grp * 101);
  long long t0 = clock64();
#pragma unroll 1
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      const unsigned a = base + (((rowsel + (unsigned)(u * 53)) & 255u) << 7) + lane_off;
      if (WIDTH == 16) {
        unsigned x, y, z, w;
        asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(x), "=r"(y), "=r"(z), "=r"(w) : "r"(a));
        acc0 += x; acc1 += y; acc2 += z; acc3 += w;
      } else if (WIDTH == 8) {
        unsigned x, y;
        asm volatile("ld.shared.v2.u32 {%0,%1}, [%2];" : "=r"(x), "=r"(y) : "r"(a));
        acc0 += x; acc1 += y;
      } else {
        unsigned x;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(x) : "r"(a));
        acc0 += x;
      }
    }
    rowsel += 7u;
  }
  long long t1 = clock64();
  if (tid == 0) out[blockIdx.x * 2] = t1 - t0;
  if (acc0 + acc1 + acc2 + acc3 == 0x12345678u) out[blockIdx.x * 2 + 1] = acc0;
}

template <int WIDTH, int G>
void run(int warps, long long* d_out) {
  const int iters = 2000;
  cudaFuncSetAttribute(lds_kernel<WIDTH, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  lds_kernel<WIDTH, G><<<148, warps * 32, 65536>>>(d_out, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return; }
  long long h[2];
  cudaMemcpy(h, d_out, 16, cudaMemcpyDeviceToHost);
  const double clk = (double)h[0];
  const double bytes = (double)warps * 32 * WIDTH * 16 * iters;
  printf("LDS.%-3d %2d lanes per row, %2d warps/SM: %5.2f clk per warp-instruction, %6.1f B/clk/SM\n", WIDTH * 8, G, warps, clk / (16.0 * iters * warps),
         bytes / clk);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 16);
  for (int warps : {8, 16, 32}) {
    run<4, 32>(warps, d_out); run<4, 16>(warps, d_out); run<4, 8>(warps, d_out);
    run<8, 16>(warps, d_out); run<8, 8>(warps, d_out); run<8, 4>(warps, d_out); run<8, 2>(warps, d_out);
    run<16, 8>(warps, d_out); run<16, 4>(warps, d_out); run<16, 2>(warps, d_out); run<16, 1>(warps, d_out);
  }
  return 0;
}
