// Standalone probe of the shared-memory staged PSCV kernel (csrc/pscv_smem.cu) for the GPU box: synthetic level-2 inputs,
// CUDA-event timing with an L2 flush between launches, and per-phase clock64 stamps of every tile (M4D_PSCV_PROF).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/build/pscv_probe tools/pscv_probe.cu
//   tools/build/pscv_probe [micro|insitu] [ctas_per_sm]
// An iteration tool, not part of the library or the bench contract.
#define M4D_PSCV_PROF 1
#include "../m4depth_b200/csrc/pscv_smem.cu"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <random>
#include <string>
#include <vector>

void m4d_set_error(const char*, ...) {}
std::atomic<uint64_t> g_m4d_launches{0};

#define CK(x)                                                                              \
  do {                                                                                     \
    cudaError_t e_ = (x);                                                                  \
    if (e_ != cudaSuccess) {                                                               \
      fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_));   \
      exit(1);                                                                             \
    }                                                                                      \
  } while (0)

int main(int argc, char** argv) {
  const std::string mode = argc > 1 ? argv[1] : "micro";
#ifndef PROBE_TH
#define PROBE_TH 8
#endif
  const int cps = argc > 2 ? atoi(argv[2]) : SCfg<32, 2, PROBE_TH>::CTAS;
  const int b = 8, h = 96, w = 320, c = 32, cuts = 2, xs = 124;
  const size_t npix = (size_t)b * h * w;
  std::mt19937 rng(1);
  std::normal_distribution<float> nd(0.f, 1.f);
  std::uniform_real_distribution<float> ud(0.f, 1.f);
  auto feat = [&]() {
    std::vector<float> f(npix * c);
    for (size_t p = 0; p < npix; ++p)
      for (int g = 0; g < cuts; ++g) {
        float n = 0.f;
        float* q = &f[p * c + g * 16];
        for (int j = 0; j < 16; ++j) { float v = nd(rng); v = v >= 0 ? v : 0.1f * v; q[j] = v; n += v * v; }
        n = 1.f / std::sqrt(n);
        for (int j = 0; j < 16; ++j) q[j] *= n;
      }
    return f;
  };
  std::vector<float> c1 = feat(), c2 = feat(), pl(npix), pt(npix), rot(b * 4), trans(b * 3), cf(b * 2), cc(b * 2);
  for (size_t p = 0; p < npix; ++p) {
    pl[p] = mode == "insitu" ? 0.6f + 1.7f * ud(rng) : std::exp(ud(rng) * (std::log(16.f) - std::log(.5f)) + std::log(.5f));
    pt[p] = std::exp(ud(rng) * (std::log(16.f) - std::log(.05f)) + std::log(.05f));
  }
  for (int i = 0; i < b; ++i) {
    float q[4] = {1.f, 0.01f * nd(rng), 0.01f * nd(rng), 0.01f * nd(rng)};
    float n = std::sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int j = 0; j < 4; ++j) rot[i * 4 + j] = q[j] / n;
    trans[i * 3 + 0] = 0.05f * nd(rng); trans[i * 3 + 1] = 0.05f * nd(rng); trans[i * 3 + 2] = 1.f + 0.3f * nd(rng);
    cf[i * 2] = 0.580948f * w; cf[i * 2 + 1] = 1.924101f * h; cc[i * 2] = 0.490788f * w; cc[i * 2 + 1] = 0.460944f * h;
  }
  auto up = [&](const std::vector<float>& v) {
    float* d;
    CK(cudaMalloc(&d, v.size() * 4));
    CK(cudaMemcpy(d, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return d;
  };
  float *d_c1 = up(c1), *d_c2 = up(c2), *d_pl = up(pl), *d_pt = up(pt), *d_rot = up(rot), *d_tr = up(trans), *d_cf = up(cf), *d_cc = up(cc);
  float* d_x;
  CK(cudaMalloc(&d_x, npix * xs * 4));
  char* d_flush;
  CK(cudaMalloc(&d_flush, 256u << 20));

  typedef SCfg<32, 2, PROBE_TH> Cfg;
  SArgs sa;
  PscvArgs& a = sa.a;
  memset(&sa, 0, sizeof(sa));
  a.c1 = d_c1; a.c2 = d_c2; a.para_t = d_pt; a.para_l = d_pl; a.rot = d_rot; a.trans = d_tr; a.cam_f = d_cf; a.cam_c = d_cc;
  a.cv = d_x; a.prev_disp = nullptr; a.centre_log = d_x + 121; a.idx_dbg = nullptr;
  a.rot_dim = 4; a.b = b; a.h = h; a.w = w; a.c = c; a.cuts = cuts; a.r = 4; a.K = 9; a.Q = c / 4;
  a.cv_stride = xs; a.pd_stride = 0; a.cl_stride = xs; a.cl_scale = 0.5f; a.npix = (int64_t)npix;
  a.one = 1.f; a.neg_one = -1.f; a.neg_zero = -0.f;
  sa.tiles_x = (w + Cfg::TW - 1) / Cfg::TW; sa.tiles_y = (h + Cfg::TH - 1) / Cfg::TH; sa.n_tiles = sa.tiles_x * sa.tiles_y * b;
  sa.cv_vec2 = 1; sa.pl_bulk = 1; sa.l2_prefetch = argc > 3 ? atoi(argv[3]) : 1;
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  const int grid = std::min(sa.n_tiles, sms * cps);
  sa.prof_iters = (sa.n_tiles + grid - 1) / grid;
  const size_t nprof = (size_t)grid * sa.prof_iters * 2 * 10;
  CK(cudaMalloc(&sa.prof, nprof * 8));
  CK(cudaMemset(sa.prof, 0, nprof * 8));
  auto kern = pscv9s_kernel<32, 2, PROBE_TH, false>;
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  CK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  int occ = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, Cfg::NT, Cfg::SMEM));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  std::vector<float> ts;
  for (int it = 0; it < 12; ++it) {
    CK(cudaMemsetAsync(d_flush, it, 256u << 20));
    CK(cudaEventRecord(e0));
    kern<<<grid, Cfg::NT, Cfg::SMEM>>>(sa);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    ts.push_back(ms * 1e3f);
  }
  std::sort(ts.begin(), ts.end());
  printf("%s: grid %d x %d threads, %d B smem, occupancy %d CTAs/SM, %d tiles (%d per CTA max): median %.1f us, min %.1f us\n", mode.c_str(), grid,
         Cfg::NT, Cfg::SMEM, occ, sa.n_tiles, sa.prof_iters, ts[ts.size() / 2], ts[0]);
  std::vector<long long> prof(nprof);
  CK(cudaMemcpy(prof.data(), sa.prof, nprof * 8, cudaMemcpyDeviceToHost));
  const char* seg[] = {"wait c1 / para_l", "phase 0 records", "barrier after phase 0", "box + issue c2 copies", "centre_log + c1 regs + wait c2",
                       "sweep (phase 1)", "output", "prefetch + end barrier"};
  for (int wsel = 0; wsel < 2; ++wsel) {
    double sum[8] = {0}, tot = 0;
    long n = 0;
    for (int cta = 0; cta < grid; ++cta)
      for (int it = 0; it < sa.prof_iters; ++it) {
        const long long* q = &prof[(((size_t)cta * sa.prof_iters + it) * 2 + wsel) * 10];
        if (q[0] == 0 || q[8] == 0) continue;
        for (int s_ = 0; s_ < 8; ++s_) sum[s_] += (double)(q[s_ + 1] - q[s_]);
        tot += (double)(q[8] - q[0]);
        ++n;
      }
    printf("warp %d: %ld tiles, mean tile time %.0f clk\n", wsel ? 7 : 0, n, tot / n);
    for (int s_ = 0; s_ < 8; ++s_) printf("   %-32s %8.0f clk  %5.1f%%\n", seg[s_], sum[s_] / n, 100.0 * sum[s_] / tot);
  }
  // whole-CTA span vs tile sum: time outside the tile loop / launch skew
  double span = 0;
  long long tmin = LLONG_MAX, tmax = 0;
  for (int cta = 0; cta < grid; ++cta) {
    long long first = prof[((size_t)cta * sa.prof_iters) * 2 * 10], last = 0;
    for (int it = 0; it < sa.prof_iters; ++it) last = std::max(last, prof[(((size_t)cta * sa.prof_iters + it) * 2) * 10 + 8]);
    span += (double)(last - first);
    (void)tmin; (void)tmax;
  }
  printf("mean CTA span %.0f clk (kernel %.1f us = %.0f clk at 1.9 GHz)\n", span / grid, ts[0], ts[0] * 1900.0);
  return 0;
}
