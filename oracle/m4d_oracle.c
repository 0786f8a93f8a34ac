/* CPU oracle, plain C: the BackProject op's forward and gradient as the reference's CUDA kernels compute them, one sample
 * per loop iteration instead of one per thread.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): built by oracle/Makefile
 * into oracle/libm4d_oracle.so and loaded by tests/test_oracle_golden.py to cross-check the torch restatement in
 * oracle/warp.py; nothing under m4depth_b200/ may link or load it.
 *
 * Follows cuda_backproject/backproject_op_gpu.cu.cc:
 *   bp_forward   :19-79   (index decomposition :33-37, guard :47, floor/ceil taps :48-51, weights :53-59, sum :74; the output is
 *                          zero where the guard fails - the launcher's cudaMemset, :91)
 *   bp_backward  :108-196 (scatter of grad * w onto the four taps :178-181 - serial here, atomicAdd there -, coordinate
 *                          gradient :183-184; both outputs start from zero, :209-210)
 * Compile with -ffp-contract=off: every product and sum is rounded separately, the reading the torch oracle uses
 * (the reference binary may contract them into FMAs; <= 1 ulp of the largest term). */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* dim = {B,H,W,S,F,C}; input [B,H,W,F,C]; coords [B,H,W,S,F,2] (x,y); out [B,H,W,S,F,C]; idx (nullable) [B,H,W,S,F,4] */
void bp_forward(const float* input, const float* coords, const int32_t* dim, float* out, int32_t* idx) {
  const int64_t H = dim[1], W = dim[2], S = dim[3], F = dim[4], C = dim[5];
  const int64_t nsamp = (int64_t)dim[0] * H * W * S * F;
  memset(out, 0, (size_t)nsamp * C * sizeof(float));
  for (int64_t index = 0; index < nsamp; ++index) {
    int64_t n = index;
    const int64_t f = n % F; n /= F;
    n /= S;
    n /= W;
    n /= H;                                                     /* n = batch index (:33-37) */
    const float x = coords[2 * index], y = coords[2 * index + 1];
    const int inside = x >= 0 && y >= 0 && x <= (float)(W - 1) && y <= (float)(H - 1);       /* false for NaN (:47) */
    if (idx) {
      int32_t* t = idx + 4 * index;
      t[0] = inside ? (int32_t)floorf(x) : -1; t[1] = inside ? (int32_t)ceilf(x) : -1;
      t[2] = inside ? (int32_t)floorf(y) : -1; t[3] = inside ? (int32_t)ceilf(y) : -1;
    }
    if (!inside) continue;
    const int64_t x0 = (int64_t)floorf(x), x1 = (int64_t)ceilf(x), y0 = (int64_t)floorf(y), y1 = (int64_t)ceilf(y);
    const float dx = x - (float)x0, dy = y - (float)y0;
    const float w00 = (1 - dy) * (1 - dx), w01 = (1 - dy) * dx, w10 = dy * (1 - dx), w11 = dy * dx;
    const int64_t offset = (n * H * W * F + f) * C;
    const float* im00 = input + offset + F * C * (y0 * W + x0);
    const float* im01 = input + offset + F * C * (y0 * W + x1);
    const float* im10 = input + offset + F * C * (y1 * W + x0);
    const float* im11 = input + offset + F * C * (y1 * W + x1);
    float* o = out + index * C;
    for (int64_t c = 0; c < C; ++c) o[c] = im00[c] * w00 + im01[c] * w01 + im10[c] * w10 + im11[c] * w11;      /* :74 */
  }
}

/* grad [B,H,W,S,F,C] -> input_grad [B,H,W,F,C], coords_grad [B,H,W,S,F,2] */
void bp_backward(const float* grad, const float* input, const float* coords, const int32_t* dim, float* input_grad,
                 float* coords_grad) {
  const int64_t H = dim[1], W = dim[2], S = dim[3], F = dim[4], C = dim[5];
  const int64_t nsamp = (int64_t)dim[0] * H * W * S * F;
  memset(input_grad, 0, (size_t)dim[0] * H * W * F * C * sizeof(float));
  memset(coords_grad, 0, (size_t)nsamp * 2 * sizeof(float));
  for (int64_t index = 0; index < nsamp; ++index) {
    int64_t n = index;
    const int64_t f = n % F; n /= F;
    n /= S;
    n /= W;
    n /= H;
    const float x = coords[2 * index], y = coords[2 * index + 1];
    if (!(x >= 0 && y >= 0 && x <= (float)(W - 1) && y <= (float)(H - 1))) continue;
    const int64_t x0 = (int64_t)floorf(x), x1 = (int64_t)ceilf(x), y0 = (int64_t)floorf(y), y1 = (int64_t)ceilf(y);
    const float dx = x - (float)x0, dy = y - (float)y0;
    const float wx0 = 1 - dx, wx1 = dx, wy0 = 1 - dy, wy1 = dy;
    const float w00 = (1 - dy) * (1 - dx), w01 = (1 - dy) * dx, w10 = dy * (1 - dx), w11 = dy * dx;
    const int64_t offset = (n * H * W * F + f) * C;
    const int64_t i00 = offset + F * C * (y0 * W + x0), i01 = offset + F * C * (y0 * W + x1);
    const int64_t i10 = offset + F * C * (y1 * W + x0), i11 = offset + F * C * (y1 * W + x1);
    const float* g = grad + index * C;
    float gx = 0, gy = 0;
    for (int64_t c = 0; c < C; ++c) {
      const float gv = g[c];
      input_grad[i00 + c] += gv * w00;                          /* atomicAdd in the reference (:178-181) */
      input_grad[i01 + c] += gv * w01;
      input_grad[i10 + c] += gv * w10;
      input_grad[i11 + c] += gv * w11;
      gx += gv * (wy0 * (input[i01 + c] - input[i00 + c]) + wy1 * (input[i11 + c] - input[i10 + c]));          /* :183 */
      gy += gv * (wx0 * (input[i10 + c] - input[i00 + c]) + wx1 * (input[i11 + c] - input[i01 + c]));          /* :184 */
    }
    coords_grad[2 * index] = gx;
    coords_grad[2 * index + 1] = gy;
  }
}
