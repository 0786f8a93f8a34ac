// libm4d: tcgen05 (5th-gen tensor core) path of the 3x3 convolution.  Placeholder until the 3xTF32 implicit-GEMM
// kernel lands: reports M4D_ENOTSUP so that algo=0 (auto) falls through to the FFMA2 kernel in conv3x3.cu.
#include "common.cuh"

int m4d_conv3x3_tc(const float*, int, const float*, const float*, int, int, int, int, int, int, float, float*, int,
                   cudaStream_t) {
  m4d_set_error("m4d_conv3x3_nhwc: the tcgen05 path is not built for this shape");
  return M4D_ENOTSUP;
}
