// C-ABI doors into the REFERENCE's own BackProject CUDA launchers (cuda_backproject/backproject_op_gpu.h:17-22), compiled
// unmodified from /root/reference by oracle/Makefile into oracle/_ref/libbackproject_ref.so.  TEST INFRASTRUCTURE ONLY: the
// parity tests use it as the oracle for m4d_backproject_fwd / m4d_backproject_bwd and the BP_FMA warp of the fused PSCV,
// bench.py times it as the GPU kernel the fused path replaces.  Nothing under m4depth_b200/ loads it.
#define GOOGLE_CUDA 1
#include "backproject_op_gpu.h"

// backproject_op_gpu.cu.cc comments its `namespace tensorflow {` out and says `using namespace tensorflow;` instead (:16-17), so
// the launchers it DEFINES live in the global namespace (the header's tensorflow:: declarations stay undefined).
bool BackProjectForwardLauncher(const float* input, const float* coords, const int dim[6], float* top, const Eigen::GpuDevice& d);
bool BackProjectBackwardLauncher(const float* grad, const float* input, const float* coords, const int dim[6], float* inputs_diff,
                                 float* coords_diff, const Eigen::GpuDevice& d);

extern "C" {

// 0 = ok.  The reference memsets on the legacy default stream (backproject_op_gpu.cu.cc:91,209-210), so callers synchronise
// around these calls; a launch error makes the reference exit(-1) (:95-100).
int ref_backproject_fwd(const float* input, const float* coords, const int dim[6], float* top, void* stream) {
  Eigen::GpuDevice d((cudaStream_t)stream);
  return ::BackProjectForwardLauncher(input, coords, dim, top, d) ? 0 : 1;
}

int ref_backproject_bwd(const float* grad, const float* input, const float* coords, const int dim[6], float* inputs_diff,
                        float* coords_diff, void* stream) {
  Eigen::GpuDevice d((cudaStream_t)stream);
  return ::BackProjectBackwardLauncher(grad, input, coords, dim, inputs_diff, coords_diff, d) ? 0 : 1;
}

}  // extern "C"
