// Stand-in for TensorFlow's tensorflow/core/framework/tensor_types.h, just large enough to compile the reference's
// cuda_backproject/backproject_op_gpu.cu.cc UNMODIFIED outside TensorFlow (test infrastructure: oracle/Makefile builds
// oracle/_ref/libbackproject_ref.so from the reference sources where they lie).  The reference file uses exactly two members
// of Eigen::GpuDevice: stream() for the launch (backproject_op_gpu.cu.cc:92,210) and ok() for its return value (:102,222).
#ifndef M4D_ORACLE_REF_STUB_TENSOR_TYPES_H_
#define M4D_ORACLE_REF_STUB_TENSOR_TYPES_H_
#include <cuda_runtime.h>

namespace Eigen {
struct GpuDevice {
  cudaStream_t stream_;
  explicit GpuDevice(cudaStream_t s) : stream_(s) {}
  cudaStream_t stream() const { return stream_; }
  bool ok() const { return cudaPeekAtLastError() == cudaSuccess; }
};
}  // namespace Eigen

namespace tensorflow {}
#endif
