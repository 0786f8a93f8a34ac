// Stand-in for tensorflow/core/platform/types.h (see ../framework/tensor_types.h): the reference CUDA file needs nothing from it.
#ifndef M4D_ORACLE_REF_STUB_TYPES_H_
#define M4D_ORACLE_REF_STUB_TYPES_H_
namespace tensorflow {}
#endif
