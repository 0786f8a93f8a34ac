// TensorFlow custom-op shim over libm4d's C ABI (include/m4d.h).
//
// NOT BUILT IN THIS REPOSITORY'S IMAGE: TensorFlow (headers + libtensorflow_framework) is not installed here nor on the
// GPU box, so this file is compiled only by integration/tf_shim/build.sh on a machine that has TF 2.x.  It exists so that
// a maintainer of the reference can keep `utils/dense_image_warp.py:36-52` byte for byte: the shared object registers the
// SAME op names ("BackProject", "BackProjectGrad": cuda_backproject/backproject_op.cc:32-42,163-164) and forwards them to
// m4d_backproject_fwd / m4d_backproject_bwd, plus fused ops for the hot path (M4dPscvFused + its gradient, M4dSncv) that replace
// utils/depth_operations.py:223-281 and :283-313.  No arithmetic lives here: shapes, allocation, stream, error mapping.
//
// Differences from the reference's op library, on purpose:
//   * errors become OP_REQUIRES failures (errors::Internal / InvalidArgument) instead of fprintf + exit(-1)
//     (backproject_op_gpu.cu.cc:95-100, 215-220);
//   * all device work is enqueued on the op's own stream - the reference's cudaMemset calls run on the legacy default
//     stream (:91-93, 209-213);
//   * shape functions are registered (the reference has none), so the ops work under tf.function without set_shape.
#include "tensorflow/core/framework/op.h"
#include "tensorflow/core/framework/op_kernel.h"
#include "tensorflow/core/framework/shape_inference.h"

#define EIGEN_USE_GPU
#include "tensorflow/core/util/gpu_kernel_helper.h"

#include "../../include/m4d.h"

using namespace tensorflow;  // NOLINT

namespace {

void* stream_of(OpKernelContext* ctx) { return (void*)ctx->eigen_device<Eigen::GpuDevice>().stream(); }

#define M4D_TF_CHECK(ctx, rc) OP_REQUIRES(ctx, (rc) == M4D_OK, errors::Internal("libm4d: ", m4d_last_error_string()))

}  // namespace

// ------------------------------------------------------------------------------------------------ BackProject
REGISTER_OP("BackProject")
    .Input("inputs: float32")
    .Input("coords: float32")
    .Output("output: float32")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      shape_inference::ShapeHandle in, co;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 5, &in));   // [B,H,W,F,C]
      TF_RETURN_IF_ERROR(c->WithRank(c->input(1), 6, &co));   // [B,H,W,S,F,2]
      c->set_output(0, c->MakeShape({c->Dim(co, 0), c->Dim(co, 1), c->Dim(co, 2), c->Dim(co, 3), c->Dim(co, 4), c->Dim(in, 4)}));
      return Status();
    });

REGISTER_OP("BackProjectGrad")
    .Input("inputs: float32")
    .Input("coords: float32")
    .Input("grad: float32")
    .Output("inputs_grad: float32")
    .Output("coords_grad: float32")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      c->set_output(0, c->input(0));
      c->set_output(1, c->input(1));
      return Status();
    });

class BackProjectM4d : public OpKernel {
 public:
  explicit BackProjectM4d(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& in = ctx->input(0);
    const Tensor& co = ctx->input(1);
    OP_REQUIRES(ctx, in.dims() == 5 && co.dims() == 6 && co.dim_size(5) == 2,
                errors::InvalidArgument("BackProject: inputs must be [B,H,W,F,C] and coords [B,H,W,S,F,2]"));
    const int32_t dim[6] = {(int32_t)co.dim_size(0), (int32_t)co.dim_size(1), (int32_t)co.dim_size(2),
                            (int32_t)co.dim_size(3), (int32_t)co.dim_size(4), (int32_t)in.dim_size(4)};   // backproject_op.cc:62-76
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({dim[0], dim[1], dim[2], dim[3], dim[4], dim[5]}), &out));
    if (out->NumElements() == 0) return;
    M4D_TF_CHECK(ctx, m4d_backproject_fwd(in.flat<float>().data(), co.flat<float>().data(), dim, out->flat<float>().data(),
                                          nullptr, stream_of(ctx)));
  }
};

class BackProjectGradM4d : public OpKernel {
 public:
  explicit BackProjectGradM4d(OpKernelConstruction* c) : OpKernel(c) {}
  void Compute(OpKernelContext* ctx) override {
    const Tensor& in = ctx->input(0);
    const Tensor& co = ctx->input(1);
    const Tensor& gr = ctx->input(2);
    OP_REQUIRES(ctx, in.dims() == 5 && co.dims() == 6 && gr.dims() == 6,
                errors::InvalidArgument("BackProjectGrad: inputs [B,H,W,F,C], coords [B,H,W,S,F,2], grad [B,H,W,S,F,C]"));
    const int32_t dim[6] = {(int32_t)co.dim_size(0), (int32_t)co.dim_size(1), (int32_t)co.dim_size(2),
                            (int32_t)co.dim_size(3), (int32_t)co.dim_size(4), (int32_t)in.dim_size(4)};   // backproject_op.cc:117-144
    Tensor *ig = nullptr, *cg = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, in.shape(), &ig));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, co.shape(), &cg));
    if (gr.NumElements() == 0) return;
    M4D_TF_CHECK(ctx, m4d_backproject_bwd(gr.flat<float>().data(), in.flat<float>().data(), co.flat<float>().data(), dim,
                                          ig->flat<float>().data(), cg->flat<float>().data(), stream_of(ctx)));
  }
};

REGISTER_KERNEL_BUILDER(Name("BackProject").Device(DEVICE_GPU), BackProjectM4d);
REGISTER_KERNEL_BUILDER(Name("BackProjectGrad").Device(DEVICE_GPU), BackProjectGradM4d);

// ------------------------------------------------------------------------------------------------ fused PSCV
// (cv, prev_disp) = get_parallax_sweeping_cv(c1, c2, disp_prev_t, disp, rot, trans, camera, search_range, nbre_cuts)
REGISTER_OP("M4dPscvFused")
    .Input("c1: float32")            // [b,h,w,c]
    .Input("c2: float32")            // [b,h,w,c]
    .Input("disp_prev_t: float32")   // [b,h,w,1]
    .Input("disp: float32")          // [b,h,w,1]
    .Input("rot: float32")           // [b,4] (w,x,y,z) or [b,3]
    .Input("trans: float32")         // [b,3]
    .Input("cam_f: float32")         // [b,2]
    .Input("cam_c: float32")         // [b,2]
    .Attr("search_range: int = 4")
    .Attr("nbre_cuts: int = 1")
    .Attr("interp: int = 0")         // M4D_INTERP_GATHER (what the reference's TF-CPU path computes) / _BP / _BP_FMA
    .Output("cv: float32")           // [b,h,w,nbre_cuts*(2r+1)]
    .Output("prev_disp: float32")    // [b,h,w,2r+1]
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      shape_inference::ShapeHandle f;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 4, &f));
      int r, cuts;
      TF_RETURN_IF_ERROR(c->GetAttr("search_range", &r));
      TF_RETURN_IF_ERROR(c->GetAttr("nbre_cuts", &cuts));
      c->set_output(0, c->MakeShape({c->Dim(f, 0), c->Dim(f, 1), c->Dim(f, 2), cuts * (2 * r + 1)}));
      c->set_output(1, c->MakeShape({c->Dim(f, 0), c->Dim(f, 1), c->Dim(f, 2), 2 * r + 1}));
      return Status();
    });

class M4dPscvFused : public OpKernel {
 public:
  explicit M4dPscvFused(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("search_range", &r_));
    OP_REQUIRES_OK(c, c->GetAttr("nbre_cuts", &cuts_));
    OP_REQUIRES_OK(c, c->GetAttr("interp", &interp_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& c1 = ctx->input(0);
    OP_REQUIRES(ctx, c1.dims() == 4 && ctx->input(1).shape() == c1.shape(), errors::InvalidArgument("M4dPscvFused: c1/c2 must be [b,h,w,c]"));
    const int b = c1.dim_size(0), h = c1.dim_size(1), w = c1.dim_size(2), c = c1.dim_size(3), K = 2 * r_ + 1;
    const Tensor& rot = ctx->input(4);
    OP_REQUIRES(ctx, rot.dims() == 2 && rot.dim_size(0) == b, errors::InvalidArgument("M4dPscvFused: rot must be [b,3|4]"));
    Tensor *cv = nullptr, *pd = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({b, h, w, cuts_ * K}), &cv));
    OP_REQUIRES_OK(ctx, ctx->allocate_output(1, TensorShape({b, h, w, K}), &pd));
    if (cv->NumElements() == 0) return;
    auto p = [&](int i) { return ctx->input(i).flat<float>().data(); };
    M4D_TF_CHECK(ctx, m4d_pscv_fused_fwd_ex(p(0), p(1), p(2), p(3), p(4), (int)rot.dim_size(1), p(5), p(6), p(7), b, h, w, c, cuts_, r_,
                                            cv->flat<float>().data(), cuts_ * K, pd->flat<float>().data(), K, nullptr, 0, 1.0f,
                                            nullptr, interp_, stream_of(ctx)));
  }

 private:
  int r_, cuts_, interp_;
};
REGISTER_KERNEL_BUILDER(Name("M4dPscvFused").Device(DEVICE_GPU), M4dPscvFused);

// Gradient op of M4dPscvFused (gather convention), to be registered from Python:
//   @ops.RegisterGradient("M4dPscvFused")
//   def _pscv_grad(op, d_cv, d_prev_disp):
//       g = _m4d_ops.m4d_pscv_fused_grad(*op.inputs, d_cv, d_prev_disp, search_range=op.get_attr("search_range"),
//                                        nbre_cuts=op.get_attr("nbre_cuts"))
//       return [g.d_c1, g.d_c2, g.d_disp_prev_t, g.d_disp, None, None, None, None]      # pose and camera are data
REGISTER_OP("M4dPscvFusedGrad")
    .Input("c1: float32")
    .Input("c2: float32")
    .Input("disp_prev_t: float32")
    .Input("disp: float32")
    .Input("rot: float32")
    .Input("trans: float32")
    .Input("cam_f: float32")
    .Input("cam_c: float32")
    .Input("d_cv: float32")
    .Input("d_prev_disp: float32")
    .Attr("search_range: int = 4")
    .Attr("nbre_cuts: int = 1")
    .Output("d_c1: float32")
    .Output("d_c2: float32")
    .Output("d_disp_prev_t: float32")
    .Output("d_disp: float32")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      for (int i = 0; i < 4; ++i) c->set_output(i, c->input(i));
      return Status();
    });

class M4dPscvFusedGrad : public OpKernel {
 public:
  explicit M4dPscvFusedGrad(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("search_range", &r_));
    OP_REQUIRES_OK(c, c->GetAttr("nbre_cuts", &cuts_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& c1 = ctx->input(0);
    OP_REQUIRES(ctx, c1.dims() == 4 && ctx->input(1).shape() == c1.shape(), errors::InvalidArgument("M4dPscvFusedGrad: c1/c2 must be [b,h,w,c]"));
    const int b = c1.dim_size(0), h = c1.dim_size(1), w = c1.dim_size(2), c = c1.dim_size(3), K = 2 * r_ + 1;
    const Tensor& rot = ctx->input(4);
    Tensor* out[4];
    for (int i = 0; i < 4; ++i) OP_REQUIRES_OK(ctx, ctx->allocate_output(i, ctx->input(i).shape(), &out[i]));
    if (c1.NumElements() == 0) return;
    auto p = [&](int i) { return ctx->input(i).flat<float>().data(); };
    M4D_TF_CHECK(ctx, m4d_pscv_fused_bwd(p(0), p(1), p(2), p(3), p(4), (int)rot.dim_size(1), p(5), p(6), p(7), b, h, w, c, cuts_, r_, p(8),
                                         cuts_ * K, p(9), K, out[0]->flat<float>().data(), out[1]->flat<float>().data(),
                                         out[2]->flat<float>().data(), out[3]->flat<float>().data(), stream_of(ctx)));
  }

 private:
  int r_, cuts_;
};
REGISTER_KERNEL_BUILDER(Name("M4dPscvFusedGrad").Device(DEVICE_GPU), M4dPscvFusedGrad);

// ------------------------------------------------------------------------------------------------ SNCV
// out = cost_volume(c1, c2, search_range, nbre_cuts=nbre_cuts)   (leaky_relu(0.1) included, utils/depth_operations.py:311)
REGISTER_OP("M4dSncv")
    .Input("c1: float32")
    .Input("c2: float32")
    .Attr("search_range: int = 3")
    .Attr("nbre_cuts: int = 1")
    .Output("out: float32")
    .SetShapeFn([](shape_inference::InferenceContext* c) {
      shape_inference::ShapeHandle f;
      TF_RETURN_IF_ERROR(c->WithRank(c->input(0), 4, &f));
      int r, cuts;
      TF_RETURN_IF_ERROR(c->GetAttr("search_range", &r));
      TF_RETURN_IF_ERROR(c->GetAttr("nbre_cuts", &cuts));
      c->set_output(0, c->MakeShape({c->Dim(f, 0), c->Dim(f, 1), c->Dim(f, 2), cuts * (2 * r + 1) * (2 * r + 1)}));
      return Status();
    });

class M4dSncv : public OpKernel {
 public:
  explicit M4dSncv(OpKernelConstruction* c) : OpKernel(c) {
    OP_REQUIRES_OK(c, c->GetAttr("search_range", &r_));
    OP_REQUIRES_OK(c, c->GetAttr("nbre_cuts", &cuts_));
  }
  void Compute(OpKernelContext* ctx) override {
    const Tensor& c1 = ctx->input(0);
    OP_REQUIRES(ctx, c1.dims() == 4 && ctx->input(1).shape() == c1.shape(), errors::InvalidArgument("M4dSncv: c1/c2 must be [b,h,w,c]"));
    const int b = c1.dim_size(0), h = c1.dim_size(1), w = c1.dim_size(2), c = c1.dim_size(3), n = 2 * r_ + 1;
    Tensor* out = nullptr;
    OP_REQUIRES_OK(ctx, ctx->allocate_output(0, TensorShape({b, h, w, cuts_ * n * n}), &out));
    if (out->NumElements() == 0) return;
    M4D_TF_CHECK(ctx, m4d_sncv_fwd(c1.flat<float>().data(), ctx->input(1).flat<float>().data(), b, h, w, c, cuts_, r_,
                                   out->flat<float>().data(), cuts_ * n * n, stream_of(ctx)));
  }

 private:
  int r_, cuts_;
};
REGISTER_KERNEL_BUILDER(Name("M4dSncv").Device(DEVICE_GPU), M4dSncv);
