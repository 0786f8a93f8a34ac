"""Oracle: the reference's OWN BackProject CUDA kernels, compiled unmodified (test infrastructure, see oracle/__init__.py).

``oracle/Makefile`` compiles ``/root/reference/cuda_backproject/backproject_op_gpu.cu.cc`` where it lies, against two stand-in
TensorFlow headers (``oracle/ref_stub``: the file only uses ``Eigen::GpuDevice::stream()`` / ``ok()``), together with
``oracle/ref_wrap.cu`` (``extern "C"`` doors onto ``BackProjectForwardLauncher`` / ``BackProjectBackwardLauncher``,
backproject_op_gpu.h:17-22) into ``oracle/_ref/libbackproject_ref.so``.  The library is built in the build container (which has
the reference) and travels to the GPU box with the snapshot; this module only loads it.

Used by the GPU parity tests as the oracle of ``m4d_backproject_fwd`` / ``m4d_backproject_bwd`` and of the BP_FMA warp inside
the fused PSCV, and by ``bench.py`` as the GPU kernel the fused path replaces.  Needs a CUDA device (the kernels are CUDA).
"""
import ctypes as C
import os

import torch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "libbackproject_ref.so")
_lib = None


def available():
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise FileNotFoundError(f"{LIB_PATH} is missing: run `make -C oracle` in a container that has /root/reference")
        _lib = C.CDLL(LIB_PATH)
        _lib.ref_backproject_fwd.restype = C.c_int
        _lib.ref_backproject_fwd.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        _lib.ref_backproject_bwd.restype = C.c_int
        _lib.ref_backproject_bwd.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


def _dims(inputs, coords):
    B, H, W, Fd, Cc = inputs.shape
    S = coords.shape[3]
    assert tuple(coords.shape) == (B, H, W, S, Fd, 2), (inputs.shape, coords.shape)
    return (C.c_int * 6)(B, H, W, S, Fd, Cc), (B, H, W, S, Fd, Cc)


def back_project(inputs, coords, keep_on_device=False):
    """inputs [B,H,W,F,C], coords [B,H,W,S,F,2] (x,y) -> [B,H,W,S,F,C] through the reference's compiled kernel.  CPU tensors
    are moved to cuda:0 and the result is returned on the CPU (so this can stand in for ``oracle.back_project``)."""
    dev_in = inputs.is_cuda
    i_d, c_d = inputs.cuda().float().contiguous(), coords.cuda().float().contiguous()
    dim, (B, H, W, S, Fd, Cc) = _dims(i_d, c_d)
    out = torch.empty(B, H, W, S, Fd, Cc, dtype=torch.float32, device=i_d.device)
    torch.cuda.synchronize()                        # the reference memsets on the legacy default stream
    rc = lib().ref_backproject_fwd(i_d.data_ptr(), c_d.data_ptr(), dim, out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0, "reference launcher reported a CUDA error"
    return out if (dev_in or keep_on_device) else out.cpu()


def back_project_grad(inputs, coords, grad):
    """-> (inputs_grad [B,H,W,F,C], coords_grad [B,H,W,S,F,2]) through the reference's compiled gradient kernel."""
    dev_in = inputs.is_cuda
    i_d, c_d, g_d = (t.cuda().float().contiguous() for t in (inputs, coords, grad))
    dim, (B, H, W, S, Fd, Cc) = _dims(i_d, c_d)
    gi = torch.empty_like(i_d)
    gc = torch.empty_like(c_d)
    torch.cuda.synchronize()
    rc = lib().ref_backproject_bwd(g_d.data_ptr(), i_d.data_ptr(), c_d.data_ptr(), dim, gi.data_ptr(), gc.data_ptr(),
                                   torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert rc == 0, "reference launcher reported a CUDA error"
    return (gi, gc) if dev_in else (gi.cpu(), gc.cpu())
