"""A tiny stand-in for the TensorFlow 2.7 API surface that M4Depth's inference path touches.

WHY: TensorFlow is not installable in the build image (no network), so the reference
(`/root/reference`) cannot be executed as shipped.  With this package on ``sys.path`` the reference's
*own unmodified* ``utils/depth_operations.py``, ``utils/dense_image_warp.py`` and
``m4depth_network.py`` import and run, so golden vectors can be produced from the reference's real
Python op sequence (``tools/gen_golden.py``).  Only the TF *primitives* are restated here, each on
torch-CPU fp32 with one rounded op per TF op:

* elementwise math, reshape/stack/concat/tile/transpose/pad/slice/split/gather/meshgrid/range
* ``a @ b`` / ``tf.linalg.matmul``: inner-dimension sum taken left to right (TF leaves it unspecified)
* ``tf.reduce_mean`` over float16: fp16-rounded inputs summed in fp32 in index order, divided by n,
  rounded once to fp16 (TF leaves the order unspecified; this is the repo's "fp32acc" contract)
* Keras ``Conv2D(3, padding='same')`` with TF SAME padding, ``Layer``/``Model`` build-on-first-call,
  ``add_weight``/``Variable.assign``
* ``tf.compat.v1.image.resize_bilinear`` (legacy, no half-pixel), ``tf.image.resize(NEAREST)``
* ``tf.linalg.normalize``, ``tf.math.l2_normalize``, ``tf.math.reduce_variance``

This is test tooling; it is never imported by the product package.  It does not implement autograd,
graphs, devices, or anything the training path needs.
"""
import builtins as _builtins
import math as _math
import sys as _sys
import types as _types

import numpy as _np
import torch as _torch

__version__ = "2.7.0-shim"
_builtin_range = _builtins.range


# ------------------------------------------------------------------------------------------ dtypes
class DType:
    def __init__(self, name, tdt):
        self.name, self.t = name, tdt

    def __eq__(self, o):
        return as_dtype(o).t == self.t

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "tf." + self.name


float16 = half = DType("float16", _torch.float16)
float32 = DType("float32", _torch.float32)
float64 = DType("float64", _torch.float64)
int32 = DType("int32", _torch.int32)
int64 = DType("int64", _torch.int64)
bool_ = DType("bool", _torch.bool)
_BY_T = {d.t: d for d in (float16, float32, float64, int32, int64, bool_)}
_BY_N = {d.name: d for d in (float16, float32, float64, int32, int64, bool_)}


def as_dtype(d):
    if isinstance(d, DType):
        return d
    if isinstance(d, str):
        return _BY_N[d]
    if isinstance(d, _torch.dtype):
        return _BY_T[d]
    raise TypeError(d)


# ------------------------------------------------------------------------------------------ tensors
class TensorShape(list):
    def as_list(self):
        return list(self)

    def __getitem__(self, k):
        r = list.__getitem__(self, k)
        return TensorShape(r) if isinstance(k, _builtins.slice) else r

    def __add__(self, o):
        return TensorShape(list(self) + list(o))


def _raw(x, like=None):
    """python / numpy / Tensor -> torch tensor or python scalar."""
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, (int, float, bool)):
        return x
    if isinstance(x, _torch.Tensor):
        return x
    if isinstance(x, _np.ndarray):
        return _torch.from_numpy(x)
    if isinstance(x, (list, tuple)):
        if any(isinstance(e, (Tensor, _torch.Tensor)) for e in x):
            return _torch.stack([_torch.as_tensor(_raw(e)) for e in x])
        t = _torch.tensor(x)
        if t.dtype == _torch.float64:
            t = t.to(_torch.float32)
        if like is not None and t.is_floating_point() == like.is_floating_point():
            t = t.to(like.dtype)
        return t
    raise TypeError(type(x))


def _ints(seq):
    """shape-like (list with 0-d tensors / Tensor) -> list[int]."""
    if isinstance(seq, Tensor):
        return [int(v) for v in seq.t.reshape(-1).tolist()]
    return [int(_raw(v)) for v in seq]


class Tensor:
    __array_priority__ = 100

    def __init__(self, t):
        self.t = t if isinstance(t, _torch.Tensor) else _torch.as_tensor(t)

    # structure
    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    @property
    def dtype(self):
        return _BY_T[self.t.dtype]

    def numpy(self):
        return self.t.numpy()

    def __len__(self):
        return self.t.shape[0]

    def __bool__(self):
        return bool(self.t)

    def __int__(self):
        return int(self.t)

    def __float__(self):
        return float(self.t)

    def __index__(self):
        return int(self.t)

    def __iter__(self):
        return (Tensor(v) for v in self.t)

    def __getitem__(self, k):
        def conv(e):
            return int(e) if isinstance(e, Tensor) else e
        k = tuple(conv(e) for e in k) if isinstance(k, tuple) else conv(k)
        return Tensor(self.t[k])

    def __repr__(self):
        return "shimTensor(%r)" % (self.t,)

    # arithmetic (one rounded torch op each)
    def _b(self, o, f, rev=False):
        o = _raw(o, like=self.t)
        return Tensor(f(o, self.t) if rev else f(self.t, o))

    def __add__(self, o): return self._b(o, _torch.add)
    def __radd__(self, o): return self._b(o, _torch.add, True)
    def __sub__(self, o): return self._b(o, _torch.sub)
    def __rsub__(self, o): return self._b(o, lambda a, b: _torch.sub(_torch.as_tensor(a, dtype=b.dtype) if not isinstance(a, _torch.Tensor) else a, b), True)
    def __mul__(self, o): return self._b(o, _torch.mul)
    def __rmul__(self, o): return self._b(o, _torch.mul, True)
    def __truediv__(self, o): return self._b(o, _torch.true_divide)
    def __rtruediv__(self, o): return self._b(o, lambda a, b: _torch.true_divide(_torch.as_tensor(a, dtype=b.dtype) if not isinstance(a, _torch.Tensor) else a, b), True)
    def __neg__(self): return Tensor(-self.t)
    def __matmul__(self, o): return matmul(self, o)

    def __pow__(self, p):
        if p == 2:
            return Tensor(self.t * self.t)          # graph mode lowers x**2 to Square
        return Tensor(_torch.pow(self.t, p))

    def __lt__(self, o): return self._b(o, _torch.lt)
    def __le__(self, o): return self._b(o, _torch.le)
    def __gt__(self, o): return self._b(o, _torch.gt)
    def __ge__(self, o): return self._b(o, _torch.ge)


class Variable(Tensor):
    def __init__(self, initial_value=None, trainable=True, name=None, dtype=None, **kw):
        t = _raw(initial_value)
        t = _torch.as_tensor(t)
        if dtype is not None:
            t = t.to(as_dtype(dtype).t)
        super().__init__(t.clone())
        self.name, self.trainable = name, trainable

    def assign(self, v):
        self.t = _torch.as_tensor(_raw(v)).to(self.t.dtype).clone().reshape(self.t.shape)
        return self

    def assign_add(self, v):
        self.t = self.t + _raw(v)
        return self


def convert_to_tensor(x, dtype=None, name=None):
    t = x if isinstance(x, Tensor) else Tensor(_torch.as_tensor(_raw(x)))
    return cast(t, dtype) if dtype is not None else t


def constant(v, dtype=None, name=None):
    t = _torch.as_tensor(v)
    if dtype is not None:
        t = t.to(as_dtype(dtype).t)
    elif t.dtype == _torch.float64:
        t = t.to(_torch.float32)
    return Tensor(t)


def function(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


class _Scope:
    def __init__(self, *a, **k): pass
    def __enter__(self): return self
    def __exit__(self, *a): return False


name_scope = _Scope


def identity(x, name=None): return x
def stop_gradient(x, name=None): return x


def cast(x, dtype, name=None):
    x = convert_to_tensor(x)
    return Tensor(x.t.to(as_dtype(dtype).t))


def _dt(dtype, default=float32):
    return as_dtype(dtype).t if dtype is not None else default.t


def ones(shape, dtype=None, name=None): return Tensor(_torch.ones(_ints(shape), dtype=_dt(dtype)))
def zeros(shape, dtype=None, name=None): return Tensor(_torch.zeros(_ints(shape), dtype=_dt(dtype)))
def ones_initializer(): return lambda shape, dtype=None: ones(shape, dtype)
def zeros_initializer(): return lambda shape, dtype=None: zeros(shape, dtype)


def shape(input=None, name=None, **kw):
    return Tensor(_torch.tensor(list(convert_to_tensor(input).t.shape), dtype=_torch.int32))


def reshape(tensor, shape, name=None):
    return Tensor(convert_to_tensor(tensor).t.reshape(_ints(shape)))


def stack(values, axis=0, name=None):
    ts = [_torch.as_tensor(_raw(v)) for v in values]
    dt = next((t.dtype for t in ts if t.is_floating_point()), ts[0].dtype)
    return Tensor(_torch.stack([t.to(dt) for t in ts], dim=axis))


def unstack(value, axis=0, num=None, name=None):
    return [Tensor(t) for t in convert_to_tensor(value).t.unbind(axis)]


def concat(values, axis, name=None):
    ts = [_torch.as_tensor(_raw(v)) for v in values]
    if ts[0].dim() == 1 and not ts[0].is_floating_point():
        ts = [t.to(_torch.int32) for t in ts]
    return Tensor(_torch.cat(ts, dim=axis))


def expand_dims(input, axis, name=None): return Tensor(convert_to_tensor(input).t.unsqueeze(axis))
def squeeze(input, axis=None, name=None):
    t = convert_to_tensor(input).t
    return Tensor(t.squeeze() if axis is None else t.squeeze(axis))


def tile(input, multiples, name=None): return Tensor(convert_to_tensor(input).t.repeat(*_ints(multiples)))
def transpose(a, perm=None, name=None): return Tensor(convert_to_tensor(a).t.permute(*perm).contiguous())
def reverse(tensor, axis, name=None): return Tensor(_torch.flip(convert_to_tensor(tensor).t, dims=list(axis)))


def range(start, limit=None, delta=1, dtype=None, name=None):   # noqa: A001
    if limit is None:
        start, limit = 0, start
    s, l, d = (_raw(v) for v in (start, limit, delta))
    s, l, d = (v.item() if isinstance(v, _torch.Tensor) else v for v in (s, l, d))
    if dtype is None:
        dtype = float32 if any(isinstance(v, float) for v in (s, l, d)) else int32
    n = max(int(_math.ceil((l - s) / d)), 0)
    # TF evaluates start + i*delta per element
    return Tensor((_torch.arange(n, dtype=_torch.float64) * d + s).to(as_dtype(dtype).t))


def meshgrid(*args, indexing='xy'):
    return [Tensor(g.contiguous()) for g in _torch.meshgrid(*[convert_to_tensor(a).t for a in args], indexing=indexing)]


def pad(tensor, paddings, mode="CONSTANT", constant_values=0, name=None):
    t = convert_to_tensor(tensor).t
    flat = []
    for lo, hi in reversed([tuple(p) for p in paddings]):
        flat += [int(lo), int(hi)]
    return Tensor(_torch.nn.functional.pad(t, flat, value=constant_values))


def slice(input_, begin, size, name=None):   # noqa: A001
    t = convert_to_tensor(input_).t
    idx = tuple(_builtins.slice(b, None if s == -1 else b + s) for b, s in zip(_ints(begin), _ints(size)))
    return Tensor(t[idx])


def split(value, num_or_size_splits, axis=0, num=None, name=None):
    t = convert_to_tensor(value).t
    if isinstance(num_or_size_splits, int):
        return [Tensor(c) for c in _torch.chunk(t, num_or_size_splits, dim=axis)]
    return [Tensor(c) for c in _torch.split(t, list(num_or_size_splits), dim=axis)]


def gather(params, indices, axis=0, name=None, **kw):
    p, i = convert_to_tensor(params).t, convert_to_tensor(indices).t.to(_torch.int64)
    return Tensor(p[i]) if axis in (0, None) else Tensor(_torch.index_select(p, axis, i.reshape(-1)).reshape(
        *p.shape[:axis], *i.shape, *p.shape[axis + 1:]))


def _un(f):
    return lambda x, name=None: Tensor(f(convert_to_tensor(x).t))


def _ieee_sqrt(t):
    # torch.sqrt (MKL VML) is not correctly rounded in fp32; TF's Eigen sqrt is.  numpy's sqrt is the hardware one.
    import numpy as _np
    return _torch.from_numpy(_np.sqrt(t.detach().contiguous().numpy())).reshape(t.shape)


sqrt, exp, log, floor, ceil, abs, square = (_un(f) for f in (   # noqa: A001
    _ieee_sqrt, _torch.exp, _torch.log, _torch.floor, _torch.ceil, _torch.abs, lambda t: t * t))
rsqrt = _un(_torch.rsqrt)


def _bi(f):
    def g(x, y, name=None):
        x = convert_to_tensor(x)
        return Tensor(f(x.t, _torch.as_tensor(_raw(y, like=x.t), dtype=x.t.dtype) if not isinstance(y, Tensor) else y.t))
    return g


multiply, divide, add, subtract = _bi(_torch.mul), _bi(_torch.true_divide), _bi(_torch.add), _bi(_torch.sub)
maximum, minimum = _bi(_torch.maximum), _bi(_torch.minimum)
greater, less = _bi(_torch.gt), _bi(_torch.lt)


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    return minimum(maximum(t, clip_value_min), clip_value_max)


def _axes(axis, nd):
    if axis is None:
        return list(_builtin_range(nd))
    return [axis] if isinstance(axis, int) else list(axis)


def reduce_sum(input_tensor, axis=None, keepdims=False, name=None):
    t = convert_to_tensor(input_tensor).t
    return Tensor(t.sum(dim=_axes(axis, t.dim()), keepdim=keepdims))


def reduce_mean(input_tensor, axis=None, keepdims=False, name=None):
    t = convert_to_tensor(input_tensor).t
    ax = _axes(axis, t.dim())
    if t.dtype == _torch.float16:
        # contract "fp32acc": ordered fp32 sum of the fp16 values, /n, one rounding to fp16
        assert len(ax) == 1
        a = ax[0]
        acc = _torch.zeros_like(t.select(a, 0), dtype=_torch.float32)
        for j in _builtin_range(t.shape[a]):
            acc = acc + t.select(a, j).to(_torch.float32)
        m = (acc / float(t.shape[a])).to(_torch.float16)
        return Tensor(m.unsqueeze(a) if keepdims else m)
    return Tensor(t.mean(dim=ax, keepdim=keepdims))



def matmul(a, b, name=None, **kw):
    """[..., m, k] @ [..., k, n] with the k-sum taken left to right, products rounded separately."""
    a, b = convert_to_tensor(a).t, convert_to_tensor(b).t
    acc = None
    for k in _builtin_range(a.shape[-1]):
        term = a[..., :, k:k + 1] * b[..., k:k + 1, :]
        acc = term if acc is None else acc + term
    return Tensor(acc)


def norm(tensor, ord='euclidean', axis=None, keepdims=None, name=None):
    t = convert_to_tensor(tensor).t
    return Tensor(_ieee_sqrt((t * t).sum(dim=_axes(axis, t.dim()), keepdim=bool(keepdims))))


def load_op_library(path):
    raise NotImplementedError("tf_shim cannot load TF op libraries")


# ------------------------------------------------------------------------------------------ image
def _resize_bilinear_legacy(images, size, align_corners=False, half_pixel_centers=False, name=None):
    assert not align_corners and not half_pixel_centers
    x = convert_to_tensor(images).t
    oh, ow = _ints(size)
    b, h, w, c = x.shape

    def axis(n_in, n_out):
        scale = _torch.tensor(n_in / n_out, dtype=_torch.float32)
        src = _torch.arange(n_out, dtype=_torch.float32) * scale
        lo_f = _torch.floor(src)
        return (lo_f.to(_torch.int64).clamp(min=0), _torch.ceil(src).to(_torch.int64).clamp(max=n_in - 1), src - lo_f)

    ylo, yhi, ly = axis(h, oh)
    xlo, xhi, lx = axis(w, ow)
    lx, ly = lx.view(1, 1, ow, 1), ly.view(1, oh, 1, 1)
    tl, tr = x[:, ylo][:, :, xlo], x[:, ylo][:, :, xhi]
    bl, br = x[:, yhi][:, :, xlo], x[:, yhi][:, :, xhi]
    top = tl + (tr - tl) * lx
    bot = bl + (br - bl) * lx
    return Tensor(top + (bot - top) * ly)


class _ResizeMethod:
    NEAREST_NEIGHBOR = "nearest"
    BILINEAR = "bilinear"


def _resize(images, size, method="bilinear", **kw):
    x = convert_to_tensor(images).t
    oh, ow = _ints(size)
    if method != "nearest":
        raise NotImplementedError("tf_shim: only NEAREST tf.image.resize is on the inference path")
    b, h, w, c = x.shape

    def axis(n_in, n_out):
        scale = _torch.tensor(n_in / n_out, dtype=_torch.float32)
        return _torch.floor((_torch.arange(n_out, dtype=_torch.float32) + 0.5) * scale).to(_torch.int64).clamp(max=n_in - 1)

    return Tensor(x[:, axis(h, oh)][:, :, axis(w, ow)])


def _mod(name, **attrs):
    m = _types.ModuleType(name)
    m.__dict__.update(attrs)
    _sys.modules[name] = m
    return m


image = _mod("tensorflow.image", resize=_resize, ResizeMethod=_ResizeMethod, resize_bilinear=_resize_bilinear_legacy)


def _leaky_relu(features, alpha=0.2, name=None):
    t = convert_to_tensor(features).t
    return Tensor(_torch.where(t >= 0, t, t * alpha))


nn = _mod("tensorflow.nn", leaky_relu=_leaky_relu)


def _l2_normalize(x, axis=None, epsilon=1e-12, name=None):
    t = convert_to_tensor(x).t
    sq = (t * t).sum(dim=_axes(axis, t.dim()), keepdim=True)
    return Tensor(t * _torch.rsqrt(_torch.clamp(sq, min=epsilon)))


def _reduce_variance(input_tensor, axis=None, keepdims=False, name=None):
    t = convert_to_tensor(input_tensor).t
    ax = _axes(axis, t.dim())
    dev = t - t.mean(dim=ax, keepdim=True)
    return Tensor((dev * dev).mean(dim=ax, keepdim=keepdims))


def _multiply_no_nan(x, y, name=None):
    x, y = convert_to_tensor(x).t, convert_to_tensor(y).t
    return Tensor(_torch.where(y == 0, _torch.zeros_like(x * y), x * y))


math = _mod("tensorflow.math", reduce_mean=reduce_mean, reduce_variance=_reduce_variance, l2_normalize=_l2_normalize,
            log=log, exp=exp, sqrt=sqrt, abs=abs, multiply_no_nan=_multiply_no_nan, maximum=maximum, minimum=minimum,
            less=less, greater=greater, reduce_sum=reduce_sum,
            squared_difference=lambda a, b: Tensor((_raw(a) - _raw(b)) * (_raw(a) - _raw(b))))


def _linalg_normalize(tensor, ord='euclidean', axis=None, name=None):
    n = norm(tensor, ord, axis, keepdims=True)
    return Tensor(convert_to_tensor(tensor).t / n.t), n


linalg = _mod("tensorflow.linalg", matmul=matmul, normalize=_linalg_normalize)

_v1_image = _mod("tensorflow.compat.v1.image", resize_bilinear=_resize_bilinear_legacy)
_v1 = _mod("tensorflow.compat.v1", name_scope=_Scope, image=_v1_image)
compat = _mod("tensorflow.compat", v1=_v1)
summary = _mod("tensorflow.summary", image=lambda *a, **k: None, scalar=lambda *a, **k: None)
errors = _mod("tensorflow.errors")


# ------------------------------------------------------------------------------------------ keras
_init_gen = _torch.Generator().manual_seed(7)


def set_initializer_seed(seed):
    _init_gen.manual_seed(seed)


class _HeNormal:
    def __call__(self, shape, dtype=None):
        shape = _ints(shape)
        fan_in = int(_np.prod(shape[:-1]))
        return Tensor(_torch.randn(shape, generator=_init_gen, dtype=_torch.float32) * _math.sqrt(2.0 / fan_in))


class _Reg:
    def __init__(self, *a, **k): pass
    def __call__(self, x): return Tensor(_torch.zeros(()))


class Layer:
    def __init__(self, trainable=True, name=None, **kw):
        self.trainable, self.built = trainable, False

    def add_weight(self, name=None, shape=None, dtype=None, initializer=None, trainable=True, **kw):
        init = initializer if initializer is not None else zeros_initializer()
        return Variable(init(shape, dtype or 'float32'), trainable=trainable, name=name)

    def add_loss(self, *a, **k): pass

    def build(self, input_shape): pass

    def __call__(self, *args, **kwargs):
        if not self.built:
            first = args[0]
            self.build(TensorShape(first.shape) if isinstance(first, Tensor) else None)
            self.built = True
        return self.call(*args, **kwargs)


class Conv2D(Layer):
    def __init__(self, filters, kernel_size, strides=(1, 1), padding='valid', kernel_initializer=None,
                 kernel_regularizer=None, **kw):
        super().__init__()
        assert kernel_size == 3 and padding == 'same'
        self.filters, self.stride = filters, strides[0]
        self.kinit = kernel_initializer or _HeNormal()

    def build(self, input_shape):
        self.kernel = Variable(self.kinit([3, 3, input_shape[-1], self.filters], 'float32'), name="kernel")
        self.bias = Variable(_torch.zeros(self.filters), name="bias")

    def call(self, x):
        t = x.t
        b, h, w, c = t.shape
        s = self.stride

        def padding(n):
            out = -(-n // s)
            tot = max((out - 1) * s + 3 - n, 0)
            return tot // 2, tot - tot // 2
        (pt, pb), (pl, pr) = padding(h), padding(w)
        xn = _torch.nn.functional.pad(t.permute(0, 3, 1, 2), (pl, pr, pt, pb))
        y = _torch.nn.functional.conv2d(xn, self.kernel.t.permute(3, 2, 0, 1).contiguous(), self.bias.t, stride=s)
        return Tensor(y.permute(0, 2, 3, 1).contiguous())


class Model(Layer):
    pass


class _Mean:
    def __init__(self, name=None, **kw): self.name = name


_layers = _mod("tensorflow.keras.layers", Layer=Layer, Conv2D=Conv2D)
_models = _mod("tensorflow.keras.models", Model=Model)
_inits = _mod("tensorflow.keras.initializers", HeNormal=_HeNormal)
_regs = _mod("tensorflow.keras.regularizers", L1=_Reg, L2=_Reg)
_metrics = _mod("tensorflow.keras.metrics", Mean=_Mean)
keras = _mod("tensorflow.keras", layers=_layers, models=_models, initializers=_inits, regularizers=_regs,
             metrics=_metrics, Model=Model)

# ------------------------------------------------------------------- tensorflow.python.* (dense_image_warp.py:25-31)
_this = _sys.modules[__name__]


class _OpsNS:
    name_scope = _Scope
    convert_to_tensor = staticmethod(convert_to_tensor)
    control_dependencies = staticmethod(lambda deps: _Scope())

    @staticmethod
    def RegisterGradient(name):
        return lambda f: f


_ops_mod = _mod("tensorflow.python.framework.ops", **{k: getattr(_OpsNS, k) for k in
                                                     ("name_scope", "convert_to_tensor", "control_dependencies", "RegisterGradient")})
_dtypes_mod = _mod("tensorflow.python.framework.dtypes", int32=int32, float32=float32, float16=float16)
_const_mod = _mod("tensorflow.python.framework.constant_op", constant=constant)
_fw = _mod("tensorflow.python.framework", ops=_ops_mod, dtypes=_dtypes_mod, constant_op=_const_mod)
_array_ops = _mod("tensorflow.python.ops.array_ops", shape=shape, unstack=unstack, expand_dims=expand_dims,
                  reshape=reshape, gather=gather, meshgrid=meshgrid, stack=stack)
_math_ops = _mod("tensorflow.python.ops.math_ops", cast=cast, minimum=minimum, maximum=maximum, floor=floor, range=range)
_check_ops = _mod("tensorflow.python.ops.check_ops")
_pyops = _mod("tensorflow.python.ops", array_ops=_array_ops, math_ops=_math_ops, check_ops=_check_ops)
python = _mod("tensorflow.python", framework=_fw, ops=_pyops)
