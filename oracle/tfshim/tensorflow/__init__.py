"""Minimal ``tensorflow`` stand-in over torch (CPU), TEST INFRASTRUCTURE ONLY.

Purpose: let the reference's own modules -- /root/reference/tensorBNN/
{BNN_functions,layer,activationFunctions,likelihood,network,paramAdapter,
metrics}.py -- be imported and executed UNMODIFIED in this container, where
TensorFlow itself cannot be installed, so that known-answer fixtures for the
parity tests come from the reference's source instead of from a restatement
(tests/golden/make_ref_golden.py).  Only the ~70 ``tf.*`` symbols those files
touch exist here, each with TensorFlow's documented eager semantics
(broadcasting, reduce-over-all-axes defaults, ``tf.where`` select,
``tf.transpose`` reversing all axes, ``tf.cast`` accepting numpy or tf dtypes,
``tf.while_loop`` as a plain loop, ``tf.function`` as the identity).
Tensors are torch CPU tensors, so reverse-mode gradients of the reference's
closures come from torch autograd (the reference gets them from TF autodiff
inside TFP's leapfrog).

Never imported by the product (tensorbnn_b200) or at GPU-box run time.
"""
import math as _math
import sys as _sys
import types as _types

import numpy as _np
import torch as _torch

_torch.set_grad_enabled(True)

# ---------------------------------------------------------------------------
# dtypes
# ---------------------------------------------------------------------------
float16 = _torch.float16
float32 = _torch.float32
float64 = _torch.float64
int32 = _torch.int32
int64 = _torch.int64
bool = _torch.bool  # noqa: A001  (tf.bool)

_NP2T = {"float16": float16, "float32": float32, "float64": float64, "int32": int32,
         "int64": int64, "bool": _torch.bool}


def _dt(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, _torch.dtype):
        return dtype
    return _NP2T[_np.dtype(dtype).name]


class Tensor(_torch.Tensor):
    """torch tensor with the two eager-TF conveniences the reference uses:
    ``.numpy()`` on a tensor that carries a graph, and ``len()`` of a 1-D
    tensor (torch already has it)."""

    def numpy(self):  # tf.Tensor.numpy()
        return _torch.Tensor.numpy(self.detach().as_subclass(_torch.Tensor))

    def __format__(self, spec):
        return format(self.detach().as_subclass(_torch.Tensor).item(), spec)


def _wrap(t):
    return t.as_subclass(Tensor) if isinstance(t, _torch.Tensor) and not isinstance(t, Tensor) else t


def convert_to_tensor(value, dtype=None):
    """tf.convert_to_tensor: nested python lists may hold tensors (stacked,
    gradient-preserving, as TF packs them); python floats become float32,
    python ints int32 (TF's defaults)."""
    dtype = _dt(dtype)
    if isinstance(value, _torch.Tensor):
        out = value if dtype is None or value.dtype == dtype else value.to(dtype)
        return _wrap(out)
    if isinstance(value, (list, tuple)):
        if len(value) == 0:
            return _wrap(_torch.zeros((0,), dtype=dtype or float32))
        if any(isinstance(v, (list, tuple, _torch.Tensor)) for v in value):
            parts = [convert_to_tensor(v, dtype) for v in value]
            if dtype is None:
                dt0 = parts[0].dtype
                for p in parts:
                    if p.dtype.is_floating_point:
                        dt0 = p.dtype
                        break
                parts = [p.to(dt0) for p in parts]
            return _wrap(_torch.stack([p.as_subclass(_torch.Tensor) for p in parts]))
    arr = _np.asarray(value)
    py = _python_leaves(value)
    if dtype is None:
        if arr.dtype == _np.float64 and py:
            dtype = float32          # python floats default to float32 in TF (numpy scalars keep their dtype)
        elif arr.dtype == _np.int64 and py:
            dtype = int32
        else:
            dtype = _dt(arr.dtype)
    if py and not dtype.is_floating_point and dtype != _torch.bool and arr.dtype.kind == "f":
        arr = _np.trunc(arr)
    if arr.dtype == _np.bool_ and dtype == _torch.bool:
        return _wrap(_torch.as_tensor(arr))
    return _wrap(_torch.as_tensor(arr.astype(_np.dtype(str(dtype).replace("torch.", "")))))


def _python_leaves(value):
    """True when every leaf of a (nested) python container is a plain python number."""
    if isinstance(value, (list, tuple)):
        return all(_python_leaves(v) for v in value)
    return type(value) in (int, float, type(True))


def _t(x, like=None):
    """operand coercion for binary ops: python scalars adopt the dtype of the
    tensor operand (TF's behaviour for python constants)."""
    if isinstance(x, _torch.Tensor):
        return x
    if like is not None and isinstance(x, (int, float)) and not isinstance(x, type(True)):
        return _torch.as_tensor(x, dtype=like.dtype)
    return convert_to_tensor(x)


def _pair(a, b):
    ta, tb = isinstance(a, _torch.Tensor), isinstance(b, _torch.Tensor)
    if ta and not tb:
        return a, _t(b, a)
    if tb and not ta:
        return _t(a, b), b
    return _t(a), _t(b)


# ---------------------------------------------------------------------------
# construction / casting / shape
# ---------------------------------------------------------------------------
def constant(value, dtype=None, shape=None):
    t = convert_to_tensor(value, dtype)
    if shape is not None:
        t = t.reshape(tuple(shape)) if t.numel() != 1 else t.expand(tuple(shape)).clone()
    return _wrap(t)


def Variable(value, dtype=None):
    return constant(value, dtype)


def cast(x, dtype):
    return _wrap(convert_to_tensor(x).to(_dt(dtype))) if not isinstance(x, _torch.Tensor) \
        else _wrap(x.to(_dt(dtype)))


def shape(input):  # noqa: A002
    return _wrap(_torch.tensor(list(_t(input).shape), dtype=int32))


def rank(x):
    return _wrap(_torch.tensor(_t(x).dim(), dtype=int32))


def size(input, out_type=int32):  # noqa: A002
    return _wrap(_torch.tensor(_t(input).numel(), dtype=_dt(out_type)))


def _shape_arg(s):
    if isinstance(s, _torch.Tensor):
        return tuple(int(v) for v in s.reshape(-1).tolist())
    if isinstance(s, (int, _np.integer)):
        return (int(s),)
    return tuple(int(v) for v in s)


def reshape(tensor, shape):  # noqa: A002
    return _wrap(_t(tensor).reshape(_shape_arg(shape)))


def pad(tensor, paddings, constant_values=0):
    t = _t(tensor)
    assert t.dim() == 1 and len(paddings) == 1
    lo, hi = (int(v) for v in paddings[0])
    return _wrap(_torch.nn.functional.pad(t, (lo, hi), value=constant_values))


def transpose(a, perm=None):
    t = _t(a)
    if perm is None:
        perm = tuple(reversed(range(t.dim())))
    return _wrap(t.permute(*perm)) if t.dim() else _wrap(t)


def squeeze(input, axis=None):  # noqa: A002
    t = _t(input)
    return _wrap(t.squeeze() if axis is None else t.squeeze(axis))


def expand_dims(input, axis):  # noqa: A002
    return _wrap(_t(input).unsqueeze(axis))


def concat(values, axis):
    parts = [_t(v) for v in values]
    dt0 = parts[0].dtype
    keep = []
    for p in parts:
        if p.dim() == 1 and p.numel() == 0 and parts[0].dim() == 2:
            # tf.concat([K, [[]]]) of an empty python row: shape [1,0]
            p = p.reshape(1, 0)
        keep.append(p.to(dt0))
    return _wrap(_torch.cat(keep, dim=axis))


def split(value, num_or_size_splits, axis=0):
    t = _t(value)
    n = int(num_or_size_splits)
    return [_wrap(c) for c in _torch.chunk(t, n, dim=axis)]


def ones(shape, dtype=float32):  # noqa: A002
    return _wrap(_torch.ones(_shape_arg(shape), dtype=_dt(dtype)))


def zeros(shape, dtype=float32):  # noqa: A002
    return _wrap(_torch.zeros(_shape_arg(shape), dtype=_dt(dtype)))


def ones_like(x):
    return _wrap(_torch.ones_like(_t(x)))


def eye(n, dtype=float32):
    return _wrap(_torch.eye(int(n), dtype=_dt(dtype)))


def linspace(start, stop, num):
    s, e = _t(start), _t(stop)
    dt0 = s.dtype if s.dtype.is_floating_point else float32
    n = int(num)
    # TF: start + delta * range(num) with delta = (stop-start)/(num-1), in the input dtype
    if n == 1:
        return _wrap(s.to(dt0).reshape(1))
    step = (e.to(dt0) - s.to(dt0)) / _torch.tensor(n - 1, dtype=dt0)
    idx = _torch.arange(n, dtype=dt0)
    out = s.to(dt0) + step * idx
    out[-1] = e.to(dt0)
    return _wrap(out)


# ---------------------------------------------------------------------------
# elementwise / reductions
# ---------------------------------------------------------------------------
def add(x, y):
    a, b = _pair(x, y)
    return _wrap(a + b)


def subtract(x, y):
    a, b = _pair(x, y)
    return _wrap(a - b)


def multiply(x, y):
    a, b = _pair(x, y)
    return _wrap(a * b)


def divide(x, y):
    a, b = _pair(x, y)
    return _wrap(a / b)


def maximum(x, y):
    # TF's _MaximumMinimumGrad routes the whole gradient to x where x >= y (torch.maximum would split ties)
    a, b = _pair(x, y)
    return _wrap(_torch.where(a >= b, a, b))


def minimum(x, y):
    a, b = _pair(x, y)
    return _wrap(_torch.where(a <= b, a, b))


def less(x, y):
    a, b = _pair(x, y)
    return _wrap(a < b)


def where(condition, x=None, y=None):
    c = _t(condition)
    if isinstance(condition, (type(True), _np.bool_)):
        c = _torch.tensor(condition)
    a, b = _pair(x, y)
    if a.dtype != b.dtype:
        dt0 = a.dtype if a.dtype.is_floating_point else b.dtype
        a, b = a.to(dt0), b.to(dt0)
    return _wrap(_torch.where(c, a, b))


def clip_by_value(t, clip_value_min, clip_value_max):
    x = _t(t)
    lo = _torch.as_tensor(clip_value_min, dtype=x.dtype)
    hi = _torch.as_tensor(clip_value_max, dtype=x.dtype)
    # tf.clip_by_value = maximum(minimum(t, hi), lo); the gradient passes on [lo, hi] inclusive, zero where clipped
    return maximum(minimum(x, hi), lo)


def exp(x):
    return _wrap(_torch.exp(_t(x)))


def abs(x):  # noqa: A001
    return _wrap(_torch.abs(_t(x)))


def square(x):
    t = _t(x)
    return _wrap(t * t)


def round(x):  # noqa: A001
    return _wrap(_torch.round(_t(x)))     # both round half to even


def matmul(a, b):
    x, y = _pair(a, b)
    return _wrap(x @ y)


def _axis(axis):
    return None if axis is None else axis


def reduce_sum(input_tensor, axis=None):
    t = _stack_if_list(input_tensor)
    return _wrap(t.sum() if axis is None else t.sum(dim=axis))


def _stack_if_list(v):
    if isinstance(v, _torch.Tensor):
        return v
    return convert_to_tensor(v)


def reduce_mean(input_tensor, axis=None):
    t = _stack_if_list(input_tensor)
    return _wrap(t.mean() if axis is None else t.mean(dim=axis))


def reduce_max(input_tensor, axis=None):
    t = _stack_if_list(input_tensor)
    return _wrap(t.max() if axis is None else t.max(dim=axis).values)


def reduce_std(input_tensor, axis=None):
    t = _stack_if_list(input_tensor)       # population standard deviation (ddof = 0)
    return _wrap(t.std(unbiased=False) if axis is None else t.std(dim=axis, unbiased=False))


def print(*args, **kwargs):  # noqa: A001  (tf.print)
    import builtins
    builtins.print(*args, **kwargs)


def function(func=None, **kwargs):
    """tf.function / tf.function(jit_compile=..., experimental_relax_shapes=...):
    tracing and XLA change scheduling, not arithmetic -- identity here."""
    if func is not None and callable(func):
        return func

    def deco(f):
        return f
    return deco


def while_loop(cond, body, loop_vars, **kwargs):
    vars_ = list(loop_vars)
    while _truth(cond(*vars_)):
        vars_ = list(body(*vars_))
    return vars_


def _truth(v):
    if isinstance(v, _torch.Tensor):
        return v.item() is True or (v.dtype != _torch.bool and v.item() != 0)
    return True if v else False


# ---------------------------------------------------------------------------
# sub-modules: tf.math, tf.linalg, tf.random, tf.nn, tf.errors
# ---------------------------------------------------------------------------
def _mod(name):
    m = _types.ModuleType(__name__ + "." + name)
    _sys.modules[m.__name__] = m
    return m


math = _mod("math")
math.log = lambda x: _wrap(_torch.log(_t(x)))
math.exp = exp
math.abs = abs
math.tanh = lambda x: _wrap(_torch.tanh(_t(x)))
math.sigmoid = lambda x: _wrap(_torch.sigmoid(_t(x)))
math.square = square
math.multiply = multiply
math.less = less
math.reduce_sum = reduce_sum
math.reduce_mean = reduce_mean
math.reduce_max = reduce_max
math.reduce_std = reduce_std
math.scalar_mul = lambda scalar, x: _wrap(_t(scalar, _t(x)) * _t(x))
math.log1p = lambda x: _wrap(_torch.log1p(_t(x)))
math.softplus = lambda x: _wrap(_torch.nn.functional.softplus(_t(x)))


def _squared_difference(x, y):
    a, b = _pair(x, y)
    return _wrap((a - b) * (a - b))


def _multiply_no_nan(x, y):
    a, b = _pair(x, y)
    return _wrap(_torch.where(b == 0, _torch.zeros_like(a * b), a * b))


math.squared_difference = _squared_difference
math.multiply_no_nan = _multiply_no_nan

linalg = _mod("linalg")


class _Errors(object):
    class InvalidArgumentError(Exception):
        pass


errors = _Errors()


def _inv(m):
    t = _t(m)
    try:
        return _wrap(_torch.linalg.inv(t))
    except Exception as e:  # singular input: TF raises InvalidArgumentError
        raise errors.InvalidArgumentError(str(e))


linalg.inv = _inv
linalg.diag = lambda d: _wrap(_torch.diag(convert_to_tensor(d)))

nn = _mod("nn")
nn.leaky_relu = lambda features, alpha=0.2: _wrap(_torch.where(_t(features) < 0, _t(alpha, _t(features)) * _t(features),
                                                               _t(features)))

random = _mod("random")
_gen = _torch.Generator().manual_seed(0)
# Injection hooks so fixture generation controls every draw (TF's own stateful
# Philox stream cannot be reproduced outside TF).
random.normal_hook = None
random.uniform_hook = None


def _set_seed(seed):
    _gen.manual_seed(int(seed))


def _normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None):
    shp = _shape_arg(shape)
    if random.normal_hook is not None:
        z = random.normal_hook(shp, _dt(dtype))
    else:
        g = _gen if seed is None else _torch.Generator().manual_seed(int(seed))
        z = _torch.randn(shp, dtype=_dt(dtype), generator=g)
    return _wrap(z * _t(stddev, z) + _t(mean, z))


def _uniform(shape, minval=0, maxval=1, dtype=float32, seed=None):
    shp = _shape_arg(shape)
    if random.uniform_hook is not None:
        u = random.uniform_hook(shp, _dt(dtype))
    else:
        u = _torch.rand(shp, dtype=_dt(dtype), generator=_gen)
    return _wrap(u * (maxval - minval) + minval)


random.set_seed = _set_seed
random.normal = _normal
random.uniform = _uniform

# ``from tensorflow.python.ops import gen_nn_ops`` (activationFunctions.py:4)
python = _mod("python")
python.ops = _mod("python.ops")
gen_nn_ops = _mod("python.ops.gen_nn_ops")
python.ops.gen_nn_ops = gen_nn_ops
gen_nn_ops.relu = lambda x: _wrap(_torch.relu(_t(x)))
gen_nn_ops.elu = lambda x: _wrap(_torch.nn.functional.elu(_t(x)))
gen_nn_ops.softmax = lambda x: _wrap(_torch.softmax(_t(x), dim=-1))


class _Keras(object):
    """tf.keras is used only by the Keras pre-training helpers
    (BNN_functions.py:60-298), which are out of scope."""

    def __getattr__(self, name):
        raise NotImplementedError("tf.keras is not part of the shim")


keras = _Keras()
