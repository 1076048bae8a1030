"""Eager stand-in for the TensorFlow-1 names the reference's loss / model code touches, on torch CPU float64.

TEST INFRASTRUCTURE ONLY.  TensorFlow 1.x cannot be installed in this image (no Python 3.12 build), so the
reference's own `src.losses.*`, `src.models.gaussian_process`, `src.utils.*` and `src.misc.sharded_fn` modules are
imported UNMODIFIED from the reference checkout with this package standing in for `tensorflow`: every `tf.*` call
is executed at once on torch tensors (wrapped in `Tensor`, which adds the few TF tensor attributes the reference reads:
`get_shape()`, `.shape.ndims`, `.dtype.as_numpy_dtype`, ...).  `tests/golden/make_golden_acq.py` uses it to write
golden vectors (acquisition values, posterior moments, gradients through torch autograd) that pin the oracle and the
CUDA path to the reference's code rather than to a restatement of it.  Semantics mirrored where TF-1 and torch could
differ: reduce_min / reduce_max split the gradient evenly among ties (torch.amin / amax do), relu'(0) = 0,
abs'(0) = 0, clip_by_value passes the gradient inside the range, cholesky's gradient is the symmetrised one.
Only what the myopic Monte-Carlo / closed-form path needs is implemented; anything else raises NotImplementedError.
"""
import contextlib
import sys
import types

import numpy as np
import torch

__version__ = "1.x-eager-shim"


# ---- dtypes -----------------------------------------------------------------------------------------------
class DType:
    def __init__(self, name, tdt, ndt):
        self.name, self.torch, self.as_numpy_dtype = name, tdt, ndt

    @property
    def base_dtype(self):
        return self

    def __eq__(self, other):
        try:
            return as_dtype(other).name == self.name
        except Exception:
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __hash__(self):
        return hash(self.name)

    def __repr__(self):
        return "tf." + self.name


float64 = DType("float64", torch.float64, np.float64)
float32 = DType("float32", torch.float32, np.float32)
int32 = DType("int32", torch.int32, np.int32)
int64 = DType("int64", torch.int64, np.int64)
bool = DType("bool", torch.bool, np.bool_)          # noqa: A001  (tf.bool)
_DTYPES = {d.name: d for d in (float64, float32, int32, int64, bool)}
_DTYPES["int"] = int64
_DTYPES["float"] = float64


def as_dtype(x):
    if isinstance(x, DType):
        return x
    if isinstance(x, str):
        return _DTYPES[x]
    if isinstance(x, torch.dtype):
        return next(d for d in _DTYPES.values() if d.torch == x)
    return _DTYPES[np.dtype(x).name]


# ---- static shapes ----------------------------------------------------------------------------------------
class Int(int):
    """Python int that also answers the tensor idioms the reference applies to shape entries."""

    @property
    def value(self):
        return int(self)

    @property
    def dtype(self):
        return int64

    def __getitem__(self, key):                       # num_shards[None]
        assert key is None
        return [self]


class TensorShape(tuple):
    def __new__(cls, dims=()):
        return super().__new__(cls, [None if d is None else Int(d) for d in dims])

    @property
    def ndims(self):
        return len(self)

    def as_list(self):
        return list(self)

    def concatenate(self, other):
        return TensorShape(tuple(self) + tuple(other))

    def __getitem__(self, key):
        out = tuple.__getitem__(self, key)
        return TensorShape(out) if isinstance(key, slice) else out


Dimension = Int


# ---- tensors ----------------------------------------------------------------------------------------------
def _raw(x):
    if isinstance(x, Tensor):
        return x.t
    if isinstance(x, np.ndarray):
        return torch.as_tensor(x)
    if isinstance(x, Int):
        return int(x)
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return torch.stack([torch.as_tensor(_raw(e)) for e in x])
    if isinstance(x, (list, tuple)) and x and all(isinstance(e, (int, np.integer)) for e in x):
        return [int(e) for e in x]                         # static shape vector
    return x


def _wrap(t):
    return Tensor(t) if isinstance(t, torch.Tensor) else t


class Tensor:
    __array_ufunc__ = None                            # numpy scalars defer to the reflected operators below
    __hash__ = object.__hash__                        # identity, like a TF-1 graph tensor

    def __init__(self, t):
        self.t = t

    @property
    def shape(self):
        return TensorShape(self.t.shape)

    def get_shape(self):
        return self.shape

    def set_shape(self, shape):
        return None

    @property
    def dtype(self):
        return as_dtype(self.t.dtype)

    @property
    def name(self):
        return getattr(self, "_name", "tensor")

    def __add__(self, o): return Tensor(self.t + _raw(o))
    def __radd__(self, o): return Tensor(_raw(o) + self.t)
    def __sub__(self, o): return Tensor(self.t - _raw(o))
    def __rsub__(self, o): return Tensor(_raw(o) - self.t)
    def __mul__(self, o): return Tensor(self.t * _raw(o))
    def __rmul__(self, o): return Tensor(_raw(o) * self.t)
    def __truediv__(self, o): return Tensor(self.t / _raw(o))
    def __rtruediv__(self, o): return Tensor(_raw(o) / self.t)
    def __pow__(self, o): return Tensor(self.t ** _raw(o))
    def __neg__(self): return Tensor(-self.t)
    def __getitem__(self, key): return Tensor(self.t[key])
    def __lt__(self, o): return Tensor(self.t < _raw(o))
    def __le__(self, o): return Tensor(self.t <= _raw(o))
    def __gt__(self, o): return Tensor(self.t > _raw(o))
    def __ge__(self, o): return Tensor(self.t >= _raw(o))
    def __bool__(self): return builtins_bool(self.t)
    def __int__(self): return int(self.t)
    def __index__(self): return int(self.t)
    def __float__(self): return float(self.t)
    def __len__(self): return self.t.shape[0]
    def __iter__(self): return (Tensor(t) for t in self.t)
    def __repr__(self): return "shim.Tensor(%r)" % (self.t,)

    def numpy(self):
        return self.t.detach().numpy()


class Variable(Tensor):
    def initialized_value(self):
        return self


import builtins as _builtins  # noqa: E402

builtins_bool = _builtins.bool


def _is_static(x):
    return isinstance(x, (int, float, np.integer, np.floating)) and not isinstance(x, builtins_bool)


# ---- scopes / bookkeeping ---------------------------------------------------------------------------------
@contextlib.contextmanager
def name_scope(name=None, *a, **k):
    yield name


@contextlib.contextmanager
def control_dependencies(deps=None):
    yield None


class _Scope:
    def __init__(self, name):
        self.name = name


@contextlib.contextmanager
def variable_scope(scope=None, reuse=None, **k):
    yield _Scope(scope if isinstance(scope, str) else getattr(scope, "name", None))


def get_variable(name=None, shape=None, dtype=None, initializer=None, trainable=True, **k):
    if initializer is None:
        raise NotImplementedError("shim get_variable needs an initializer")
    dt = as_dtype(dtype).torch if dtype is not None else None
    v = Variable(torch.as_tensor(_raw(initializer), dtype=dt).clone())
    v._name = name
    return v


def _noop_assert(*a, **k):
    return None


assert_non_negative = assert_equal = assert_positive = assert_rank = assert_less = assert_less_equal = _noop_assert
assert_greater_equal = Assert = _noop_assert


# ---- shapes and casts -------------------------------------------------------------------------------------
def shape(x, name=None):
    return [Int(s) for s in _raw(x).shape]


def rank(x):
    return Int(_raw(x).dim())


def size(x):
    return Int(_raw(x).numel())


def cast(x, dtype, name=None):
    dt = as_dtype(dtype)
    r = _raw(x)
    if isinstance(r, torch.Tensor):
        return Tensor(r.to(dt.torch))
    return Tensor(torch.as_tensor(r, dtype=dt.torch))


def constant(x, dtype=None, name=None):
    return cast(x, dtype) if dtype is not None else Tensor(torch.as_tensor(_raw(x)))


def identity(x, name=None):
    return x


def reshape(x, shp, name=None):
    return Tensor(_raw(x).reshape([int(s) for s in shp]))


def transpose(x, perm=None, name=None):
    r = _raw(x)
    return Tensor(r.permute(*perm) if perm is not None else r.permute(*reversed(range(r.dim()))))


def expand_dims(x, axis=None, name=None):
    return Tensor(_raw(x).unsqueeze(axis))


def squeeze(x, axis=None, name=None):
    return Tensor(_raw(x).squeeze() if axis is None else _raw(x).squeeze(axis))


def tile(x, multiples, name=None):
    return Tensor(_raw(x).repeat(*[int(m) for m in multiples]))


def concat(values, axis=0, name=None):
    if all(isinstance(v, (list, tuple)) and not any(isinstance(e, Tensor) for e in v) for v in values):
        out = []
        for v in values:
            out.extend(v)
        return out
    return Tensor(torch.cat([torch.as_tensor(_raw(v)) for v in values], dim=axis))


def stack(values, axis=0, name=None):
    return Tensor(torch.stack([torch.as_tensor(_raw(v)) for v in values], dim=axis))


def unstack(x, axis=0, name=None):
    if isinstance(x, (list, tuple)):
        return list(x)
    return [Tensor(t) for t in _raw(x).unbind(axis)]


def zeros(shp, dtype=float64, name=None):
    return Tensor(torch.zeros([int(s) for s in shp], dtype=as_dtype(dtype).torch))


def ones(shp, dtype=float64, name=None):
    return Tensor(torch.ones([int(s) for s in shp], dtype=as_dtype(dtype).torch))


def fill(dims, value, name=None):
    v = torch.as_tensor(_raw(value))
    return Tensor(torch.full([int(s) for s in dims], v.item(), dtype=v.dtype))


def zeros_like(x, name=None):
    return Tensor(torch.zeros_like(torch.as_tensor(_raw(x))))


def ones_like(x, name=None):
    return Tensor(torch.ones_like(torch.as_tensor(_raw(x))))


def eye(n, dtype=float64, name=None):
    return Tensor(torch.eye(int(n), dtype=as_dtype(dtype).torch))


# ---- elementwise ------------------------------------------------------------------------------------------
def _unary(fn):
    def op(x, name=None):
        return Tensor(fn(torch.as_tensor(_raw(x))))
    return op


exp = _unary(torch.exp)
log = _unary(torch.log)
sqrt = _unary(torch.sqrt)
abs = _unary(torch.abs)                                # noqa: A001
sign = _unary(torch.sign)
sigmoid = _unary(torch.sigmoid)
reciprocal = _unary(torch.reciprocal)
negative = _unary(torch.neg)
square = _unary(torch.square)
erfc = _unary(torch.special.erfc)
is_nan = _unary(torch.isnan)
ceil = _unary(torch.ceil)


def _binary(fn, static=None):
    def op(a, b, name=None):
        if static is not None and _is_static(a) and _is_static(b):
            return static(a, b)
        return Tensor(fn(torch.as_tensor(_raw(a)), torch.as_tensor(_raw(b))))
    return op


add = _binary(torch.add)
subtract = _binary(torch.sub)
multiply = _binary(torch.mul)
divide = _binary(torch.div)
div = _binary(torch.div, static=lambda a, b: Int(a // b) if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) else a / b)
floor_div = _binary(torch.floor_divide, static=lambda a, b: Int(a // b))
minimum = _binary(torch.minimum, static=lambda a, b: Int(min(a, b)) if isinstance(a, (int, np.integer)) and isinstance(b, (int, np.integer)) else min(a, b))
maximum = _binary(torch.maximum, static=lambda a, b: max(a, b))
squared_difference = _binary(lambda a, b: (a - b) ** 2)
less = _binary(torch.lt, static=lambda a, b: a < b)
less_equal = _binary(torch.le, static=lambda a, b: a <= b)
greater = _binary(torch.gt, static=lambda a, b: a > b)
greater_equal = _binary(torch.ge, static=lambda a, b: a >= b)
equal = _binary(torch.eq, static=lambda a, b: a == b)
logical_and = _binary(torch.logical_and, static=lambda a, b: a and b)
logical_or = _binary(torch.logical_or, static=lambda a, b: a or b)


def clip_by_value(x, lo, hi, name=None):
    return Tensor(torch.clamp(_raw(x), min=lo, max=hi))


def where(cond, x=None, y=None, name=None):
    return Tensor(torch.where(_raw(cond), torch.as_tensor(_raw(x)), torch.as_tensor(_raw(y))))


def cond(pred, true_fn, false_fn, name=None):
    return true_fn() if builtins_bool(_raw(pred)) else false_fn()


def add_n(values, name=None):
    out = _raw(values[0])
    for v in values[1:]:
        out = out + _raw(v)
    return Tensor(out)


# ---- reductions -------------------------------------------------------------------------------------------
def _axes(r, axis):
    if axis is None:
        return list(range(r.dim()))
    return [axis] if isinstance(axis, (int, np.integer)) else list(axis)


def reduce_sum(x, axis=None, keepdims=False, name=None):
    r = _raw(x)
    return Tensor(r.sum(dim=_axes(r, axis), keepdim=keepdims))


def reduce_mean(x, axis=None, keepdims=False, name=None):
    r = _raw(x)
    return Tensor(r.mean(dim=_axes(r, axis), keepdim=keepdims))


def reduce_min(x, axis=None, keepdims=False, name=None):          # gradient split evenly among ties, as in TF
    r = _raw(x)
    return Tensor(r.amin(dim=_axes(r, axis), keepdim=keepdims))


def reduce_max(x, axis=None, keepdims=False, name=None):
    r = _raw(x)
    return Tensor(r.amax(dim=_axes(r, axis), keepdim=keepdims))


def reduce_prod(x, axis=None, keepdims=False, name=None):
    if isinstance(x, (list, tuple)):
        return Int(int(np.prod(x)))
    r = _raw(x)
    for a in sorted(_axes(r, axis), reverse=True):
        r = r.prod(dim=a, keepdim=keepdims)
    return Tensor(r)


# ---- linear algebra ---------------------------------------------------------------------------------------
def cholesky(a, name=None):
    return Tensor(torch.linalg.cholesky(_raw(a)))


def cholesky_solve(chol, rhs, name=None):
    return Tensor(torch.cholesky_solve(_raw(rhs), _raw(chol)))


def matrix_inverse(a, name=None):
    return Tensor(torch.linalg.inv(_raw(a)))


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    a, b = _raw(a), _raw(b)
    if transpose_a:
        a = a.transpose(-1, -2)
    if transpose_b:
        b = b.transpose(-1, -2)
    return Tensor(a @ b)


def einsum(equation, *operands, name=None):
    return Tensor(torch.einsum(equation, *[_raw(o) for o in operands]))


def tensordot(a, b, axes, name=None):
    a, b = torch.as_tensor(_raw(a)), torch.as_tensor(_raw(b))
    if isinstance(axes, int):
        return Tensor(torch.tensordot(a, b, dims=axes))
    ax_a = [x % a.dim() for x in axes[0]]
    ax_b = [x % b.dim() for x in axes[1]]
    return Tensor(torch.tensordot(a, b, dims=(ax_a, ax_b)))


def diag_part(a, name=None):
    return Tensor(torch.diagonal(_raw(a)))


class _NN:
    @staticmethod
    def relu(x, name=None):
        return Tensor(torch.relu(_raw(x)))


nn = _NN()


# ---- control flow -----------------------------------------------------------------------------------------
def _map_structure(fn, *structs):
    s0 = structs[0]
    if isinstance(s0, (tuple, list)):
        return s0.__class__([_map_structure(fn, *[s[i] for s in structs]) for i in range(len(s0))])
    return fn(*structs)


def map_fn(fn, elems, dtype=None, **k):
    """tf.map_fn over the leading axis: run fn on every slice at once and stack the (nested) results."""
    leaves = []
    _map_structure(lambda e: leaves.append(e), elems)
    count = _raw(leaves[0]).shape[0]
    outs = [fn(_map_structure(lambda e, i=i: Tensor(_raw(e)[i]), elems)) for i in range(count)]
    return _map_structure(lambda *parts: Tensor(torch.stack([_raw(p) for p in parts], dim=0)), *outs)


def random_normal(shp, dtype=float64, **k):
    raise NotImplementedError("the golden vectors always pass base samples explicitly")


Tensor.__module__ = __name__


# ---- everything else: importable, not callable --------------------------------------------------------------
class _Missing:
    def __init__(self, name):
        self._name = name

    def __call__(self, *a, **k):
        raise NotImplementedError("tensorflow shim: %s is not implemented" % self._name)

    def __getattr__(self, item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Missing(self._name + "." + item)


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    return _Missing("tf." + name)


def _stub_module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    m.__getattr__ = lambda item: _Missing(name + "." + item)
    sys.modules[name] = m
    return m


contrib = _stub_module(__name__ + ".contrib")
contrib.opt = _stub_module(__name__ + ".contrib.opt", ScipyOptimizerInterface=_Missing("ScipyOptimizerInterface"))
python = _stub_module(__name__ + ".python")
python.client = _stub_module(__name__ + ".python.client", timeline=_Missing("timeline"))
