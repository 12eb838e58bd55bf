"""TEST INFRASTRUCTURE ONLY - a stand-in for the few dozen TensorFlow 2 eager ops that gpbasics 2.0.0 touches, written
on torch CPU float64, so that the UNMODIFIED reference sources under /root/reference can be imported and executed in
the build container (TensorFlow itself is not installable there: no network, not in /opt/wheelhouse).

It exists for one purpose: tests/golden/make_golden.py runs the reference's own Python (its op order, its
hyper-parameter slicing, its index bookkeeping) on top of this module and freezes the outputs as golden vectors.
Only the leaf numerical ops (matmul, exp, sin, cholesky, triangular_solve, ...) are supplied here, by their torch
equivalents of identical IEEE-754 float64 meaning; reverse-mode differentiation (tf.GradientTape) is torch.autograd.

Nothing in the product package or in the GPU tests imports this module.
"""
import builtins as _builtins
import contextlib
import types

import numpy as _np
import torch as _torch

__version__ = "2.99.0-shim"


class DType:
    def __init__(self, name, t):
        self.name = name
        self.torch = t

    def __repr__(self):
        return "tf." + self.name


float64 = DType("float64", _torch.float64)
float32 = DType("float32", _torch.float32)
int32 = DType("int32", _torch.int32)
int64 = DType("int64", _torch.int64)
bool = DType("bool", _torch.bool)  # noqa: A001
_BY_TORCH = {d.torch: d for d in (float64, float32, int32, int64, bool)}


def _td(dtype):
    if dtype is None:
        return None
    if isinstance(dtype, DType):
        return dtype.torch
    if isinstance(dtype, _torch.dtype):
        return dtype
    return _torch.from_numpy(_np.zeros(1, dtype=dtype)).dtype


class TensorShape:
    def __init__(self, dims):
        self._d = [int(v) for v in dims]

    def __eq__(self, other):
        if isinstance(other, TensorShape):
            return self._d == other._d
        try:
            return self._d == [int(v) for v in other]
        except TypeError:
            return False

    def __ne__(self, other):
        return not self.__eq__(other)

    def __len__(self):
        return len(self._d)

    def __getitem__(self, i):
        r = self._d[i]
        return TensorShape(r) if isinstance(i, _builtins.slice) else r

    def __iter__(self):
        return iter(self._d)

    def as_list(self):
        return list(self._d)

    def __repr__(self):
        return "TensorShape(%r)" % (self._d,)


def _raw(x, dtype=None):
    """anything -> torch tensor"""
    if isinstance(x, Tensor):
        t = x._t
    elif isinstance(x, _torch.Tensor):
        t = x
    elif isinstance(x, TensorShape):
        t = _torch.tensor(x.as_list())
    elif isinstance(x, (list, tuple)) and any(isinstance(v, (Tensor, _torch.Tensor)) for v in _flatten(x)):
        t = _stack_nested(x)
    else:
        a = _np.asarray(x)
        if a.dtype == _np.float32 or (a.dtype.kind == "f" and dtype is None):
            a = a.astype(_np.float64) if isinstance(x, (float, list, tuple)) or a.dtype == _np.float64 else a
        if a.ndim:
            t = _torch.from_numpy(_np.ascontiguousarray(a))
        else:  # python / numpy scalar: keep full double precision (torch.tensor(float) would round to float32)
            t = _torch.tensor(a.item(), dtype=_torch.float64 if a.dtype.kind == "f" else
                              (_torch.bool if a.dtype.kind == "b" else _torch.int64))
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t


def _flatten(x):
    for v in x:
        if isinstance(v, (list, tuple)):
            yield from _flatten(v)
        else:
            yield v


def _stack_nested(x):
    if isinstance(x, (list, tuple)):
        return _torch.stack([_stack_nested(v) for v in x])
    return _raw(x)


def _promote(a, b):
    """binary-op operands: python / numpy scalars take the tensor operand's dtype (TF semantics)"""
    ta = a._t if isinstance(a, Tensor) else None
    tb = b._t if isinstance(b, Tensor) else None
    if ta is None and tb is None:
        return _raw(a), _raw(b)
    if ta is None:
        ta = _raw(a, tb.dtype) if not isinstance(a, _torch.Tensor) else a
    if tb is None:
        tb = _raw(b, ta.dtype) if not isinstance(b, _torch.Tensor) else b
    return ta, tb


class Tensor:
    __array_ufunc__ = None
    __array_priority__ = 1000

    def __init__(self, t):
        self._t = t

    # -- introspection
    @property
    def shape(self):
        return TensorShape(self._t.shape)

    @property
    def dtype(self):
        return _BY_TORCH.get(self._t.dtype, DType(str(self._t.dtype), self._t.dtype))

    def numpy(self):
        return self._t.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    def __len__(self):
        return self._t.shape[0]

    def __iter__(self):
        for i in _builtins.range(self._t.shape[0]):
            yield Tensor(self._t[i])

    def __float__(self):
        return float(self._t)

    def __int__(self):
        return int(self._t)

    def __index__(self):
        return int(self._t)

    def __bool__(self):
        return builtins_bool(self._t)

    def __repr__(self):
        return "<tf_shim.Tensor shape=%s dtype=%s numpy=%r>" % (list(self._t.shape), self.dtype.name, self.numpy())

    def __copy__(self):
        return Tensor(self._t.clone())

    def __deepcopy__(self, memo):
        return type(self)(self._t.detach().clone()) if not isinstance(self, Variable) else Variable(self._t.detach().clone())

    def get_shape(self):
        return self.shape

    def __getitem__(self, idx):
        def conv(i):
            if isinstance(i, Tensor):
                return i._t
            return i
        if isinstance(idx, tuple):
            idx = tuple(conv(i) for i in idx)
        else:
            idx = conv(idx)
        return Tensor(self._t[idx])

    # -- arithmetic
    def _bin(self, other, fn, rev=False):
        a, b = _promote(self, other)
        return Tensor(fn(b, a) if rev else fn(a, b))

    def __add__(self, o): return self._bin(o, _torch.add)
    def __radd__(self, o): return self._bin(o, _torch.add, True)
    def __sub__(self, o): return self._bin(o, _torch.sub)
    def __rsub__(self, o): return self._bin(o, _torch.sub, True)
    def __mul__(self, o): return self._bin(o, _torch.mul)
    def __rmul__(self, o): return self._bin(o, _torch.mul, True)
    def __truediv__(self, o): return self._bin(o, _torch.true_divide)
    def __rtruediv__(self, o): return self._bin(o, _torch.true_divide, True)
    def __pow__(self, o): return self._bin(o, _torch.pow)
    def __neg__(self): return Tensor(-self._t)
    def __abs__(self): return Tensor(_torch.abs(self._t))
    def __matmul__(self, o): return self._bin(o, _torch.matmul)
    def __lt__(self, o): return self._bin(o, _torch.lt)
    def __le__(self, o): return self._bin(o, _torch.le)
    def __gt__(self, o): return self._bin(o, _torch.gt)
    def __ge__(self, o): return self._bin(o, _torch.ge)
    __hash__ = None

    def __eq__(self, o):
        if o is None:
            return False
        return self._bin(o, _torch.eq)

    def __ne__(self, o):
        if o is None:
            return True
        return self._bin(o, _torch.ne)


builtins_bool = _builtins.bool


class Variable(Tensor):
    """tf.Variable: a differentiable leaf"""

    def __init__(self, initial_value, dtype=None, shape=None, trainable=True, name=None):
        t = _raw(initial_value, _td(dtype)).detach().clone()
        if dtype is None and not t.dtype.is_floating_point and isinstance(initial_value, float):
            t = t.to(_torch.float64)
        if shape is not None:
            t = t.reshape(list(shape)) if t.numel() == int(_np.prod(list(shape), dtype=int)) else t.expand(list(shape)).clone()
        if t.dtype.is_floating_point:
            t.requires_grad_(True)
        super().__init__(t)

    def assign(self, value):
        with _torch.no_grad():
            self._t.copy_(_raw(value, self._t.dtype))
        return self

    def assign_sub(self, value):
        with _torch.no_grad():
            self._t.sub_(_raw(value, self._t.dtype))
        return self

    def assign_add(self, value):
        with _torch.no_grad():
            self._t.add_(_raw(value, self._t.dtype))
        return self

    def __hash__(self):
        return id(self)


def _w(t):
    return Tensor(t)


# ---- construction ----------------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name=None):
    t = _raw(value, _td(dtype))
    if isinstance(value, Tensor):
        t = t  # tf.constant(tensor) keeps the graph in eager mode
    if shape is not None:
        shape = list(shape)
        t = t.reshape(shape) if t.numel() == int(_np.prod(shape, dtype=int)) else t.expand(shape).clone()
    return _w(t)


def convert_to_tensor(value, dtype=None):
    return constant(value, dtype)


def cast(x, dtype, name=None):
    return _w(_raw(x).to(_td(dtype)))


def zeros(shape, dtype=float32, name=None):
    return _w(_torch.zeros([int(s) for s in shape], dtype=_td(dtype)))


def ones(shape, dtype=float32, name=None):
    return _w(_torch.ones([int(s) for s in shape], dtype=_td(dtype)))


def zeros_like(x, dtype=None):
    return _w(_torch.zeros_like(_raw(x), dtype=_td(dtype)))


def fill(dims, value, name=None, dtype=None):
    v = _raw(value, _td(dtype))
    return _w(v.expand([int(d) for d in dims]).clone() if v.dim() == 0 else v.reshape([int(d) for d in dims]))


def eye(num_rows, num_columns=None, dtype=float32, name=None):
    return _w(_torch.eye(int(num_rows), int(num_columns) if num_columns is not None else int(num_rows), dtype=_td(dtype)))


def linspace(start, stop, num, name=None):
    return _w(_torch.linspace(float(start), float(stop), int(num), dtype=_torch.float64))


def range(start, limit=None, delta=1, dtype=None):  # noqa: A001
    if limit is None:
        start, limit = 0, start
    return _w(_torch.arange(int(start), int(limit), int(delta), dtype=_td(dtype) or _torch.int32))


# ---- elementwise -----------------------------------------------------------------------------------------------
def _un(fn):
    def op(x, name=None):
        return _w(fn(_raw(x)))
    return op


def _bi(fn):
    def op(x, y, name=None):
        a, b = _promote(x if isinstance(x, Tensor) else (_w(_raw(x)) if not isinstance(x, (int, float)) else x),
                        y if isinstance(y, Tensor) else (_w(_raw(y)) if not isinstance(y, (int, float)) else y))
        return _w(fn(a, b))
    return op


sqrt = _un(_torch.sqrt)
square = _un(lambda t: t * t)
abs = _un(_torch.abs)  # noqa: A001
exp = _un(_torch.exp)
sin = _un(_torch.sin)
add = _bi(_torch.add)
subtract = _bi(_torch.sub)
multiply = _bi(_torch.mul)
divide = _bi(_torch.true_divide)
pow = _bi(_torch.pow)  # noqa: A001
less = _bi(_torch.lt)
equal = _bi(_torch.eq)
logical_and = _bi(_torch.logical_and)


def add_n(inputs, name=None):
    res = _raw(inputs[0])
    for t in inputs[1:]:
        res = res + _raw(t)
    return _w(res)


def where(condition, x=None, y=None, name=None):
    c = _raw(condition)
    if x is None:
        return _w(_torch.nonzero(c).to(_torch.int64))
    a, b = _promote(_w(_raw(x)), _w(_raw(y)))
    return _w(_torch.where(c, a, b))


# ---- reductions ------------------------------------------------------------------------------------------------
def _red(fn_all, fn_axis):
    def op(x, axis=None, keepdims=False, name=None):
        t = _raw(x)
        if axis is None:
            r = fn_all(t)
            if keepdims:
                r = r.reshape([1] * t.dim())
            return _w(r)
        return _w(fn_axis(t, int(axis), keepdims))
    return op


reduce_sum = _red(_torch.sum, lambda t, a, k: _torch.sum(t, dim=a, keepdim=k))
reduce_mean = _red(_torch.mean, lambda t, a, k: _torch.mean(t, dim=a, keepdim=k))
reduce_min = _red(_torch.min, lambda t, a, k: _torch.min(t, dim=a, keepdim=k).values)
reduce_max = _red(_torch.max, lambda t, a, k: _torch.max(t, dim=a, keepdim=k).values)
reduce_any = _red(_torch.any, lambda t, a, k: _torch.any(t, dim=a, keepdim=k))
reduce_all = _red(_torch.all, lambda t, a, k: _torch.all(t, dim=a, keepdim=k))


# ---- shape manipulation ----------------------------------------------------------------------------------------
def reshape(tensor, shape, name=None):
    if isinstance(shape, Tensor):
        shape = shape.numpy().tolist()
    if isinstance(shape, TensorShape):
        shape = shape.as_list()
    return _w(_raw(tensor).reshape([int(s) for s in shape]))


def shape(input, name=None):  # noqa: A002
    return _w(_torch.tensor(list(_raw(input).shape), dtype=_torch.int32))


def transpose(a, perm=None, name=None):
    t = _raw(a)
    if perm is None:
        perm = list(_builtins.range(t.dim()))[::-1]
    return _w(t.permute([int(p) for p in perm]))


def expand_dims(input, axis, name=None):  # noqa: A002
    return _w(_raw(input).unsqueeze(int(axis)))


def concat(values, axis, name=None):
    ts = [_raw(v) for v in values]
    dt = ts[0].dtype
    return _w(_torch.cat([t.to(dt) for t in ts], dim=int(axis)))


def gather(params, indices, axis=0, name=None):
    t = _raw(params)
    idx = _raw(indices).to(_torch.int64)
    if idx.dim() == 0:
        return _w(t.select(int(axis), int(idx)))
    flat = _torch.index_select(t, int(axis), idx.reshape(-1))
    new_shape = list(t.shape[:int(axis)]) + list(idx.shape) + list(t.shape[int(axis) + 1:])
    return _w(flat.reshape(new_shape))


def slice(input_, begin, size, name=None):  # noqa: A001
    t = _raw(input_)
    idx = tuple(_builtins.slice(int(b), None if int(s) < 0 else int(b) + int(s)) for b, s in zip(begin, size))
    return _w(t[idx])


def split(value, num_or_size_splits, axis=0, name=None):
    t = _raw(value)
    if isinstance(num_or_size_splits, int):
        return [_w(p) for p in _torch.chunk(t, num_or_size_splits, dim=int(axis))]
    return [_w(p) for p in _torch.split(t, [int(s) for s in num_or_size_splits], dim=int(axis))]


def sort(values, axis=-1, direction="ASCENDING", name=None):
    return _w(_torch.sort(_raw(values), dim=int(axis), descending=(direction != "ASCENDING")).values)


def tile(input, multiples, name=None):  # noqa: A002
    return _w(_raw(input).repeat([int(m) for m in multiples]))


def repeat(input, repeats, axis=None, name=None):  # noqa: A002
    t = _raw(input)
    if axis is None:
        return _w(t.reshape(-1).repeat_interleave(int(repeats)))
    return _w(t.repeat_interleave(int(repeats), dim=int(axis)))


def unique(x, name=None):
    v, idx = _torch.unique(_raw(x), sorted=False, return_inverse=True)
    return _w(v), _w(idx)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    ta, tb = _promote(a if isinstance(a, Tensor) else _w(_raw(a)), b if isinstance(b, Tensor) else _w(_raw(b)))
    if transpose_a:
        ta = ta.transpose(-1, -2)
    if transpose_b:
        tb = tb.transpose(-1, -2)
    return _w(_torch.matmul(ta, tb))


def tensordot(a, b, axes, name=None):
    return _w(_torch.tensordot(_raw(a), _raw(b), dims=axes))


def map_fn(fn, elems, **kw):
    return _w(_torch.stack([_raw(fn(_w(e))) for e in _raw(elems)]))


# ---- misc ------------------------------------------------------------------------------------------------------
@contextlib.contextmanager
def name_scope(name):
    yield


def function(func=None, **kw):
    if func is None:
        return lambda f: f
    return func


def custom_gradient(f):
    return f


def gradients(*a, **k):
    raise NotImplementedError("graph-mode tf.gradients is not used on the eager path")


class GradientTape:
    def __init__(self, persistent=False, watch_accessed_variables=True):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def watch(self, tensor):
        for t in (tensor if isinstance(tensor, (list, tuple)) else [tensor]):
            if isinstance(t, Tensor) and t._t.is_leaf and t._t.dtype.is_floating_point and not t._t.requires_grad:
                t._t.requires_grad_(True)

    def gradient(self, target, sources):
        single = not isinstance(sources, (list, tuple))
        srcs = [sources] if single else list(sources)
        gs = _torch.autograd.grad(_raw(target).sum(), [s._t for s in srcs], allow_unused=True, retain_graph=True)
        out = [None if g is None else _w(g) for g in gs]
        return out[0] if single else out


# ---- namespaces ------------------------------------------------------------------------------------------------
math = types.SimpleNamespace(
    add=add, subtract=subtract, multiply=multiply, divide=divide, exp=exp, log=_un(_torch.log), sin=sin, abs=abs,
    square=square, sqrt=sqrt, tanh=_un(_torch.tanh), sign=_un(_torch.sign), less=less, logical_and=logical_and,
    add_n=add_n, is_nan=_un(_torch.isnan), squared_difference=_bi(lambda a, b: (a - b) * (a - b)), pow=pow,
    reduce_sum=reduce_sum,
)
nn = types.SimpleNamespace(relu=_un(_torch.relu))


class _LinearOperatorFullMatrix:
    def __init__(self, matrix, **kw):
        self._m = matrix if isinstance(matrix, Tensor) else _w(_raw(matrix))

    def to_dense(self):
        return self._m

    def add_to_tensor(self, x, name=None):
        return self._m + x

    @property
    def shape(self):
        return self._m.shape


class _LinearOperatorBlockDiag:
    def __init__(self, operators, **kw):
        self.operators = list(operators)

    def to_dense(self):
        return _w(_torch.block_diag(*[_raw(o.to_dense()) for o in self.operators]))


def _cholesky(input, name=None):  # noqa: A002
    return _w(_torch.linalg.cholesky(_raw(input)))


def _triangular_solve(matrix, rhs, lower=True, adjoint=False, name=None):
    m = _raw(matrix)
    if adjoint:
        m = m.transpose(-1, -2)
        lower = not lower
    return _w(_torch.linalg.solve_triangular(m, _raw(rhs), upper=not lower))


linalg = types.SimpleNamespace(
    cholesky=_cholesky, triangular_solve=_triangular_solve,
    matrix_transpose=lambda a, name=None: _w(_raw(a).transpose(-1, -2)),
    diag_part=lambda a, name=None: _w(_torch.diagonal(_raw(a), dim1=-2, dim2=-1)),
    inv=lambda a, name=None: _w(_torch.linalg.inv(_raw(a))),
    pinv=lambda a, name=None: _w(_torch.linalg.pinv(_raw(a))),
    slogdet=lambda a, name=None: tuple(_w(v) for v in _torch.linalg.slogdet(_raw(a))),
    trace=lambda a, name=None: _w(_torch.diagonal(_raw(a), dim1=-2, dim2=-1).sum(-1)),
    eigvals=lambda a, name=None: _w(_torch.linalg.eigvals(_raw(a))),
    matmul=matmul,
    LinearOperatorFullMatrix=_LinearOperatorFullMatrix, LinearOperatorBlockDiag=_LinearOperatorBlockDiag,
)


def _gen(seed):
    g = _torch.Generator()
    g.manual_seed(int(seed) if seed is not None else 0)
    return g


_GLOBAL_GEN = _gen(1234)


def _rnormal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    g = _gen(seed) if seed is not None else _GLOBAL_GEN
    return _w(_torch.randn([int(s) for s in shape], dtype=_td(dtype), generator=g) * float(stddev) + float(mean))


def _runiform(shape, minval=0, maxval=None, dtype=float32, seed=None, name=None):
    g = _gen(seed) if seed is not None else _GLOBAL_GEN
    lo = float(minval)
    hi = float(maxval if maxval is not None else 1.0)
    dt = _td(dtype)
    if not dt.is_floating_point:
        return _w(_torch.randint(int(lo), int(hi), [int(s) for s in shape], dtype=dt, generator=g))
    return _w(_torch.rand([int(s) for s in shape], dtype=dt, generator=g) * (hi - lo) + lo)


def _rshuffle(value, seed=None, name=None):
    t = _raw(value)
    g = _gen(seed) if seed is not None else _GLOBAL_GEN
    return _w(t[_torch.randperm(t.shape[0], generator=g)])


random = types.SimpleNamespace(
    normal=_rnormal, uniform=_runiform, shuffle=_rshuffle,
    stateless_uniform=lambda shape, seed, minval=0, maxval=None, dtype=float32, name=None:
        _runiform(shape, minval, maxval, dtype, seed=int(_np.asarray(seed).reshape(-1)[0])),
    set_seed=lambda s: _GLOBAL_GEN.manual_seed(int(s)),
)

config = types.SimpleNamespace(threading=types.SimpleNamespace(
    set_inter_op_parallelism_threads=lambda n: None,
    set_intra_op_parallelism_threads=lambda n: _torch.set_num_threads(max(1, int(n)))))
