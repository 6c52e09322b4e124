"""A torch-backed, lazily evaluated stand-in for the ~45 TensorFlow-1.x symbols the reference's TF half uses.
TEST INFRASTRUCTURE ONLY (oracle/make_golden.py puts this directory on sys.path as `tensorflow`).

Purpose: TensorFlow 1.x cannot be installed here, yet the graph-building code of the reference -
policies/networks/mlp.py, policies/{base,gaussian_mlp_policy,meta_gaussian_mlp_policy}.py,
policies/distributions/diagonal_gaussian.py, meta_algos/{base,pro_mp,trpo_maml,vpg_maml}.py,
optimizers/{maml_first_order_optimizer,conjugate_gradient_optimizer}.py, meta_trainer.py - is plain Python that only
*composes* TF ops.  This module lets those files run UNMODIFIED from /root/reference: every `tf.<op>` returns a lazy
graph node, `Session.run(fetches, feed_dict)` evaluates the requested nodes with torch (CPU) tensors, `tf.gradients`
differentiates with `torch.autograd.grad(create_graph=True)` (so gradients of gradients work exactly like TF's symbolic
ones), and `tf.train.AdamOptimizer` applies the published TF1 update rule
(tensorflow/python/training/adam.py `_apply_dense` / `training_ops.apply_adam`:
lr_t = lr*sqrt(1-b2^t)/(1-b1^t); m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; var -= lr_t m/(sqrt(v)+eps), with the
beta powers kept as float32 accumulators like TF's `beta1_power` / `beta2_power` variables).

What is NOT TensorFlow here: the floating-point kernels (torch's CPU matmul / tanh / exp instead of Eigen's) and the
random streams (`random_normal` and the initialisers draw from a module RandomState, see `set_random_seed` and
`set_random_normal_hook`).  Tie-breaking of `minimum` / `maximum` gradients follows TF (`x <= y` / `x >= y` masks).

`set_compute_dtype(torch.float64)` evaluates the very same graphs in double precision (placeholders / variables /
constants are promoted), which gives golden vectors free of float32 rounding.
"""
import contextlib
import re
import sys
import types

import numpy as np
import torch

sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))

__version__ = '1.x-oracle-stub'

# ------------------------------------------------------------------------------------------------ dtypes / state


class DType(object):
    def __init__(self, name, np_dtype, torch_dtype):
        self.name, self.as_numpy_dtype, self._torch = name, np_dtype, torch_dtype

    def __repr__(self):
        return 'tf.' + self.name

    @property
    def base_dtype(self):
        return self


float32 = DType('float32', np.float32, torch.float32)
float64 = DType('float64', np.float64, torch.float64)
int32 = DType('int32', np.int32, torch.int32)
int64 = DType('int64', np.int64, torch.int64)
bool_ = DType('bool', np.bool_, torch.bool)

_state = types.SimpleNamespace(
    compute_dtype=torch.float32,       # what tf.float32 placeholders / variables evaluate in
    scope=[],                          # variable_scope stack
    variables=[],                      # all Variables in creation order
    by_name={},                        # name -> Variable (for reuse=True)
    default_session=[],
    rng=np.random.RandomState(0),
    normal_hook=None,
)


def set_compute_dtype(torch_dtype):
    """Evaluate float32 graph nodes in this torch dtype (torch.float32 = faithful, torch.float64 = rounding-free)."""
    _state.compute_dtype = torch_dtype
    for v in _state.variables:
        if v.value is not None and v.value.is_floating_point():
            v.value = v.value.to(torch_dtype)


def set_random_seed(seed):
    _state.rng = np.random.RandomState(seed % 4294967294)


def set_random_normal_hook(fn):
    """fn(shape) -> ndarray replaces the module RandomState for tf.random_normal (lets golden scripts inject / record
    the action noise)."""
    _state.normal_hook = fn


def reset_default_graph():
    _state.scope, _state.variables, _state.by_name = [], [], {}


def _tdtype(dtype):
    if dtype is None:
        return _state.compute_dtype
    if isinstance(dtype, DType):
        return _state.compute_dtype if dtype._torch in (torch.float32, torch.float64) else dtype._torch
    if isinstance(dtype, torch.dtype):
        return dtype
    return _state.compute_dtype


class TensorShape(tuple):
    def as_list(self):
        return list(self)

    @property
    def ndims(self):
        return len(self)


def _shape(s):
    if s is None:
        return None
    return TensorShape(None if d is None else int(d) for d in s)


# ------------------------------------------------------------------------------------------------ graph nodes


class Tensor(object):
    """A lazy graph node: `fn(*evaluated_inputs)` -> torch tensor."""
    __array_priority__ = 1000

    def __init__(self, fn, inputs=(), shape=None, dtype=float32, name=None):
        self._fn, self.inputs = fn, tuple(inputs)
        self._shape, self.dtype = _shape(shape), dtype
        self.name = '/'.join(_state.scope + [name or 'op']) + ':0'

    @property
    def shape(self):
        return self._shape

    def get_shape(self):
        return self._shape

    @property
    def op(self):
        return self

    def __hash__(self):
        return id(self)

    def __repr__(self):
        return '<tf-stub Tensor %s shape=%s>' % (self.name, self._shape)

    # arithmetic (TF1 does not overload ==)
    def __add__(self, o): return add(self, o)
    def __radd__(self, o): return add(o, self)
    def __sub__(self, o): return subtract(self, o)
    def __rsub__(self, o): return subtract(o, self)
    def __mul__(self, o): return multiply(self, o)
    def __rmul__(self, o): return multiply(o, self)
    def __truediv__(self, o): return divide(self, o)
    def __rtruediv__(self, o): return divide(o, self)
    def __neg__(self): return _unary(torch.neg, self, 'neg')
    def __iter__(self): raise TypeError("Tensor objects are not iterable in graph mode")
    def __bool__(self): raise TypeError("using a tf.Tensor as a Python bool is not allowed")


class Variable(Tensor):
    def __init__(self, initial_value=None, name=None, dtype=None, trainable=True, _initializer=None, _shape=None):
        self._initializer = _initializer
        if initial_value is not None:
            arr = np.asarray(initial_value)
            self._initializer = lambda shape, a=arr: a
            _shape = arr.shape
        Tensor.__init__(self, None, (), _shape, dtype or float32, name or 'Variable')
        self.trainable = trainable
        self.value = None                 # torch tensor once initialised
        _state.variables.append(self)
        _state.by_name[self.name] = self

    def initialize(self):
        arr = np.asarray(self._initializer(tuple(self._shape)))
        self.value = torch.as_tensor(arr).to(_tdtype(self.dtype)).reshape(tuple(self._shape)).clone()

    def initialized_value(self):
        return self


class Operation(object):
    """A side-effect-only node (train op, initialiser)."""

    def __init__(self, run, name='op'):
        self._run, self.name = run, name


class _Const(Tensor):
    def __init__(self, value, like=None):
        arr = np.asarray(value)
        kind = float32 if arr.dtype.kind == 'f' else (int32 if arr.dtype.kind in 'iu' else bool_)
        Tensor.__init__(self, None, (), arr.shape, kind, 'Const')
        self._arr = arr


def _as_tensor(x):
    if isinstance(x, Tensor):
        return x
    return _Const(x)


def convert_to_tensor(x, dtype=None, name=None):
    return _as_tensor(x)


def constant(value, dtype=None, shape=None, name=None):
    return _Const(np.asarray(value) if shape is None else np.broadcast_to(np.asarray(value), shape))


# ------------------------------------------------------------------------------------------------ shape helpers


def _bshape(a, b):
    sa, sb = a.shape, b.shape
    if sa is None or sb is None:
        return None
    out = []
    for i in range(1, max(len(sa), len(sb)) + 1):
        da = sa[-i] if i <= len(sa) else 1
        db = sb[-i] if i <= len(sb) else 1
        if da is None or db is None:
            out.append(None if (da is None and db in (None, 1)) or (db is None and da in (None, 1)) else (da if db is None else db))
        else:
            out.append(max(da, db))
    return tuple(reversed(out))


def _unary(fn, x, name):
    x = _as_tensor(x)
    return Tensor(fn, (x,), x.shape, x.dtype, name)


def _binary(fn, a, b, name):
    a, b = _as_tensor(a), _as_tensor(b)
    return Tensor(fn, (a, b), _bshape(a, b), a.dtype if not isinstance(a, _Const) else b.dtype, name)


# ------------------------------------------------------------------------------------------------ ops


def add(a, b, name=None): return _binary(torch.add, a, b, name or 'add')
def subtract(a, b, name=None): return _binary(torch.sub, a, b, name or 'sub')
def multiply(a, b, name=None): return _binary(torch.mul, a, b, name or 'mul')
def divide(a, b, name=None): return _binary(torch.div, a, b, name or 'truediv')
def tanh(x, name=None): return _unary(torch.tanh, x, name or 'Tanh')
def exp(x, name=None): return _unary(torch.exp, x, name or 'Exp')
def log(x, name=None): return _unary(torch.log, x, name or 'Log')
def square(x, name=None): return _unary(lambda t: t * t, x, name or 'Square')
def sqrt(x, name=None): return _unary(torch.sqrt, x, name or 'Sqrt')
def identity(x, name=None): return _unary(lambda t: t, x, name or 'Identity')
def zeros_like(x, name=None): return _unary(torch.zeros_like, x, name or 'zeros_like')
def stop_gradient(x, name=None): return _unary(lambda t: t.detach(), x, name or 'StopGradient')


def minimum(x, y, name=None):
    # math_grad._MinimumGrad: the gradient goes to x where x <= y, else to y
    return _binary(lambda a, b: torch.where(a <= b, a + torch.zeros_like(b), b + torch.zeros_like(a)), x, y, name or 'Minimum')


def maximum(x, y, name=None):
    # math_grad._MaximumGrad: the gradient goes to x where x >= y, else to y
    return _binary(lambda a, b: torch.where(a >= b, a + torch.zeros_like(b), b + torch.zeros_like(a)), x, y, name or 'Maximum')


def clip_by_value(t, clip_value_min, clip_value_max, name=None):
    # clip_ops.clip_by_value: t_min = minimum(t, max); t_max = maximum(t_min, min)
    return maximum(minimum(t, clip_value_max), clip_value_min, name=name or 'clip_by_value')


def matmul(a, b, name=None):
    a, b = _as_tensor(a), _as_tensor(b)
    shp = None
    if a.shape is not None and b.shape is not None:
        shp = tuple(a.shape[:-1]) + (b.shape[-1],)
    return Tensor(torch.matmul, (a, b), shp, a.dtype, name or 'MatMul')


def _axis_of(axis, reduction_indices):
    return reduction_indices if axis is None else axis


def _reduce(tfn, x, axis, name):
    x = _as_tensor(x)
    if axis is None:
        return Tensor(lambda t: tfn(t), (x,), (), x.dtype, name)
    shp = None
    if x.shape is not None:
        ax = axis % len(x.shape)
        shp = tuple(d for i, d in enumerate(x.shape) if i != ax)
    return Tensor(lambda t: tfn(t, dim=axis), (x,), shp, x.dtype, name)


def reduce_mean(x, axis=None, keepdims=None, name=None, reduction_indices=None):
    return _reduce(torch.mean, x, _axis_of(axis, reduction_indices), name or 'Mean')


def reduce_sum(x, axis=None, keepdims=None, name=None, reduction_indices=None):
    return _reduce(torch.sum, x, _axis_of(axis, reduction_indices), name or 'Sum')


def stack(values, axis=0, name=None):
    values = [_as_tensor(v) for v in values]
    shp = None
    if values and values[0].shape is not None and axis == 0:
        shp = (len(values),) + tuple(values[0].shape)
    return Tensor(lambda *ts: torch.stack(ts, dim=axis), values, shp, values[0].dtype, name or 'stack')


def concat(values, axis=0, name=None):
    values = [_as_tensor(v) for v in values]
    return Tensor(lambda *ts: torch.cat(ts, dim=axis), values, None, values[0].dtype, name or 'concat')


def reshape(t, shape, name=None):
    t = _as_tensor(t)
    shp = tuple(None if d == -1 else d for d in shape)
    return Tensor(lambda x: x.reshape(tuple(shape)), (t,), shp, t.dtype, name or 'Reshape')


def split(value, num_or_size_splits, axis=0, name=None):
    value = _as_tensor(value)
    n = int(num_or_size_splits)
    whole = Tensor(lambda t: torch.chunk(t, n, dim=axis) if t.shape[axis] % n == 0 else _split_error(t, n),
                   (value,), None, value.dtype, name or 'split')
    shp = None
    if value.shape is not None:
        shp = list(value.shape)
        shp[axis] = None if shp[axis] is None else shp[axis] // n
    return [Tensor(lambda parts, i=i: parts[i], (whole,), shp, value.dtype, 'split_%d' % i) for i in range(n)]


def _split_error(t, n):
    raise ValueError("Dimension %d not evenly divisible by %d" % (t.shape[0], n))


def shape(t, name=None):
    t = _as_tensor(t)
    node = Tensor(lambda x: tuple(x.shape), (t,), None, int32, name or 'Shape')
    return node


def random_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    def draw(shp):
        shp = tuple(int(d) for d in shp)
        arr = _state.normal_hook(shp) if _state.normal_hook is not None else _state.rng.standard_normal(shp)
        return torch.as_tensor(np.asarray(arr)).to(_tdtype(dtype)) * stddev + mean
    if isinstance(shape, Tensor):
        return Tensor(draw, (shape,), None, dtype, name or 'random_normal')
    return Tensor(lambda: draw(shape), (), shape, dtype, name or 'random_normal')


def zeros(shape, dtype=float32, name=None):
    return Tensor(lambda: torch.zeros(tuple(shape), dtype=_tdtype(dtype)), (), shape, dtype, name or 'zeros')


def assert_rank(x, rank, **kwargs):
    x = _as_tensor(x)
    if x.shape is not None:
        assert len(x.shape) == rank, "assert_rank failed: %s has rank %d, expected %d" % (x.name, len(x.shape), rank)
    return None


def placeholder(dtype, shape=None, name=None):
    return Tensor(None, (), shape, dtype, name or 'Placeholder')


def assign(ref, value, name=None):
    value = _as_tensor(value)

    def run(v):
        ref.value = v.detach().to(_tdtype(ref.dtype)).reshape(tuple(ref.shape)).clone()
        return ref.value
    return Tensor(run, (value,), ref.shape, ref.dtype, name or 'Assign')


# ------------------------------------------------------------------------------------------------ gradients


def _reachable(src, dst, seen=None):
    """True when `dst` is an ancestor of `src` in the lazy graph."""
    seen = set() if seen is None else seen
    stack_ = [src]
    while stack_:
        n = stack_.pop()
        if n is dst:
            return True
        if id(n) in seen:
            continue
        seen.add(id(n))
        stack_.extend(n.inputs)
    return False


def gradients(ys, xs, grad_ys=None, name='gradients', **kwargs):
    """tf.gradients: d sum(ys) / d xs as graph nodes (None for inputs `ys` does not depend on)."""
    single = not isinstance(xs, (list, tuple))
    xs = [xs] if single else list(xs)
    ys_list = list(ys) if isinstance(ys, (list, tuple)) else [ys]
    connected = [any(_reachable(y, x) for y in ys_list) for x in xs]
    live = [x for x, c in zip(xs, connected) if c]

    def run(*vals):
        yv, xv = vals[:len(ys_list)], vals[len(ys_list):]
        total = sum(y.sum() for y in yv)
        return torch.autograd.grad(total, list(xv), create_graph=True, allow_unused=True)
    whole = Tensor(run, ys_list + live, None, float32, name)
    out, k = [], 0
    for x, c in zip(xs, connected):
        if not c:
            out.append(None)
            continue
        out.append(Tensor(lambda parts, x_val, i=k: parts[i] if parts[i] is not None else torch.zeros_like(x_val),
                          (whole, x), x.shape, x.dtype, name + '_%d' % k))
        k += 1
    return out[0] if single else out


# ------------------------------------------------------------------------------------------------ variables / scopes


@contextlib.contextmanager
def variable_scope(name_or_scope, reuse=None, **kwargs):
    _state.scope.append(str(name_or_scope))
    try:
        yield name_or_scope
    finally:
        _state.scope.pop()


name_scope = variable_scope


class _Graph(object):
    def get_name_scope(self):
        return '/'.join(_state.scope)


def get_default_graph():
    return _Graph()


class GraphKeys(object):
    TRAINABLE_VARIABLES = 'trainable_variables'
    GLOBAL_VARIABLES = 'variables'


def get_collection(key, scope=None):
    out = []
    for v in _state.variables:
        if key == GraphKeys.TRAINABLE_VARIABLES and not v.trainable:
            continue
        if scope is None or re.match(scope, v.name):
            out.append(v)
    return out


def global_variables():
    return list(_state.variables)


def trainable_variables():
    return get_collection(GraphKeys.TRAINABLE_VARIABLES)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True, **kwargs):
    full = '/'.join(_state.scope + [name]) + ':0'
    if full in _state.by_name:
        return _state.by_name[full]
    return Variable(name=name, dtype=dtype or float32, trainable=trainable, _initializer=initializer, _shape=shape)


def zeros_initializer(dtype=None):
    return lambda shape: np.zeros(shape)


def constant_initializer(value=0, dtype=None):
    return lambda shape: np.full(shape, value)


def _xavier_initializer(uniform=True, seed=None, dtype=None):
    # tf.contrib.layers.xavier_initializer (uniform): U(-l, l), l = sqrt(6 / (fan_in + fan_out))
    def init(shape):
        fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (int(np.prod(shape[:-1])), shape[-1])
        lim = np.sqrt(6.0 / (fan_in + fan_out))
        return _state.rng.uniform(-lim, lim, size=shape)
    return init


contrib = types.SimpleNamespace(layers=types.SimpleNamespace(xavier_initializer=_xavier_initializer))


def _dense(inputs, units, activation=None, use_bias=True, kernel_initializer=None, bias_initializer=None, name=None,
           reuse=None, **kwargs):
    inputs = _as_tensor(inputs)
    with variable_scope(name or 'dense'):
        kernel = get_variable('kernel', shape=(inputs.shape[-1], units), initializer=kernel_initializer or _xavier_initializer())
        bias = get_variable('bias', shape=(units,), initializer=bias_initializer or zeros_initializer())
        out = add(matmul(inputs, kernel), bias)
        if activation is not None:
            out = activation(out)
    return out


layers = types.SimpleNamespace(dense=_dense)


def global_variables_initializer():
    return variables_initializer(None)


def variables_initializer(var_list, name='init'):
    def run(sess, cache):
        for v in (_state.variables if var_list is None else var_list):
            v.initialize()
    return Operation(run, name)


def is_variable_initialized(var):
    return Tensor(lambda: torch.tensor(var.value is not None), (), (), bool_, 'IsVariableInitialized')


# ------------------------------------------------------------------------------------------------ session


class Session(object):
    def __init__(self, *args, **kwargs):
        self.graph = get_default_graph()

    @contextlib.contextmanager
    def as_default(self):
        _state.default_session.append(self)
        try:
            yield self
        finally:
            _state.default_session.pop()

    def __enter__(self):
        _state.default_session.append(self)
        return self

    def __exit__(self, *exc):
        _state.default_session.pop()
        return False

    def close(self):
        pass

    # -- evaluation
    def _eval(self, node, cache):
        key = id(node)
        if key in cache:
            return cache[key]
        if isinstance(node, _Const):
            arr = node._arr
            val = torch.as_tensor(arr)
            if arr.dtype.kind == 'f':
                val = val.to(_state.compute_dtype)
        elif isinstance(node, Variable):
            if node.value is None:
                raise RuntimeError("FailedPreconditionError: uninitialized variable %s" % node.name)
            val = node.value.detach().clone()
            if val.is_floating_point():
                val.requires_grad_(True)
        elif node._fn is None:
            raise RuntimeError("InvalidArgumentError: placeholder %s was not fed" % node.name)
        else:
            val = node._fn(*[self._eval(i, cache) for i in node.inputs])
        cache[key] = val
        return val

    def _feed(self, feed_dict, cache):
        for ph, val in (feed_dict or {}).items():
            t = torch.as_tensor(np.asarray(val)).to(_tdtype(ph.dtype))
            if ph.shape is not None:
                assert t.dim() == len(ph.shape) and all(d is None or d == s for d, s in zip(ph.shape, t.shape)), \
                    "cannot feed value of shape %s for %s with shape %s" % (tuple(t.shape), ph.name, ph.shape)
            if t.is_floating_point():
                t = t.clone().requires_grad_(True)
            cache[id(ph)] = t

    @staticmethod
    def _to_numpy(val):
        if isinstance(val, tuple):
            return np.asarray(val, dtype=np.int32)
        arr = val.detach().cpu().numpy()
        if arr.dtype == np.float64 and _state.compute_dtype == torch.float32:
            arr = arr.astype(np.float32)
        return arr.copy() if arr.ndim else arr[()]

    def _fetch(self, f, cache, ops):
        if isinstance(f, (list, tuple)):
            return [self._fetch(x, cache, ops) for x in f]
        if isinstance(f, dict):
            out = type(f)()
            for k, v in f.items():
                out[k] = self._fetch(v, cache, ops)
            return out
        if isinstance(f, Operation):
            ops.append(f)
            return None
        if f is None:
            raise TypeError("Fetch argument None has invalid type")
        return self._to_numpy(self._eval(f, cache))

    def run(self, fetches, feed_dict=None):
        cache, ops = {}, []
        self._feed(feed_dict, cache)
        self._cache = cache
        out = self._fetch(fetches, cache, ops)          # tensor fetches see the pre-update variables
        for op in ops:
            op._run(self, cache)
        return out


def get_default_session():
    return _state.default_session[-1] if _state.default_session else None


InteractiveSession = Session

# ------------------------------------------------------------------------------------------------ tf.train


class _AdamOptimizer(object):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, **kwargs):
        self._lr, self._b1, self._b2, self._eps = learning_rate, beta1, beta2, epsilon
        self._slots = None

    def minimize(self, loss, var_list=None, **kwargs):
        if var_list is None:
            var_list = trainable_variables()
        elif isinstance(var_list, dict):
            var_list = [var_list[k] for k in sorted(var_list)]        # nest.flatten sorts dict keys
        var_list = list(var_list)
        grads = gradients(loss, var_list)
        f32 = np.float32
        self._slots = dict(m=[None] * len(var_list), v=[None] * len(var_list), b1p=f32(self._b1), b2p=f32(self._b2))
        slots = self._slots

        def run(sess, cache):
            gvals = [None if g is None else sess._eval(g, cache).detach() for g in grads]
            if _state.compute_dtype == torch.float32:
                lr_t = f32(f32(self._lr) * np.sqrt(f32(1) - slots['b2p'], dtype=f32) / (f32(1) - slots['b1p']))
                b1, b2, eps = f32(self._b1), f32(self._b2), f32(self._eps)
            else:
                lr_t = self._lr * np.sqrt(1.0 - float(slots['b2p'])) / (1.0 - float(slots['b1p']))
                b1, b2, eps = self._b1, self._b2, self._eps
            for i, (var, g) in enumerate(zip(var_list, gvals)):
                if g is None:
                    continue
                if slots['m'][i] is None:
                    slots['m'][i], slots['v'][i] = torch.zeros_like(var.value), torch.zeros_like(var.value)
                m = slots['m'][i] = slots['m'][i] + (g - slots['m'][i]) * float(1 - b1)
                v = slots['v'][i] = slots['v'][i] + (g * g - slots['v'][i]) * float(1 - b2)
                var.value = var.value - (m * float(lr_t)) / (torch.sqrt(v) + float(eps))
            if _state.compute_dtype == torch.float32:
                slots['b1p'], slots['b2p'] = f32(slots['b1p'] * f32(self._b1)), f32(slots['b2p'] * f32(self._b2))
            else:       # exact powers
                slots['b1p'], slots['b2p'] = float(slots['b1p']) * self._b1, float(slots['b2p']) * self._b2
        return Operation(run, 'Adam')


class _GradientDescentOptimizer(object):
    def __init__(self, learning_rate, **kwargs):
        self._lr = learning_rate

    def minimize(self, loss, var_list=None, **kwargs):
        var_list = trainable_variables() if var_list is None else (
            [var_list[k] for k in sorted(var_list)] if isinstance(var_list, dict) else list(var_list))
        grads = gradients(loss, var_list)

        def run(sess, cache):
            gvals = [None if g is None else sess._eval(g, cache).detach() for g in grads]
            for var, g in zip(var_list, gvals):
                if g is not None:
                    var.value = var.value - self._lr * g
        return Operation(run, 'GradientDescent')


train = types.SimpleNamespace(AdamOptimizer=_AdamOptimizer, GradientDescentOptimizer=_GradientDescentOptimizer,
                              Optimizer=object)
nn = types.SimpleNamespace(tanh=tanh, relu=lambda x, name=None: _unary(torch.relu, x, name or 'Relu'))
summary = types.SimpleNamespace(FileWriter=lambda *a, **k: None)
