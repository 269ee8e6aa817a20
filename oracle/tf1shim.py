"""A small lazy-graph stand-in for the TensorFlow-1.3 API surface that yell/boltzmann-machines' RBM and DBM classes use
-- TEST INFRASTRUCTURE ONLY (never imported by the product package, never shipped to users).

Why it exists: the reference's arithmetic lives in `tensorflow-gpu~=1.3.0` (requirements.txt:11), which cannot be
installed in this image (Python 3.12, no network), so the reference cannot run as it is.  Its *model code*, however
-- boltzmann_machines/rbm/base_rbm.py (graph construction, gradients, sparsity, momentum, metrics, the fit loop),
rbm/rbm.py (free energies), dbm.py (mean-field, PCD, AIS), layers.py (unit types), base/tf_model.py (sessions,
persistence) -- is plain Python that
only CALLS TensorFlow.  `install()` registers this module as `tensorflow` (plus empty stand-ins for nose / matplotlib /
seaborn / keras, which the reference imports but the path does not use); tests/golden/make_reference_golden.py then
imports the reference UNMODIFIED from /root/reference and runs its own `fit()` / `transform()` / `get_tf_params()`.
What is executed is therefore the reference's formulas, in the reference's order; what is restated here is the
semantics of ~60 TensorFlow ops on numpy arrays (matmul, sigmoid, reduce_mean, assign, ...), each a few lines.

Graph model: `Tensor` nodes are lazy; `Session.run(fetches, feed_dict)` evaluates them with one memo per run, so an op
runs once per run as in TensorFlow.  Plain variable reads inside a run see the values from BEFORE the run (the
reference fetches `train_op` and the metrics in one run; TensorFlow leaves that order undefined, and "metrics see the
pre-update parameters" is what the oracle and the CUDA engine implement); the value returned by `assign`/`assign_add`
is the new one, which is what the reference's data dependencies (q_update, dW_update) rely on.

Random ops do not draw by themselves: they ask `graph.random_provider(request)`, a callable the caller installs, with
the op kind, its name scope, shape, dtype, the graph-level seed (`tf.set_random_seed`), the op-level seed, the index of
the current `Session.run` call and the `tf.while_loop` iteration.  The golden generator answers from the Philox layout
of oracle/philox.py (TF's own stream depends on graph construction order and cannot be reproduced without TF,
SURVEY.md §8c), which is what lets the oracle and the CUDA engine be compared with the reference's results bit-level
in the draws and to float rounding in everything else.
"""
import contextlib
import math
import os
import sys
import types

import numpy as np

# ------------------------------------------------------------------------------------------
# dtypes
# ------------------------------------------------------------------------------------------


class DType(object):
    def __init__(self, name):
        self.name = name
        self.np = np.dtype(name)
        self.as_numpy_dtype = self.np.type

    def __repr__(self):
        return 'tf.' + self.name

    def __eq__(self, other):
        return as_np_dtype(other) == self.np if other is not None else False

    def __hash__(self):
        return hash(self.name)


float32, float64, int32, int64 = DType('float32'), DType('float64'), DType('int32'), DType('int64')
bool = DType('bool')            # noqa: A001  (tf.bool)


def as_np_dtype(d):
    if d is None:
        return None
    if isinstance(d, DType):
        return d.np
    return np.dtype(d)


# ------------------------------------------------------------------------------------------
# graph, name scopes, collections
# ------------------------------------------------------------------------------------------
default_random_provider = None      # what a new Graph starts with (set by the caller before building models)


class GraphKeys(object):
    GLOBAL_VARIABLES = 'variables'


class Graph(object):
    def __init__(self):
        self.collections = {}
        self.by_name = {}
        self.scope = ''
        self.used_scopes = {}
        self.used_names = {}
        self.seed = None
        self.random_provider = default_random_provider
        self.control_stack = []                    # tf.control_dependencies contexts being built

    @contextlib.contextmanager
    def as_default(self):
        global _default_graph
        old = _default_graph
        _default_graph = self
        try:
            yield self
        finally:
            _default_graph = old

    def unique(self, table, name):
        n = table.get(name, 0)
        table[name] = n + 1
        return name if n == 0 else '{0}_{1}'.format(name, n)

    def get_tensor_by_name(self, name):
        return self.by_name[name]


_default_graph = Graph()


def get_default_graph():
    return _default_graph


def reset_default_graph():
    global _default_graph
    _default_graph = Graph()


def set_random_seed(seed):
    _default_graph.seed = seed


@contextlib.contextmanager
def name_scope(name, *args, **kwargs):
    g = _default_graph
    old = g.scope
    full = (old + '/' if old else '') + name
    full = g.unique(g.used_scopes, full)
    g.scope = full
    try:
        yield full + '/'
    finally:
        g.scope = old


def add_to_collection(name, value):
    _default_graph.collections.setdefault(name, []).append(value)


def get_collection(name, scope=None):
    items = list(_default_graph.collections.get(name, []))
    if scope:
        items = [v for v in items if getattr(v, 'name', '').startswith(scope)]
    return items


# ------------------------------------------------------------------------------------------
# tensors
# ------------------------------------------------------------------------------------------
class RunContext(object):
    def __init__(self, session, feed, run_index):
        self.session, self.feed, self.run_index = session, feed, run_index
        self.memo = {}
        self.snapshot = dict(session.values)       # variable values before this run
        self.loop_iter = None                      # iteration of the innermost while_loop being executed
        self.loop_stack = ()                       # ... of every enclosing while_loop, outermost first
        self.bound = {}                            # loop variables of the while_loops being executed


class Tensor(object):
    def __init__(self, op, inputs=(), fn=None, name=None, dtype=None, attrs=None):
        g = _default_graph
        self.graph = g
        self.op, self.inputs, self.fn, self.attrs = op, list(inputs), fn, attrs or {}
        self.dtype = dtype
        base = name or op
        self.scope = g.scope
        full = (g.scope + '/' if g.scope else '') + base
        full = g.unique(g.used_names, full)
        self.name = full + ':0'
        g.by_name[self.name] = self
        # ops created under tf.control_dependencies(ops) run after `ops`; TF-1 reference variables are read when
        # the consuming op executes, so such an op sees the values those assigns left behind
        self.control_inputs = [c for frame in g.control_stack for c in frame]

    # -- evaluation -------------------------------------------------------------------
    def _value(self, ctx):
        key = id(self)
        if key in ctx.memo:
            return ctx.memo[key]
        if self in ctx.feed:
            v = ctx.feed[self]
        elif self.control_inputs:
            for c in self.control_inputs:
                _val(ctx, c)
            v = self.fn(ctx, *[_val_late(ctx, i) for i in self.inputs])
        else:
            v = self.fn(ctx, *[_val(ctx, i) for i in self.inputs])
        ctx.memo[key] = v
        return v

    def eval(self, feed_dict=None, session=None):
        return (session or _default_session).run(self, feed_dict=feed_dict)

    def get_shape(self):
        return TensorShape(self.attrs.get('static_shape'))

    def __hash__(self):
        return id(self)

    def __eq__(self, other):
        return self is other

    # -- operators --------------------------------------------------------------------
    def __add__(self, o): return _binary('add', np.add, self, o)
    def __radd__(self, o): return _binary('add', np.add, o, self)
    def __sub__(self, o): return _binary('sub', np.subtract, self, o)
    def __rsub__(self, o): return _binary('sub', np.subtract, o, self)
    def __mul__(self, o): return _binary('mul', np.multiply, self, o)
    def __rmul__(self, o): return _binary('mul', np.multiply, o, self)
    def __truediv__(self, o): return _binary('div', np.true_divide, self, o)
    def __rtruediv__(self, o): return _binary('div', np.true_divide, o, self)
    __div__, __rdiv__ = __truediv__, __rtruediv__
    def __neg__(self): return _unary('neg', np.negative, self)
    def __lt__(self, o): return _binary('less', np.less, self, o, out_dtype=bool)
    def __ge__(self, o): return _binary('greater_equal', np.greater_equal, self, o, out_dtype=bool)
    def __gt__(self, o): return _binary('greater', np.greater, self, o, out_dtype=bool)
    def __le__(self, o): return _binary('less_equal', np.less_equal, self, o, out_dtype=bool)

    def __getitem__(self, idx):
        return Tensor('strided_slice', [self], lambda ctx, x: x[idx], dtype=self.dtype)

    def __iter__(self):
        raise TypeError('a Tensor is not iterable')


def _val(ctx, x):
    if isinstance(x, Tensor):
        return x._value(ctx)
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return np.stack([np.asarray(_val(ctx, e)) for e in x])
    return x


def _val_late(ctx, x):
    """Like _val, but a Variable (also inside a list) is read NOW from the session, not from the pre-run snapshot."""
    if isinstance(x, Variable):
        if x in ctx.feed:
            return ctx.feed[x]
        return ctx.session.values[x]
    if isinstance(x, (list, tuple)) and any(isinstance(e, Tensor) for e in x):
        return np.stack([np.asarray(_val_late(ctx, e)) for e in x])
    return _val(ctx, x)


@contextlib.contextmanager
def control_dependencies(ops):
    g = _default_graph
    flat = []
    for o in (ops or []):
        flat.extend(o if isinstance(o, (list, tuple)) else [o])
    g.control_stack.append([o for o in flat if isinstance(o, Tensor)])
    try:
        yield
    finally:
        g.control_stack.pop()


def _dtype_of(x):
    return as_np_dtype(x.dtype) if isinstance(x, Tensor) and x.dtype is not None else None


def _operand(v, like):
    """A non-tensor operand of a binary op becomes a constant of the tensor operand's dtype
    (tf.convert_to_tensor with a dtype hint: `x * self.sigma` with a float64 numpy sigma stays float32)."""
    if isinstance(v, Tensor):
        return v
    if isinstance(v, (list, tuple)) and any(isinstance(e, Tensor) for e in v):
        return constant(v)
    a = np.asarray(v)
    if like is not None and a.dtype.kind in 'fiub' and like.kind in 'fiu' and not (a.dtype.kind == 'f' and like.kind in 'iu'):
        a = a.astype(like)
    return Tensor('Const', [], lambda ctx: a, name='Const', dtype=DType(a.dtype.name))


def _binary(op, f, a, b, out_dtype=None):
    da, db = _dtype_of(a), _dtype_of(b)
    a, b = _operand(a, db), _operand(b, da)
    dt = out_dtype if out_dtype is not None else (a.dtype if a.dtype is not None else b.dtype)
    return Tensor(op, [a, b], lambda ctx, x, y: f(np.asarray(x), np.asarray(y)), dtype=dt)


def _unary(op, f, a, dtype=None, name=None):
    return Tensor(op, [a], lambda ctx, x: f(np.asarray(x)), dtype=dtype or getattr(a, 'dtype', None), name=name)


def convert_to_tensor(v, dtype=None, name=None):
    if isinstance(v, Tensor):
        return v
    return constant(v, dtype=dtype, name=name or 'Const')


# ------------------------------------------------------------------------------------------
# sources: constants, placeholders, variables
# ------------------------------------------------------------------------------------------
def constant(value, dtype=None, shape=None, name='Const'):
    nd = as_np_dtype(dtype)
    if isinstance(value, (list, tuple)) and any(isinstance(e, Tensor) for e in value):
        return Tensor('pack', [list(value)], lambda ctx, x: np.asarray(x, dtype=nd) if nd else np.asarray(x), name=name, dtype=dtype)
    arr = np.asarray(value, dtype=nd) if nd is not None else np.asarray(value)
    if nd is None and arr.dtype == np.float64 and not isinstance(value, np.ndarray):
        arr = arr.astype(np.float32)               # TF's default float type
    if nd is None and arr.dtype == np.int64 and not isinstance(value, np.ndarray):
        arr = arr.astype(np.int32)
    if shape is not None:
        arr = np.broadcast_to(arr, shape).copy()
    return Tensor('Const', [], lambda ctx: arr, name=name, dtype=DType(arr.dtype.name))


def placeholder(dtype, shape=None, name='Placeholder'):
    nd = as_np_dtype(dtype)

    def fn(ctx):
        raise ValueError('placeholder {0} was not fed'.format(t.name))
    t = Tensor('Placeholder', [], fn, name=name, dtype=dtype, attrs={'static_shape': shape})
    t.feed_dtype = nd
    return t


class Variable(Tensor):
    def __init__(self, initial_value, dtype=None, name='Variable', trainable=True):
        init = convert_to_tensor(initial_value, dtype=dtype)
        nd = as_np_dtype(dtype) or as_np_dtype(init.dtype)
        Tensor.__init__(self, 'Variable', [], None, name=name, dtype=DType(nd.name))
        self.initial_value = init
        self.np_dtype = nd
        self.fn = self._read
        add_to_collection(GraphKeys.GLOBAL_VARIABLES, self)

    def _read(self, ctx):
        if self not in ctx.snapshot:
            raise RuntimeError('variable {0} is not initialized'.format(self.name))
        return ctx.snapshot[self]

    def _write(self, ctx, v):
        v = np.asarray(v, dtype=self.np_dtype)
        ctx.session.values[self] = v
        return v

    def assign(self, value):
        return Tensor('Assign', [value], lambda ctx, v: self._write(ctx, v), dtype=self.dtype)

    def assign_add(self, delta):
        # reads the pre-run value: every variable is assigned at most once per run in the reference's graphs
        return Tensor('AssignAdd', [delta], lambda ctx, d: self._write(ctx, ctx.snapshot[self] + np.asarray(d, dtype=self.np_dtype)),
                      dtype=self.dtype)

    def initialized_value(self):
        return self.initial_value


def global_variables_initializer():
    vs = get_collection(GraphKeys.GLOBAL_VARIABLES)

    def fn(ctx, *inits):
        for v, i in zip(vs, inits):
            v._write(ctx, i)
        return None
    return Tensor('init', [v.initial_value for v in vs], fn)


def group(*ops, **kwargs):
    return Tensor('group', list(ops), lambda ctx, *xs: None, name=kwargs.get('name', 'group'))


def identity(x, name=None):
    x = convert_to_tensor(x)
    return Tensor('Identity', [x], lambda ctx, v: v, name=name or 'Identity', dtype=x.dtype)


# ------------------------------------------------------------------------------------------
# element-wise, shapes, reductions, linear algebra
# ------------------------------------------------------------------------------------------
def cast(x, dtype, name=None):
    nd = as_np_dtype(dtype)
    return Tensor('Cast', [x], lambda ctx, v: np.asarray(v).astype(nd), name=name or 'Cast', dtype=dtype)


def to_float(x, name='ToFloat'):
    return cast(x, float32, name)


def to_int64(x, name='ToInt64'):
    return cast(x, int64, name)


def to_int32(x, name='ToInt32'):
    return cast(x, int32, name)


def add(a, b, name=None): return _binary('Add', np.add, a, b)
def subtract(a, b, name=None): return _binary('Sub', np.subtract, a, b)
def multiply(a, b, name=None):
    t = _binary('Mul', np.multiply, a, b)
    return identity(t, name) if name else t
def divide(a, b, name=None): return _binary('Div', np.true_divide, a, b)
def square(x, name=None): return _unary('Square', np.square, x)
def exp(x, name=None): return _unary('Exp', np.exp, x)
def log(x, name=None): return _unary('Log', np.log, convert_to_tensor(x))
def sqrt(x, name=None): return _unary('Sqrt', np.sqrt, x)
def negative(x, name=None): return _unary('Neg', np.negative, x)


def assign(ref, value, validate_shape=None, use_locking=None, name=None):
    return ref.assign(value)


def minimum(a, b, name=None): return _binary('Minimum', np.minimum, a, b)
def maximum(a, b, name=None): return _binary('Maximum', np.maximum, a, b)
def logical_and(a, b, name=None): return _binary('LogicalAnd', np.logical_and, a, b, out_dtype=bool)


def clip_by_value(x, lo, hi, name=None):
    d = _dtype_of(x)
    return Tensor('ClipByValue', [x], lambda ctx, v: np.clip(v, np.asarray(lo, dtype=d or None), np.asarray(hi, dtype=d or None)),
                  dtype=x.dtype)


def norm(x, ord='euclidean', axis=None, keep_dims=False, name=None):        # noqa: A002
    def fn(ctx, v):
        v = np.asarray(v)
        if ord in (np.inf, 'inf'):
            r = np.max(np.abs(v), axis=axis, keepdims=keep_dims)
        elif ord in ('euclidean', 2, 'fro'):
            r = np.sqrt(np.sum(np.square(v), axis=axis, keepdims=keep_dims))
        elif ord == 1:
            r = np.sum(np.abs(v), axis=axis, keepdims=keep_dims)
        else:
            raise NotImplementedError(ord)
        return np.asarray(r, dtype=v.dtype)
    return Tensor('Norm', [x], fn, dtype=x.dtype)


class TensorShape(object):
    def __init__(self, dims=None):
        self.dims = dims


def _lgamma(x):
    x = np.asarray(x)
    f = np.vectorize(math.lgamma, otypes=[x.dtype if x.dtype.kind == 'f' else np.float32])
    return f(x)


def lgamma(x, name=None):
    x = convert_to_tensor(x)
    return _unary('Lgamma', _lgamma, x)


def _softplus(x):
    return np.logaddexp(x, 0).astype(x.dtype, copy=False)


def _sigmoid(x):
    # the stable two-branch form, evaluated in the array's dtype
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1.0 / (1.0 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1.0 + e)
    return out


def log_sigmoid(x, name=None):
    return _unary('LogSigmoid', lambda v: (-_softplus(-v)).astype(v.dtype, copy=False), x)


def shape(x, name=None):
    return Tensor('Shape', [x], lambda ctx, v: np.asarray(np.shape(v), dtype=np.int32), dtype=int32)


def reshape(x, shp, name=None):
    return Tensor('Reshape', [x, shp], lambda ctx, v, s: np.reshape(v, [int(i) for i in np.asarray(s).reshape(-1)]),
                  dtype=getattr(x, 'dtype', None))


def transpose(x, perm=None, name=None):
    x = convert_to_tensor(x)
    return Tensor('Transpose', [x], lambda ctx, v: np.transpose(v, perm), dtype=x.dtype)


def expand_dims(x, axis, name=None):
    return Tensor('ExpandDims', [x], lambda ctx, v: np.expand_dims(v, axis), dtype=x.dtype)


def _shape_list(ctx_shape):
    return [int(i) for i in np.asarray(ctx_shape).reshape(-1)]


def zeros(shp, dtype=float32, name=None):
    nd = as_np_dtype(dtype)
    return Tensor('zeros', [shp], lambda ctx, s: np.zeros(_shape_list(s), dtype=nd), dtype=dtype)


def ones(shp, dtype=float32, name=None):
    nd = as_np_dtype(dtype)
    return Tensor('ones', [shp], lambda ctx, s: np.ones(_shape_list(s), dtype=nd), dtype=dtype)


def zeros_like(x, dtype=None, name=None):
    nd = as_np_dtype(dtype)
    return Tensor('zeros_like', [x], lambda ctx, v: np.zeros_like(v, dtype=nd), dtype=dtype or x.dtype)


def ones_like(x, dtype=None, name=None):
    nd = as_np_dtype(dtype)
    return Tensor('ones_like', [x], lambda ctx, v: np.ones_like(v, dtype=nd), dtype=dtype or x.dtype)


def range(*args, **kwargs):      # noqa: A001  (tf.range)
    return Tensor('Range', list(args), lambda ctx, *a: np.arange(*[int(i) for i in a], dtype=np.int32), dtype=int32)


def _reduce(op, f):
    def make(x, axis=None, keep_dims=False, name=None):
        def fn(ctx, v):
            v = np.asarray(v)
            r = f(v, axis=axis, keepdims=keep_dims)
            return r.astype(v.dtype, copy=False) if v.dtype.kind == 'f' else r
        return Tensor(op, [x], fn, dtype=getattr(x, 'dtype', None))
    return make


reduce_sum = _reduce('Sum', np.sum)
reduce_mean = _reduce('Mean', np.mean)
reduce_max = _reduce('Max', np.max)


def matmul(a, b, transpose_a=False, transpose_b=False, name=None):
    def fn(ctx, x, y):
        x = x.T if transpose_a else x
        y = y.T if transpose_b else y
        return x @ y
    return Tensor('MatMul', [a, b], fn, dtype=getattr(a, 'dtype', None) or getattr(b, 'dtype', None))


def einsum(eq, *xs):
    hint = next((_dtype_of(x) for x in xs if _dtype_of(x) is not None), None)
    xs = [_operand(x, hint) for x in xs]
    return Tensor('Einsum', list(xs), lambda ctx, *vs: np.einsum(eq, *vs),
                  dtype=next((x.dtype for x in xs if isinstance(x, Tensor)), None))


# ------------------------------------------------------------------------------------------
# sparse (the pseudo-likelihood corruption, base_rbm.py:496-509)
# ------------------------------------------------------------------------------------------
class SparseTensor(object):
    def __init__(self, indices, values, dense_shape):
        self.indices, self.values, self.dense_shape = indices, values, dense_shape


def sparse_tensor_to_dense(sp, default_value=0, name=None):
    def fn(ctx, idx, vals, shp):
        vals = np.asarray(vals)
        out = np.full(_shape_list(shp), default_value, dtype=vals.dtype)
        out[tuple(np.asarray(idx, dtype=np.int64).T)] = vals
        return out
    return Tensor('SparseToDense', [sp.indices, sp.values, sp.dense_shape], fn, dtype=getattr(sp.values, 'dtype', None))


def sparse_add(a, b, name=None):
    dense, sp = (a, b) if isinstance(b, SparseTensor) else (b, a)

    def fn(ctx, d, idx, vals):
        out = np.array(d, copy=True)
        np.add.at(out, tuple(np.asarray(idx, dtype=np.int64).T), np.asarray(vals, dtype=out.dtype))
        return out
    return Tensor('SparseAdd', [dense, sp.indices, sp.values], fn, dtype=dense.dtype)


# ------------------------------------------------------------------------------------------
# random ops: values come from graph.random_provider
# ------------------------------------------------------------------------------------------
class RandomRequest(object):
    """What a random op asks the provider for.  kind in {'uniform', 'uniform_int', 'normal', 'bernoulli',
    'multinomial', 'dropout_uniform'}; `args` holds the op's evaluated inputs (probs, loc/scale, ...)."""
    def __init__(self, **kw):
        self.__dict__.update(kw)


def _random(kind, inputs, shape_of, dtype, op_seed=None, name=None, post=None, **extra):
    t_holder = []

    def fn(ctx, *vals):
        t = t_holder[0]
        prov = t.graph.random_provider
        if prov is None:
            raise RuntimeError('no random provider installed on the graph (tf1shim)')
        req = RandomRequest(kind=kind, scope=t.scope, name=t.name, shape=tuple(shape_of(*vals)),
                            dtype=as_np_dtype(dtype), graph_seed=t.graph.seed, op_seed=op_seed, graph=t.graph,
                            run_index=ctx.run_index, loop_iter=ctx.loop_iter, loop_stack=ctx.loop_stack,
                            args=vals, **extra)
        out = prov(req)
        return post(out, *vals) if post else out
    t = Tensor(kind, inputs, fn, name=name or kind, dtype=dtype)
    t.op_seed, t.random_kind = op_seed, kind
    t_holder.append(t)
    return t


def random_uniform(shp, minval=0, maxval=None, dtype=float32, seed=None, name=None):
    nd = as_np_dtype(dtype)
    if nd.kind in 'iu':
        return _random('uniform_int', [shp, minval, maxval], lambda s, lo, hi: _shape_list(s), dtype, seed, name,
                       post=lambda words, s, lo, hi: (int(lo) + (np.asarray(words, dtype=np.uint32) % np.uint32(int(hi) - int(lo)))).astype(nd))
    hi = 1.0 if maxval is None else maxval
    return _random('uniform', [shp], lambda s: _shape_list(s), dtype, seed, name,
                   post=lambda u, s: (minval + (hi - minval) * u).astype(nd) if (minval != 0 or hi != 1) else u.astype(nd))


def random_normal(shp, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    nd = as_np_dtype(dtype)
    return _random('normal', [shp], lambda s: _shape_list(s), dtype, seed, name, stddev=stddev,
                   post=lambda z, s: (mean + z).astype(nd) if mean != 0.0 else np.asarray(z, dtype=nd))


class _NN(object):
    @staticmethod
    def sigmoid(x, name=None):
        return _unary('Sigmoid', _sigmoid, x)

    @staticmethod
    def softplus(x, name=None):
        return _unary('Softplus', _softplus, x)

    @staticmethod
    def softmax(x, name=None):
        def f(v):
            z = v - v.max(axis=-1, keepdims=True)
            e = np.exp(z)
            return (e / e.sum(axis=-1, keepdims=True)).astype(v.dtype, copy=False)
        return _unary('Softmax', f, x)

    @staticmethod
    def l2_loss(x, name=None):
        return Tensor('L2Loss', [x], lambda ctx, v: (np.sum(np.square(v), dtype=np.float64) / 2).astype(np.asarray(v).dtype),
                      dtype=x.dtype)

    @staticmethod
    def dropout(x, keep_prob, noise_shape=None, seed=None, name=None):
        # tf.nn.dropout: x / keep_prob * floor(keep_prob + uniform)
        u = _random('dropout_uniform', [x], lambda v: np.shape(v), x.dtype, seed, 'dropout/random_uniform')

        def fn(ctx, v, keep, uu):
            keep = np.asarray(keep, dtype=v.dtype)
            return v / keep * np.floor(keep + uu.astype(v.dtype))
        return Tensor('dropout', [x, keep_prob, u], fn, dtype=x.dtype)


nn = _NN()


class _Distribution(object):
    pass


class Bernoulli(_Distribution):
    def __init__(self, logits=None, probs=None, dtype=int32, name='Bernoulli'):
        self.probs = probs if probs is not None else nn.sigmoid(logits)

    def sample(self, sample_shape=(), seed=None, name='sample'):
        return _random('bernoulli', [self.probs], lambda p: np.shape(p), int32, seed, 'Bernoulli/sample')


class Normal(_Distribution):
    def __init__(self, loc, scale, name='Normal'):
        self.loc, self.scale = loc, scale

    def sample(self, sample_shape=(), seed=None, name='sample'):
        return _random('normal_loc_scale', [self.loc, self.scale], lambda m, s: np.shape(m), getattr(self.loc, 'dtype', float32),
                       seed, 'Normal/sample')


class Multinomial(_Distribution):
    def __init__(self, total_count, logits=None, probs=None, name='Multinomial'):
        self.total_count, self.logits, self.probs = total_count, logits, probs

    def sample(self, sample_shape=(), seed=None, name='sample'):
        src = self.probs if self.probs is not None else self.logits
        return _random('multinomial', [src], lambda p: np.shape(p), float32, seed, 'Multinomial/sample',
                       total_count=self.total_count, from_logits=self.probs is None)


# ------------------------------------------------------------------------------------------
# control flow
# ------------------------------------------------------------------------------------------
def _flatten(struct):
    if isinstance(struct, (list, tuple)):
        out = []
        for e in struct:
            out.extend(_flatten(e))
        return out
    return [struct]


def _unflatten(struct, flat):
    """Rebuild `struct`'s nesting (lists of lists) from the iterator `flat`."""
    if isinstance(struct, (list, tuple)):
        return [_unflatten(e, flat) for e in struct]
    return next(flat)


def while_loop(cond, body, loop_vars, shape_invariants=None, back_prop=True, parallel_iterations=10, name=None, **kwargs):
    """Builds cond/body once on symbolic loop variables (nested lists allowed); every output tensor re-runs the loop
    when evaluated in a run (memoised per run, so the outputs of one loop share one execution).  A loop built under
    tf.control_dependencies runs after those ops and reads Variable loop inputs as they are then."""
    g = _default_graph
    control = [c for frame in g.control_stack for c in frame]
    with name_scope(name or 'while'):
        init = [v if isinstance(v, Tensor) else convert_to_tensor(v) for v in _flatten(loop_vars)]
        syms = []
        for i, v in enumerate(init):
            s = Tensor('loop_var', [], None, name='loop_var_%d' % i, dtype=v.dtype)
            s.fn = (lambda ss: (lambda ctx: ctx.bound[ss]))(s)
            s.control_inputs = []
            syms.append(s)
        sym_struct = _unflatten(loop_vars, iter(syms))
        cond_t = convert_to_tensor(cond(*sym_struct))
        outs = [o if isinstance(o, Tensor) else convert_to_tensor(o) for o in _flatten(list(body(*_unflatten(loop_vars, iter(syms)))))]
        assert len(outs) == len(init), 'while_loop body returned another structure'
    state = {}

    def run_loop(ctx):
        key = ('while', id(state))
        if key in ctx.memo:
            return ctx.memo[key]
        for c in control:
            _val(ctx, c)
        vals = [(_val_late if control else _val)(ctx, v) for v in init]
        outer_memo, outer_bound, outer_iter, outer_stack = ctx.memo, ctx.bound, ctx.loop_iter, ctx.loop_stack
        it = 0
        while True:
            ctx.memo = dict(outer_memo)             # nodes of the enclosing run / iteration stay evaluated once
            ctx.bound = dict(outer_bound)
            ctx.bound.update(zip(syms, vals))
            ctx.loop_iter, ctx.loop_stack = it, outer_stack + (it,)
            if not np.asarray(_val(ctx, cond_t)).item():
                break
            vals = [_val(ctx, o) for o in outs]
            it += 1
        ctx.memo, ctx.bound, ctx.loop_iter, ctx.loop_stack = outer_memo, outer_bound, outer_iter, outer_stack
        ctx.memo[key] = vals
        return vals
    results = []
    for i in builtins_range(len(init)):
        t = Tensor('while_out', [], (lambda k: (lambda ctx: run_loop(ctx)[k]))(i), name='Exit', dtype=init[i].dtype)
        t.control_inputs = []                       # handled inside run_loop
        results.append(t)
    return _unflatten(loop_vars, iter(results))


import builtins as _builtins      # noqa: E402
builtins_range = _builtins.range


# ------------------------------------------------------------------------------------------
# sessions, savers, summaries
# ------------------------------------------------------------------------------------------
_default_session = None


class ConfigProto(object):
    def __init__(self, *args, **kwargs):
        pass


class Session(object):
    def __init__(self, config=None, graph=None, target=''):
        self.graph = graph or _default_graph
        self.values = {}
        self.n_runs = 0

    def __enter__(self):
        global _default_session
        self._outer = _default_session
        _default_session = self
        return self

    def __exit__(self, *exc):
        global _default_session
        _default_session = self._outer
        return False

    def close(self):
        pass

    def _resolve(self, t):
        return self.graph.get_tensor_by_name(t) if isinstance(t, str) else t

    def run(self, fetches, feed_dict=None):
        feed = {}
        for k, v in (feed_dict or {}).items():
            t = self._resolve(k)
            nd = getattr(t, 'feed_dtype', None) or as_np_dtype(t.dtype)
            feed[t] = np.asarray(v, dtype=nd)
        is_init = isinstance(fetches, Tensor) and fetches.op == 'init'
        ctx = RunContext(self, feed, self.n_runs)
        if not is_init:
            self.n_runs += 1                         # initialisation does not count as a step of the model
        single = not isinstance(fetches, (list, tuple))
        out = [None if f is None else _val(ctx, self._resolve(f)) for f in ([fetches] if single else fetches)]
        return out[0] if single else out


class _Saver(object):
    """Variables -> <path>[-<step>].npz; the graph itself stays in this process (`_META`), keyed by the .meta path the
    reference re-imports in `run_in_tf_session`."""
    def __init__(self, graph=None, **kwargs):
        self.graph = graph or _default_graph

    def save(self, sess, save_path, global_step=None, **kwargs):
        path = save_path if global_step is None else '{0}-{1}'.format(save_path, global_step)
        d = os.path.dirname(path)
        if d and not os.path.isdir(d):
            os.makedirs(d)
        vs = self.graph.collections.get(GraphKeys.GLOBAL_VARIABLES, [])
        np.savez(path + '.npz', **{v.name: sess.values[v] for v in vs if v in sess.values})
        open(path + '.meta', 'w').write('tf1shim graph\n')
        _META[os.path.abspath(path + '.meta')] = self.graph
        _META[os.path.abspath(save_path + '.meta')] = self.graph
        np.savez(save_path + '.npz', **{v.name: sess.values[v] for v in vs if v in sess.values})
        return path

    def restore(self, sess, save_path):
        data = np.load(save_path + '.npz')
        for v in self.graph.collections.get(GraphKeys.GLOBAL_VARIABLES, []):
            if v.name in data.files:
                sess.values[v] = np.asarray(data[v.name], dtype=v.np_dtype)


_META = {}


class _Train(object):
    Saver = _Saver

    @staticmethod
    def import_meta_graph(meta_path, **kwargs):
        src = _META[os.path.abspath(meta_path)]
        g = _default_graph
        for attr in ('collections', 'by_name', 'used_scopes', 'used_names'):
            setattr(g, attr, getattr(src, attr))
        if g.random_provider is None:
            g.random_provider = src.random_provider
        for t in g.by_name.values():
            t.graph = g
        return _Saver(g)


train = _Train()


class _Summary(object):
    class FileWriter(object):
        def __init__(self, *args, **kwargs):
            self.events = []

        def add_summary(self, summary, global_step=None):
            self.events.append((global_step, summary))

        def close(self):
            pass

        def flush(self):
            pass

    @staticmethod
    def _nop(*args, **kwargs):
        return Tensor('summary', [], lambda ctx: None)

    scalar = histogram = image = _nop

    @staticmethod
    def merge_all(*args, **kwargs):
        return Tensor('merge_summaries', [], lambda ctx: None)


summary = _Summary()


class _SummaryProto(object):
    class Value(object):
        def __init__(self, **kw):
            self.__dict__.update(kw)

    def __init__(self, **kw):
        self.__dict__.update(kw)


# ------------------------------------------------------------------------------------------
# installation
# ------------------------------------------------------------------------------------------
def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install(provider_hook=None):
    """Register this module as `tensorflow` (and inert stand-ins for the plotting / test-runner / keras imports of
    the reference's utils package) unless the real packages exist.  Returns this module."""
    me = sys.modules[__name__]
    sys.modules['tensorflow'] = me
    pb = _module('tensorflow.core.framework.summary_pb2', Summary=_SummaryProto)
    fw = _module('tensorflow.core.framework', summary_pb2=pb)
    core = _module('tensorflow.core', framework=fw)
    dist = _module('tensorflow.contrib.distributions', Bernoulli=Bernoulli, Multinomial=Multinomial, Normal=Normal)
    contrib = _module('tensorflow.contrib', distributions=dist)
    me.core, me.contrib = core, contrib
    sys.modules.update({'tensorflow.core': core, 'tensorflow.core.framework': fw,
                        'tensorflow.core.framework.summary_pb2': pb, 'tensorflow.contrib': contrib,
                        'tensorflow.contrib.distributions': dist})

    def inert(name, **attrs):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _module(name, **attrs)
        return sys.modules[name]
    tools = inert('nose.tools', nottest=lambda f: f)
    inert('nose', tools=tools, run=lambda *a, **k: None)
    plt = inert('matplotlib.pyplot')
    anim = inert('matplotlib.animation', FuncAnimation=object)
    inert('matplotlib', pyplot=plt, animation=anim)
    inert('seaborn')
    return me


Tensor.__module__ = 'tensorflow'
