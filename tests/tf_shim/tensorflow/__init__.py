"""A stand-in for the `tensorflow` package, just large enough to EXECUTE the reference's own model code.

TEST INFRASTRUCTURE ONLY (used by oracle/make_golden_mlp.py and tests/test_tf_shim.py; the product never imports
it).  TensorFlow is not installable in this image, so the reference's `src/linear_model.py` (TF1 graph API) and
`src/top_vae_3d_pose/models.py::PoseBase` (TF2 eager API) cannot run as shipped.  This package implements only the
symbols those two touch, on torch-CPU autograd, so that the graph *wiring* - which op feeds which, where the bias,
BatchNorm, ReLU, dropout, residual and clip sit, what the loss averages over, which variables are trainable, what the
train op depends on - is taken from the reference's unmodified source instead of from a restatement.

What stays restated here is the semantics of each TensorFlow op, written from TensorFlow's published sources
(r1.x; file and function named at each op):

  tf.clip_by_norm           python/ops/clip_ops.py::clip_by_norm (axes=None: whole-tensor L2 norm)
  tf.layers.batch_normalization / keras.layers.BatchNormalization
                            python/keras/layers/normalization.py::BatchNormalizationBase (non-fused path for rank-2
                            input): nn.moments (biased variance), nn.batch_normalization, assign_moving_average,
                            momentum 0.99, epsilon 1e-3, update ops collected in GraphKeys.UPDATE_OPS
  tf.nn.dropout             python/ops/nn_ops.py::dropout (TF1 keep_prob form: x/keep * floor(keep + U[0,1)))
  tf.train.AdamOptimizer    python/training/adam.py + core/kernels/training_ops.cc::ApplyAdam
  tf.train.exponential_decay python/training/learning_rate_decay.py (non-staircase)
  tf.truncated_normal       core/kernels/random_op.cc (re-draw beyond two standard deviations)

Graph tensors are lazy nodes evaluated per `Session.run` with a per-run cache, so that - as in one TensorFlow graph
execution - `loss`, `outputs` and the gradients of one `run` all come from the same forward pass on the pre-update
variables.  The dropout noise of a run can be injected (`set_dropout_masks`) so that runs are reproducible.
"""
from __future__ import annotations

import contextlib

import numpy as np
import torch

__version__ = "1.15-shim"


# --------------------------------------------------------------------------------------------------- dtypes
class DType:
    def __init__(self, name, tdt, ndt):
        self.name, self.torch, self.as_numpy_dtype = name, tdt, ndt

    def __repr__(self):
        return "tf." + self.name


float32 = DType("float32", torch.float32, np.float32)
float64 = DType("float64", torch.float64, np.float64)
float16 = DType("float16", torch.float16, np.float16)
int32 = DType("int32", torch.int64, np.int64)
int64 = DType("int64", torch.int64, np.int64)
bool = DType("bool", torch.bool, np.bool_)          # noqa: A001  (tf.bool)
float = float32                                       # noqa: A001  (models.py mentions tf.float)


def _as_dtype(d):
    if d is None:
        return float32
    if isinstance(d, DType):
        return d
    return {np.float32: float32, np.float64: float64}[d]


# --------------------------------------------------------------------------------------------------- graph state
class GraphKeys:
    UPDATE_OPS = "update_ops"
    GLOBAL_VARIABLES = "variables"
    TRAINABLE_VARIABLES = "trainable_variables"


class _Graph:
    def __init__(self):
        self.collections = {GraphKeys.UPDATE_OPS: [], GraphKeys.GLOBAL_VARIABLES: [],
                            GraphKeys.TRAINABLE_VARIABLES: []}
        self.scopes = []              # variable_scope / name_scope stack
        self.control = []             # control_dependencies stack
        self.dropout_ops = []         # in graph-construction order
        self.dropout_masks = None     # injected: list aligned with dropout_ops (0/1 arrays), or None
        self.rng = np.random.RandomState(0)


_G = _Graph()


def reset_default_graph():
    global _G
    _G = _Graph()


def set_random_seed(seed):
    _G.rng = np.random.RandomState(seed)


def set_dropout_masks(masks):
    """Shim control: masks[i] (0/1 array, shape of the i-th dropout op's input, graph order) replaces
    floor(keep_prob + U[0,1)) in the runs that follow; None restores random noise."""
    _G.dropout_masks = masks


def get_collection(key):
    return list(_G.collections.get(key, []))


def global_variables():
    return get_collection(GraphKeys.GLOBAL_VARIABLES)


def trainable_variables():
    return get_collection(GraphKeys.TRAINABLE_VARIABLES)


def _scope_prefix():
    return "".join(s + "/" for s in _G.scopes)


@contextlib.contextmanager
def variable_scope(name, reuse=None):
    _G.scopes.append(name)
    try:
        yield name
    finally:
        _G.scopes.pop()


name_scope = variable_scope


@contextlib.contextmanager
def control_dependencies(ops):
    _G.control.append(list(ops))
    try:
        yield
    finally:
        _G.control.pop()


# --------------------------------------------------------------------------------------------------- tensors
class _Run:
    """One graph execution: the feed and a cache (each node is evaluated once)."""

    def __init__(self, feed=None):
        self.feed = {} if feed is None else {id(k): v for k, v in feed.items()}
        self.cache = {}


def _val(x, run, like=None):
    if isinstance(x, Tensor):
        return x._eval(run)
    if isinstance(x, torch.Tensor):
        return x
    dt = like.dtype if isinstance(like, torch.Tensor) and like.dtype.is_floating_point else torch.float64
    return torch.as_tensor(np.asarray(x), dtype=dt)


def _static_shape(x):
    if isinstance(x, Tensor):
        return x._shape
    return list(np.shape(x))


def _bshape(a, b):
    sa, sb = _static_shape(a), _static_shape(b)
    if sa is None or sb is None:
        return sa if sb is None else sb
    return sa if len(sa) >= len(sb) else sb


class Tensor:
    def __init__(self, fn, inputs=(), shape=None, name=None, dtype=None):
        self._fn, self._inputs, self._shape, self.name, self.dtype = fn, tuple(inputs), shape, name, dtype
        self._control = [op for ops in _G.control for op in ops]

    # evaluation -------------------------------------------------------------------------------------
    def _eval(self, run):
        k = id(self)
        if k not in run.cache:
            for op in self._control:
                op._eval(run)
            run.cache[k] = self._fn(run)
        return run.cache[k]

    def eval(self, session=None, feed_dict=None):
        return _to_numpy(self._eval(_Run(feed_dict)))

    def numpy(self):
        return self.eval()

    def get_shape(self):
        return self._shape

    @property
    def shape(self):
        return self._shape

    # arithmetic -------------------------------------------------------------------------------------
    def _bin(self, other, f, rev=False):
        a, b = (other, self) if rev else (self, other)

        def fn(run):
            ta = _val(a, run, like=_val(self, run))
            tb = _val(b, run, like=_val(self, run))
            return f(ta, tb)
        return Tensor(fn, [t for t in (a, b) if isinstance(t, Tensor)], _bshape(a, b), dtype=self.dtype)

    def __add__(self, o): return self._bin(o, torch.add)
    def __radd__(self, o): return self._bin(o, torch.add, True)
    def __sub__(self, o): return self._bin(o, torch.sub)
    def __rsub__(self, o): return self._bin(o, torch.sub, True)
    def __mul__(self, o): return self._bin(o, torch.mul)
    def __rmul__(self, o): return self._bin(o, torch.mul, True)
    def __truediv__(self, o): return self._bin(o, torch.div)
    def __rtruediv__(self, o): return self._bin(o, torch.div, True)
    def __neg__(self): return Tensor(lambda run: -self._eval(run), [self], self._shape, dtype=self.dtype)
    __hash__ = object.__hash__


def _to_numpy(v):
    if isinstance(v, torch.Tensor):
        a = v.detach().numpy()
        return a.copy() if a.ndim else a[()]
    if isinstance(v, (list, tuple)):
        return [_to_numpy(u) for u in v]
    return v


def placeholder(dtype, shape=None, name=None):
    dtype = _as_dtype(dtype)
    t = Tensor(None, (), list(shape) if shape is not None else None, name, dtype)

    def fn(run):
        if id(t) not in run.feed:
            raise ValueError("placeholder %r was not fed" % (name,))
        v = run.feed[id(t)]
        if dtype is bool:
            return v if isinstance(v, Tensor) else builtins_bool(v)
        return torch.as_tensor(np.asarray(v), dtype=dtype.torch)     # the feed casts (fp64 arrays -> fp32 graph)
    t._fn = fn
    return t


import builtins as _builtins   # noqa: E402
builtins_bool = _builtins.bool


def constant(v, dtype=None):
    dt = _as_dtype(dtype).torch if dtype is not None else None
    tv = torch.as_tensor(np.asarray(v), dtype=dt)
    return Tensor(lambda run: tv, (), list(tv.shape), dtype=dtype)


class Variable(Tensor):
    """tf.Variable (TF1 positional initial_value; TF2 keyword initial_value / shape)."""

    def __init__(self, initial_value=None, trainable=True, dtype=None, name=None, shape=None):
        if isinstance(initial_value, Tensor):
            init = initial_value._eval(_Run()).detach().clone()
            if dtype is not None:
                init = init.to(_as_dtype(dtype).torch)
        elif isinstance(initial_value, (int, np.integer)) and not isinstance(initial_value, builtins_bool) \
                and dtype is None:
            init = torch.tensor(int(initial_value), dtype=torch.int64)
        else:
            init = torch.as_tensor(np.asarray(initial_value),
                                   dtype=_as_dtype(dtype).torch if dtype is not None else None)
            if dtype is None and init.dtype == torch.float64 and np.asarray(initial_value).dtype != np.float64:
                init = init.to(torch.float32)
        self.value = init
        if trainable and init.dtype.is_floating_point:
            self.value.requires_grad_(True)
        self.trainable = builtins_bool(trainable)
        full = _scope_prefix() + (name or "Variable")
        super().__init__(lambda run: self.value, (), list(init.shape), full, dtype)
        _G.collections[GraphKeys.GLOBAL_VARIABLES].append(self)
        if self.trainable:
            _G.collections[GraphKeys.TRAINABLE_VARIABLES].append(self)

    def load(self, value, session=None):
        with torch.no_grad():
            self.value.copy_(torch.as_tensor(np.asarray(value), dtype=self.value.dtype))

    def assign(self, value):
        self.load(value)
        return self

    def eval(self, session=None, feed_dict=None):
        return _to_numpy(self.value)


def get_variable(name, shape=None, dtype=None, initializer=None, trainable=True):
    dtype = _as_dtype(dtype)
    init = initializer(list(shape), dtype)       # an initializer is called as f(shape, dtype[, partition_info])
    return Variable(init, trainable=trainable, dtype=dtype, name=name)


def global_variables_initializer():
    return Tensor(lambda run: None)


# --------------------------------------------------------------------------------------------------- ops
def truncated_normal(shape, mean=0.0, stddev=1.0, dtype=float32, seed=None, name=None):
    """Normal samples re-drawn while they lie beyond two standard deviations."""
    dtype = _as_dtype(dtype)
    n = int(np.prod(shape))
    v = _G.rng.standard_normal(n)
    bad = np.abs(v) > 2.0
    while bad.any():
        v[bad] = _G.rng.standard_normal(int(bad.sum()))
        bad = np.abs(v) > 2.0
    tv = torch.as_tensor((v * stddev + mean).reshape(shape), dtype=dtype.torch)
    return Tensor(lambda run: tv, (), list(shape), dtype=dtype)


def sqrt(x, name=None):
    if isinstance(x, Tensor):
        return Tensor(lambda run: torch.sqrt(x._eval(run)), [x], x._shape, dtype=x.dtype)
    return builtins_float(np.sqrt(x))


builtins_float = _builtins.float


def exp(x, name=None):
    return Tensor(lambda run: torch.exp(_val(x, run)), [x] if isinstance(x, Tensor) else (), _static_shape(x))


def square(x, name=None):
    return Tensor(lambda run: torch.square(x._eval(run)), [x], x._shape, dtype=x.dtype)


def matmul(a, b, name=None):
    def fn(run):
        tb = _val(b, run)
        ta = _val(a, run, like=tb)
        return torch.matmul(ta.to(tb.dtype), tb)
    sa, sb = _static_shape(a), _static_shape(b)
    return Tensor(fn, [t for t in (a, b) if isinstance(t, Tensor)],
                  [sa[0] if sa else None, sb[1] if sb else None],
                  dtype=getattr(b, "dtype", None))


def reduce_mean(x, axis=None, name=None):
    def fn(run):
        v = x._eval(run)
        return v.mean() if axis is None else v.mean(dim=axis)
    return Tensor(fn, [x], [], dtype=x.dtype)


def clip_by_norm(t, clip_norm, axes=None, name=None):
    """clip_ops.clip_by_norm: l2sum = sum(t*t, axes) [axes=None: over everything]; l2norm = sqrt(l2sum) where
    l2sum > 0; result = t*clip_norm / maximum(l2norm, clip_norm).  Differentiable; the gradient flows through the norm."""
    if axes is not None:
        raise NotImplementedError("the reference never passes axes")

    def fn(run):
        v = t._eval(run)
        l2sum = (v * v).sum()
        pred = l2sum > 0
        l2sum_safe = torch.where(pred, l2sum, torch.ones_like(l2sum))
        l2norm = torch.where(pred, torch.sqrt(l2sum_safe), l2sum)
        cn = torch.as_tensor(clip_norm, dtype=v.dtype)
        return (v * cn) / torch.maximum(l2norm, cn)
    return Tensor(fn, [t], t._shape, dtype=t.dtype)


class _nn:
    @staticmethod
    def relu(x, name=None):
        return Tensor(lambda run: torch.relu(x._eval(run)), [x], x._shape, dtype=x.dtype)

    @staticmethod
    def dropout(x, keep_prob=None, noise_shape=None, seed=None, name=None, rate=None):
        """nn_ops.dropout (TF1): random_tensor = keep_prob + random_uniform(shape(x)); binary = floor(random_tensor);
        ret = (x / keep_prob) * binary.  keep_prob is a tensor here, so the op always runs (no short cut at 1.0)."""
        if keep_prob is None:
            keep_prob = 1.0 - rate
        idx = len(_G.dropout_ops)

        def fn(run):
            v = x._eval(run)
            kp = _val(keep_prob, run, like=v).to(v.dtype)
            masks = _G.dropout_masks
            if masks is not None and masks[idx] is not None:
                binary = torch.as_tensor(np.asarray(masks[idx]), dtype=v.dtype)
            else:
                u = torch.as_tensor(_G.rng.uniform(size=tuple(v.shape)), dtype=v.dtype)
                binary = torch.floor(kp + u)
            return (v / kp) * binary
        t = Tensor(fn, [x] + ([keep_prob] if isinstance(keep_prob, Tensor) else []), x._shape, dtype=x.dtype)
        _G.dropout_ops.append(t)
        return t


nn = _nn()


# --------------------------------------------------------------------------------------------------- batch norm
def _moments(v):
    """nn_impl.moments(x, axes=[0]): mean = reduce_mean(x); variance = reduce_mean(squared_difference(x,
    stop_gradient(mean))) - the biased (population) variance."""
    mean = v.mean(dim=0)
    var = torch.square(v - mean.detach()).mean(dim=0)
    return mean, var


def _batch_normalization(v, mean, var, offset, scale, eps):
    """nn_impl.batch_normalization: inv = rsqrt(variance + eps) * scale; x*inv + (offset - mean*inv)."""
    inv = torch.rsqrt(var + eps) * scale
    return v * inv + (offset - mean * inv)


class _BatchNorm:
    """keras/layers/normalization.py::BatchNormalizationBase, axis=-1, momentum=0.99, epsilon=1e-3, center, scale,
    non-fused (rank-2 inputs).  Variables: gamma (ones), beta (zeros), moving_mean (zeros), moving_variance (ones)."""

    def __init__(self, name, momentum=0.99, epsilon=1e-3):
        self.name, self.momentum, self.epsilon = name, momentum, epsilon
        self.built = False

    def build(self, n, dtype):
        with variable_scope(self.name):
            self.gamma = Variable(np.ones(n), True, dtype, "gamma")
            self.beta = Variable(np.zeros(n), True, dtype, "beta")
            self.moving_mean = Variable(np.zeros(n), False, dtype, "moving_mean")
            self.moving_variance = Variable(np.ones(n), False, dtype, "moving_variance")
        self.built = True

    def _is_training(self, training, run):
        if isinstance(training, Tensor):
            return builtins_bool(training._eval(run))
        return builtins_bool(training)

    def apply(self, x, training, graph_mode):
        if not self.built:
            self.build(x._shape[-1], x.dtype or float32)
        layer = self
        stats = Tensor(lambda run: _moments(x._eval(run)), [x])

        def fn(run):
            v = x._eval(run)
            if layer._is_training(training, run):
                mean, var = stats._eval(run)
                if not graph_mode:                       # TF2 eager: the layer updates its moving statistics in call()
                    layer._update(mean, var)
            else:
                mean, var = layer.moving_mean.value, layer.moving_variance.value
            return _batch_normalization(v, mean, var, layer.beta.value, layer.gamma.value, layer.epsilon)
        out = Tensor(fn, [x, stats, self.gamma, self.beta, self.moving_mean, self.moving_variance], x._shape,
                     dtype=x.dtype)
        if graph_mode:
            # smart_cond(training, assign_moving_average, identity): one update op per moving statistic, collected in
            # UPDATE_OPS; they run only if something depends on them (the reference ties them to the train op).
            def upd(run):
                if layer._is_training(training, run):
                    mean, var = stats._eval(run)
                    layer._update(mean, var)
            _G.collections[GraphKeys.UPDATE_OPS].append(Tensor(upd, [stats]))
        return out

    def _update(self, mean, var):
        """assign_moving_average: variable -= (variable - value) * (1 - momentum); the variance that goes in is the
        biased batch variance of nn.moments (no Bessel correction on the non-fused path)."""
        decay = 1.0 - self.momentum
        with torch.no_grad():
            self.moving_mean.value -= (self.moving_mean.value - mean.detach()) * decay
            self.moving_variance.value -= (self.moving_variance.value - var.detach()) * decay


class _layers:
    @staticmethod
    def batch_normalization(inputs, training=False, name=None, momentum=0.99, epsilon=1e-3):
        return _BatchNorm(name or "batch_normalization", momentum, epsilon).apply(inputs, training, graph_mode=True)


layers = _layers()


# --------------------------------------------------------------------------------------------------- training
class _AdamOptimizer:
    """training/adam.py: slots m, v (zeros), beta1_power/beta2_power start at beta1/beta2 and are multiplied after
    every apply; ApplyAdam: lr_t = lr*sqrt(1-beta2_power)/(1-beta1_power); m += (g-m)*(1-beta1);
    v += (g*g-v)*(1-beta2); var -= lr_t*m/(sqrt(v)+epsilon)."""

    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8):
        self.lr, self.b1, self.b2, self.eps = learning_rate, beta1, beta2, epsilon
        self.b1p, self.b2p = beta1, beta2
        self.m, self.v = {}, {}

    def compute_gradients(self, loss, var_list=None):
        vs_ = trainable_variables() if var_list is None else list(var_list)
        reach = _reachable(loss)
        all_g = Tensor(lambda run: torch.autograd.grad(loss._eval(run), [v.value for v in vs_ if id(v) in reach],
                                                       retain_graph=True, allow_unused=True), [loss])
        out, k = [], 0
        for v in vs_:
            if id(v) not in reach:
                out.append((None, v))
                continue
            out.append((Tensor((lambda kk: lambda run: all_g._eval(run)[kk])(k), [all_g], v._shape, dtype=v.dtype), v))
            k += 1
        return out

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        gv = [(g, v) for g, v in grads_and_vars if g is not None]

        def fn(run):
            lr = _val(self.lr, run)
            grads = [g._eval(run) for g, _ in gv]
            with torch.no_grad():
                for g, (_, v) in zip(grads, gv):
                    if id(v) not in self.m:
                        self.m[id(v)] = torch.zeros_like(v.value)
                        self.v[id(v)] = torch.zeros_like(v.value)
                    lr_t = lr.to(v.value.dtype) * np.sqrt(1.0 - self.b2p) / (1.0 - self.b1p)
                    m, vv = self.m[id(v)], self.v[id(v)]
                    m += (g - m) * (1.0 - self.b1)
                    vv += (g * g - vv) * (1.0 - self.b2)
                    v.value -= lr_t * m / (torch.sqrt(vv) + self.eps)
                self.b1p *= self.b1
                self.b2p *= self.b2
                if global_step is not None:
                    global_step.value += 1
            return None
        return Tensor(fn, [g for g, _ in gv])

    def slot(self, var, which):
        return _to_numpy({"m": self.m, "v": self.v}[which][id(var)])


def _reachable(t):
    seen, stack = set(), [t]
    while stack:
        n = stack.pop()
        if id(n) in seen:
            continue
        seen.add(id(n))
        stack.extend(n._inputs)
    return seen


def _exponential_decay(learning_rate, global_step, decay_steps, decay_rate, staircase=False, name=None):
    """learning_rate_decay.exponential_decay: p = global_step / decay_steps (in the learning rate's dtype, floored if
    staircase); learning_rate * decay_rate ** p."""
    def fn(run):
        lr = _val(learning_rate, run)
        p = _val(global_step, run).to(lr.dtype) / torch.as_tensor(decay_steps, dtype=lr.dtype)
        if staircase:
            p = torch.floor(p)
        return (lr * torch.pow(torch.as_tensor(decay_rate, dtype=lr.dtype), p)).detach()
    return Tensor(fn, [learning_rate, global_step], [], dtype=getattr(learning_rate, "dtype", None))


class _Saver:
    def __init__(self, var_list=None, max_to_keep=5):
        self.var_list = var_list

    def save(self, *a, **k):
        raise NotImplementedError

    restore = save


class _train:
    AdamOptimizer = _AdamOptimizer
    exponential_decay = staticmethod(_exponential_decay)
    Saver = _Saver


train = _train()


# --------------------------------------------------------------------------------------------------- summaries
class _FileWriter:
    def __init__(self, logdir=None, graph=None):
        self.logdir, self.events = logdir, []

    def add_summary(self, summary, global_step=None):
        self.events.append((global_step, summary))

    def add_graph(self, graph):
        pass

    def flush(self):
        pass


class _summary:
    FileWriter = _FileWriter

    @staticmethod
    def scalar(tag, tensor):
        return Tensor(lambda run: {tag: builtins_float(_val(tensor, run).detach())}, [tensor] if isinstance(tensor, Tensor) else ())


summary = _summary()


# --------------------------------------------------------------------------------------------------- session
class ConfigProto:
    def __init__(self, **kw):
        self.kw = kw


class Session:
    def __init__(self, config=None, graph=None):
        self.graph = _G

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def run(self, fetches, feed_dict=None):
        run = _Run(feed_dict)
        return self._fetch(fetches, run)

    def _fetch(self, f, run):
        if isinstance(f, (list, tuple)):
            return [self._fetch(u, run) for u in f]
        return _to_numpy(f._eval(run))


# --------------------------------------------------------------------------------------------------- TF2 surface
class _random:
    truncated_normal = staticmethod(truncated_normal)

    @staticmethod
    def normal(shape, mean=0.0, stddev=1.0, dtype=float32):
        tv = torch.as_tensor(_G.rng.standard_normal(tuple(shape)) * stddev + mean, dtype=_as_dtype(dtype).torch)
        return Tensor(lambda run: tv, (), list(shape))


random = _random()


class _math:
    sqrt = staticmethod(sqrt)


math = _math()


def function(f=None, **kw):
    return f if f is not None else (lambda g: g)


from . import keras  # noqa: E402,F401
