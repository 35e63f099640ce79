"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (mpg_b200/).

A minimal stand-in for the third-party packages the reference imports at module level
(`tensorflow`, `tensorflow_probability`, `gym`, `matplotlib`) so that the reference's OWN,
UNMODIFIED Python (read from /root/reference at fixture-generation time, in the build container
only) can be executed to produce golden vectors (tests/golden/make_golden.py).

TensorFlow itself is not installable here (no network, no wheel), so the tensor *kernels* below
are PyTorch CPU ops; the *algorithm* that is executed -- dynamics, rewards, rollout loop, Q
bootstrap, lambda-weights, GradientTape.gradient call sites, clip_by_global_norm -- is the
reference's source, line for line.  Only the subset of the TF API touched by
  envs_and_models/{path_tracking_env,inverted_pendulum_model,inverted_double_pendulum_model}.py
  model.py, policy.py, preprocessor.py, learners/{mpg_learner,nadp,ampc}.py
is provided.  Semantics that matter for parity and how they are mirrored:
  * tf.clip_by_value gradient: pass-through on the closed interval  -> torch.clamp (same mask)
  * tf.where gradient: selects branch gradient                      -> torch.where
  * tf.keras Dense: y = act(x @ kernel + bias), kernel (in,out)     -> same, kernel (in,out)
  * elu: x>0 ? x : exp(x)-1                                         -> torch.nn.functional.elu
  * tf.clip_by_global_norm(t, c): t*c/max(norm,c)                   -> same formula
  * tfd.Normal(loc, scale).sample(): RNG is *not* part of the contract; the shim draws
    loc + scale * eps where eps comes from a caller-installed queue (`set_noise_source`) so
    the oracle and the CUDA path can be fed the identical noise tensor.
`DTYPE` (float32 = reference precision, float64 = "truth") is switchable with `set_dtype`.
"""
import contextlib
import sys
import types

import numpy as np
import torch

DTYPE = torch.float32
_noise_source = None


def set_dtype(dt):
    global DTYPE
    DTYPE = dt


def set_noise_source(fn):
    """fn(shape) -> standard-normal torch tensor; called once per tfd.Normal(...).sample()."""
    global _noise_source
    _noise_source = fn


def _t(x):
    if isinstance(x, torch.Tensor):
        return x if (x.dtype == DTYPE or not x.is_floating_point()) else x.to(DTYPE)
    if isinstance(x, Variable):
        return x.value
    if isinstance(x, (list, tuple)) and len(x) > 0 and isinstance(x[0], (torch.Tensor, Variable)):
        return torch.stack([_t(v) for v in x])
    a = np.asarray(x)
    if a.dtype.kind in 'iub' and a.dtype != np.bool_:
        return torch.as_tensor(a).to(DTYPE)
    if a.dtype == np.bool_:
        return torch.as_tensor(a)
    return torch.as_tensor(a.astype(np.float64)).to(DTYPE)


class Variable:
    """tf.Variable: a leaf tensor; arithmetic goes through .value"""

    def __init__(self, init, dtype=None, trainable=True, name=None):
        self.value = _t(init).clone().detach().requires_grad_(bool(trainable))
        self.name = name

    def assign(self, v):
        with torch.no_grad():
            self.value.copy_(_t(v))

    def numpy(self):
        return self.value.detach().cpu().numpy()

    @property
    def shape(self):
        return tuple(self.value.shape)

    # arithmetic used by the Polyak target updates (policy.py:158-171): tau * source + (1 - tau) * target
    def __mul__(self, o):
        return self.value.detach() * o

    __rmul__ = __mul__

    def __add__(self, o):
        return self.value.detach() + o

    __radd__ = __add__


# make `tensor.numpy()` legal on tensors that carry grad history (TF eager tensors allow it)
_orig_numpy = torch.Tensor.numpy


def _numpy_detached(self, *a, **k):
    return _orig_numpy(self.detach(), *a, **k)


def _patch_numpy_operands():
    """TF eager tensors accept numpy operands on either side of + - * /; torch does not."""
    for name in ('__mul__', '__rmul__', '__add__', '__radd__', '__sub__', '__rsub__',
                 '__truediv__', '__rtruediv__'):
        orig = getattr(torch.Tensor, name)

        def wrapped(self, other, _orig=orig):
            if isinstance(other, np.ndarray):
                other = _t(other)
                if self.is_floating_point() and self.dtype != DTYPE:
                    self = self.to(DTYPE)
            return _orig(self, other)
        setattr(torch.Tensor, name, wrapped)
    # TF tensors are immutable: `x += y` rebinds x to a NEW tensor (the reference relies on this,
    # e.g. rewards_sum_tile in learners/nadp.py:146 is appended to a list and then `+=`-ed).
    for iname, name in (('__iadd__', '__add__'), ('__isub__', '__sub__'), ('__imul__', '__mul__'),
                        ('__itruediv__', '__truediv__')):
        setattr(torch.Tensor, iname, lambda self, other, _n=name: getattr(self, _n)(other))


class GradientTape:
    def __init__(self, persistent=False):
        self.persistent = persistent

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def gradient(self, target, sources):
        srcs = [s.value if isinstance(s, Variable) else s for s in sources]
        grads = torch.autograd.grad(target, srcs, retain_graph=True, allow_unused=True)
        # TF returns None for unconnected sources; Keras Dense params here are always connected
        # except the unused log-std columns, which are still *connected* (zero-valued grads).
        return [g if g is not None else torch.zeros_like(s) for g, s in zip(grads, srcs)]


def _clip_by_global_norm(t_list, clip_norm):
    t_list = [_t(t) for t in t_list]
    norm = torch.sqrt(sum((t * t).sum() for t in t_list))
    scale = clip_norm * torch.minimum(1.0 / norm, torch.tensor(1.0 / clip_norm, dtype=norm.dtype))
    return [t * scale for t in t_list], norm


class _Dense:
    def __init__(self, units, activation=None, kernel_initializer=None, bias_initializer=None, dtype=None):
        self.units = units
        self.activation = activation
        self.kernel = None
        self.bias = None

    def build(self, in_dim):
        g = torch.Generator().manual_seed(in_dim * 1000 + self.units)
        self.kernel = Variable(torch.randn(in_dim, self.units, generator=g) / np.sqrt(in_dim))
        self.bias = Variable(torch.zeros(self.units))
        return self.units

    def __call__(self, x):
        y = _t(x) @ self.kernel.value + self.bias.value
        act = self.activation
        if act in (None, 'linear'):
            return y
        if act == 'elu':
            return torch.nn.functional.elu(y)
        if act == 'tanh':
            return torch.tanh(y)
        if act == 'relu':
            return torch.relu(y)
        raise NotImplementedError(act)

    @property
    def trainable_weights(self):
        return [self.kernel, self.bias]


class _Sequential:
    def __init__(self, layers=()):
        self.layers = list(layers)

    def build(self, in_dim):
        for l in self.layers:
            in_dim = l.build(in_dim)
        return in_dim

    def __call__(self, x):
        for l in self.layers:
            x = l(x)
        return x

    @property
    def trainable_weights(self):
        return [w for l in self.layers for w in l.trainable_weights]


class _KerasModel:
    """keras.Model: attribute-order layer tracking, build(), get/set_weights, trainable_weights."""

    def __init__(self, name=None, **kw):
        object.__setattr__(self, '_tracked', [])
        self.name = name

    def __setattr__(self, k, v):
        if isinstance(v, (_Dense, _Sequential, Variable)):
            self._tracked.append(v)
        object.__setattr__(self, k, v)

    def build(self, input_shape):
        d = input_shape[-1]
        for l in self._tracked:
            if not isinstance(l, Variable):
                d = l.build(d)

    def __call__(self, x, **kw):
        return self.call(_t(x), **kw)

    @property
    def trainable_weights(self):
        out = []
        for l in self._tracked:
            out += [l] if isinstance(l, Variable) else l.trainable_weights
        return out

    def get_weights(self):
        return [w.numpy().copy() for w in self.trainable_weights]

    def set_weights(self, ws):
        tw = self.trainable_weights
        assert len(tw) == len(ws)
        for v, w in zip(tw, ws):
            assert tuple(v.shape) == tuple(np.shape(w)), (v.shape, np.shape(w))
            v.assign(w)


class _PolynomialDecay:
    """tf.keras.optimizers.schedules.PolynomialDecay(initial_learning_rate, decay_steps, end_learning_rate),
    power = 1, cycle = False (published Keras semantics; TF itself is not installable here)."""

    def __init__(self, initial, decay_steps, end=0.0001, power=1.0):
        self.initial, self.decay_steps, self.end, self.power = initial, decay_steps, end, power

    def __call__(self, step):
        frac = min(float(step), float(self.decay_steps)) / float(self.decay_steps)
        return (self.initial - self.end) * (1.0 - frac) ** self.power + self.end


class _Adam:
    """tf.keras.optimizers.Adam (OptimizerV2, TF 2.2-2.4): beta_1 0.9, beta_2 0.999, epsilon 1e-7, no amsgrad.
    lr_t = lr(iterations) * sqrt(1 - b2^t) / (1 - b1^t), t = iterations + 1;
    m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2; var -= lr_t * m / (sqrt(v) + eps)."""

    def __init__(self, lr=None, name=None):
        self._name, self.lr, self.iterations, self.slots = name, lr, 0, {}

    def apply_gradients(self, gv):
        t = self.iterations + 1
        lr = self.lr(self.iterations) if callable(self.lr) else self.lr
        lr_t = lr * np.sqrt(1.0 - 0.999 ** t) / (1.0 - 0.9 ** t)
        for g, var in gv:
            g = _t(g).detach()
            m, v = self.slots.get(id(var), (torch.zeros_like(g), torch.zeros_like(g)))
            m = 0.9 * m + (1.0 - 0.9) * g
            v = 0.999 * v + (1.0 - 0.999) * g * g
            self.slots[id(var)] = (m, v)
            var.assign(var.value.detach() - lr_t * m / (torch.sqrt(v) + 1e-7))
        self.iterations = t


class _Normal:
    def __init__(self, loc, scale):
        self.loc, self.scale = _t(loc), scale

    def sample(self):
        assert _noise_source is not None, 'install a noise source with tf_shim.set_noise_source'
        eps = _noise_source(tuple(self.loc.shape)).to(DTYPE)
        return self.loc + self.scale * eps


def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


@contextlib.contextmanager
def _name_scope(name):
    yield name


def _convert(x, dtype=None):
    return _t(x)


def _reduce(fn):
    def f(x, axis=None):
        x = _t(x)
        return fn(x) if axis is None else fn(x, dim=axis)
    return f


def _reduce_min(x, axis=None):
    x = _t(x)
    return x.min() if axis is None else x.min(dim=axis).values


def _variance(x, axis=None):
    x = _t(x)
    return x.var(unbiased=False) if axis is None else x.var(dim=axis, unbiased=False)


def _split(x, num_or_size_splits, axis=-1):
    x = _t(x)
    return torch.chunk(x, num_or_size_splits, dim=axis)


def _squeeze(x, axis=None):
    return _t(x).squeeze(axis) if axis is not None else _t(x).squeeze()


def _where(c, a, b):
    if not isinstance(c, torch.Tensor):
        c = torch.as_tensor(np.asarray(c))
    return torch.where(c, _t(a), _t(b))


def _clip_by_value(x, lo, hi):
    return torch.clamp(_t(x), min=float(lo), max=float(hi))


def install():
    """Register the fake modules in sys.modules (idempotent)."""
    if getattr(sys.modules.get('tensorflow'), '_is_mpg_shim', False):
        return
    torch.Tensor.numpy = _numpy_detached
    if not hasattr(np, 'int'):
        np.int = int    # removed numpy aliases the reference still uses (path_tracking_env.py:371, dummy_vec_env.py:22)
    if not hasattr(np, 'bool'):
        np.bool = bool
    _patch_numpy_operands()
    noop = lambda *a, **k: None
    tf = _mod(
        'tensorflow', _is_mpg_shim=True,
        float32='float32', int32='int32',
        Variable=Variable, Module=object, GradientTape=GradientTape,
        function=lambda f: f, name_scope=_name_scope,
        convert_to_tensor=_convert, constant=_convert,
        zeros=lambda shape, dtype=None: torch.zeros(shape, dtype=DTYPE),
        ones=lambda shape, dtype=None: torch.ones(shape, dtype=DTYPE),
        zeros_like=lambda x, dtype=None: torch.zeros_like(_t(x)),
        ones_like=lambda x, dtype=None: torch.ones_like(_t(x)),
        stack=lambda xs, axis=0: torch.stack([_t(x) for x in xs], dim=axis),
        concat=lambda xs, axis=0: torch.cat([_t(x) for x in xs], dim=axis),
        reshape=lambda x, shape: _t(x).reshape(tuple(shape)),
        squeeze=_squeeze,
        tile=lambda x, mult: _t(x).repeat(*mult),
        where=_where, split=_split,
        sqrt=lambda x: torch.sqrt(_t(x)), square=lambda x: _t(x) * _t(x),
        sin=lambda x: torch.sin(_t(x)), cos=lambda x: torch.cos(_t(x)),
        atan=lambda x: torch.atan(_t(x)), atan2=lambda y, x: torch.atan2(_t(y), _t(x)),
        abs=lambda x: torch.abs(_t(x)), tanh=lambda x: torch.tanh(_t(x)), exp=lambda x: torch.exp(_t(x)),
        pow=lambda x, y: torch.pow(_t(x), y),
        matmul=lambda a, b: _t(a) @ _t(b),
        clip_by_value=_clip_by_value, clip_by_global_norm=_clip_by_global_norm,
        reduce_mean=_reduce(torch.mean), reduce_sum=_reduce(torch.sum), reduce_min=_reduce_min,
        stop_gradient=lambda x: _t(x).detach(),
    )
    tf.nn = _mod('tensorflow.nn', softmax=lambda x: torch.softmax(_t(x), dim=-1))
    tf.math = _mod('tensorflow.math', reduce_variance=_variance)
    tf.linalg = _mod('tensorflow.linalg', inv=lambda m: torch.linalg.inv(_t(m)))
    tf.config = _mod('tensorflow.config',
                     experimental=_mod('e', set_visible_devices=noop),
                     threading=_mod('t', set_inter_op_parallelism_threads=noop,
                                    set_intra_op_parallelism_threads=noop))
    tf.summary = _mod('tensorflow.summary', trace_on=noop, trace_export=noop, scalar=noop)
    tf.train = _mod('tensorflow.train', Checkpoint=object)
    inits = _mod('tensorflow.keras.initializers', Orthogonal=lambda *a, **k: None, Constant=lambda *a, **k: None)
    sched = _mod('tensorflow.keras.optimizers.schedules', PolynomialDecay=_PolynomialDecay)
    opts = _mod('tensorflow.keras.optimizers', Adam=_Adam, schedules=sched)
    layers = _mod('tensorflow.keras.layers', Dense=_Dense)
    keras = _mod('tensorflow.keras', Model=_KerasModel, Sequential=_Sequential, layers=layers,
                 initializers=inits, optimizers=opts)
    tf.keras = keras
    tfp = _mod('tensorflow_probability', distributions=_mod('tfd', Normal=_Normal),
               bijectors=_mod('tfb'))

    class _Wrapper:
        def __init__(self, env):
            self.env = env

    class _Env:
        pass

    class _DummyEnv:
        def __init__(self, *a, **k):
            self.num_agent = k.get('num_agent', 1)

        def reset(self, **k):
            return np.zeros(4, np.float32)

    class _Box:
        def __init__(self, low=None, high=None, dtype=np.float32, **k):
            self.low, self.high = np.asarray(low, dtype), np.asarray(high, dtype)

    def _make(env_id=None, *a, **k):
        # the reference relies on a gym registration of 'PathTracking-v0' that is not in its repository; the class
        # it must point to is envs_and_models.path_tracking_env.PathTrackingEnv
        if env_id == 'PathTracking-v0':
            import importlib
            return importlib.import_module('envs_and_models.path_tracking_env').PathTrackingEnv(**k)
        return _DummyEnv(*a, **k)

    gym = _mod('gym', make=_make,
               Env=_Env, core=_mod('gym.core', Wrapper=_Wrapper),
               spaces=_mod('gym.spaces', Box=_Box))
    gym.utils = _mod('gym.utils', seeding=_mod('gym.utils.seeding', np_random=lambda s=None: (np.random.RandomState(s), s)))
    plt = _mod('matplotlib.pyplot', ion=noop, cla=noop)
    mpl = _mod('matplotlib', pyplot=plt)
    for name, m in {
        'tensorflow': tf, 'tensorflow.keras': keras, 'tensorflow.keras.layers': layers,
        'tensorflow.keras.optimizers': opts, 'tensorflow.keras.optimizers.schedules': sched,
        'tensorflow.keras.initializers': inits, 'tensorflow_probability': tfp,
        'gym': gym, 'gym.core': gym.core, 'gym.spaces': gym.spaces, 'gym.utils': gym.utils,
        'gym.utils.seeding': gym.utils.seeding, 'matplotlib': mpl, 'matplotlib.pyplot': plt,
    }.items():
        sys.modules[name] = m
