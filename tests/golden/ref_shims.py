"""Import shims that let the UNMODIFIED reference modules under /root/reference run in a
container without gym / TensorFlow / kfac, so that tests/golden/make_golden.py can produce golden
vectors from the reference's own Python code.

  * ``gym``        : just the class skeleton the wrappers subclass (Wrapper, ObservationWrapper,
                     RewardWrapper, Env, spaces.Box/Discrete).  No emulator.
  * ``tensorflow`` : an EAGER restatement of the ~30 TF-1 symbols the hot-path modules touch,
                     evaluated with torch-CPU float64/float32 tensors.  ``tf.placeholder`` returns
                     the value registered for that placeholder name (FEEDS), ``tf.get_variable``
                     returns the injected parameter (VARIABLES) - so building the reference's
                     AtariModel / A2CObjective *is* evaluating them on that data.
  * ``kfac``       : KfacOptimizer / LayerCollection skeleton.  The K-FAC arithmetic itself is NOT
                     available; the skeleton KfacOptimizer RECORDS which of its pieces the
                     reference's subclass (kfac_utils.ColdStartPeriodicInvUpdateKfacOpt) runs
                     and in which order (covariance thunks, inverse thunks, the base
                     apply_gradients, which increments global_step like kfac's does), so that
                     the schedule golden (schedule.npz) comes from the reference class itself.

Only used by make_golden.py (run in the build container; /root/reference does not exist on the
GPU box).  The TF-op semantics restated here (conv2d NHWC/HWIO VALID cross-correlation, softmax
cross-entropy log_prob, Categorical entropy) are the documented TF-1 ones.
"""
import contextlib
import sys
import types

import numpy as np
import torch

FEEDS = {}        # placeholder name -> np.ndarray
VARIABLES = {}    # "scope/name" -> torch tensor (requires_grad)
_SCOPE = []
COMPUTE_DTYPE = torch.float64


# ----------------------------------------------------------------------------- gym
def _make_gym():
    gym = types.ModuleType("gym")

    class Env(object):
        observation_space = None
        action_space = None
        metadata = {}
        reward_range = (-float("inf"), float("inf"))

        def step(self, action):
            raise NotImplementedError

        def reset(self, **kwargs):
            raise NotImplementedError

        def close(self):
            pass

    class Wrapper(Env):
        def __init__(self, env):
            self.env = env
            self.observation_space = env.observation_space
            self.action_space = env.action_space

        def step(self, action):
            return self.env.step(action)

        def reset(self, **kwargs):
            return self.env.reset(**kwargs)

        @property
        def unwrapped(self):
            return getattr(self.env, "unwrapped", self.env)

    class ObservationWrapper(Wrapper):
        def step(self, action):
            obs, reward, done, info = self.env.step(action)
            return self.observation(obs), reward, done, info

        def reset(self, **kwargs):
            return self.observation(self.env.reset(**kwargs))

    class RewardWrapper(Wrapper):
        def step(self, action):
            obs, reward, done, info = self.env.step(action)
            return obs, self.reward(reward), done, info

    spaces = types.ModuleType("gym.spaces")

    class Space(object):
        pass

    class Box(Space):
        def __init__(self, low, high, shape=None, dtype=np.float32):
            if shape is not None:
                low = np.full(shape, low, dtype=dtype)
                high = np.full(shape, high, dtype=dtype)
            self.low = np.asarray(low, dtype=dtype)
            self.high = np.asarray(high, dtype=dtype)
            self.shape = self.low.shape
            self.dtype = np.dtype(dtype)

    class Discrete(Space):
        def __init__(self, n):
            self.n = n
            self.shape = ()
            self.dtype = np.dtype(np.int64)

    spaces.Space, spaces.Box, spaces.Discrete = Space, Box, Discrete
    gym.Env, gym.Wrapper, gym.ObservationWrapper, gym.RewardWrapper = Env, Wrapper, ObservationWrapper, RewardWrapper
    gym.spaces = spaces
    gym.make = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("no emulator in the shim"))
    return gym, spaces


# ----------------------------------------------------------------------------- tensorflow (eager, torch-backed)
def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


class _DType(object):
    def __init__(self, name, np_dtype, torch_dtype):
        self.name, self.np, self.torch = name, np.dtype(np_dtype), torch_dtype

    def as_numpy_dtype(self):
        return self.np


def _make_tf():
    tf = types.ModuleType("tensorflow")
    tf.float32 = _DType("float32", np.float32, None)        # float -> COMPUTE_DTYPE
    tf.int32 = _DType("int32", np.int32, torch.int64)
    tf.bool = _DType("bool", np.bool_, torch.bool)
    tf.uint8 = _DType("uint8", np.uint8, torch.uint8)

    def as_dtype(d):
        d = np.dtype(d)
        return {np.dtype(np.uint8): tf.uint8, np.dtype(np.float32): tf.float32,
                np.dtype(np.int32): tf.int32, np.dtype(np.bool_): tf.bool}[d]
    tf.as_dtype = as_dtype

    def _torch_dtype(d):
        if d is tf.float32:
            return COMPUTE_DTYPE
        return d.torch

    @contextlib.contextmanager
    def name_scope(*args, **kwargs):
        yield
    tf.name_scope = name_scope

    @contextlib.contextmanager
    def variable_scope(name_or_scope=None, default_name=None, reuse=None, **kwargs):
        name = name_or_scope if name_or_scope is not None else default_name
        _SCOPE.append(name)
        try:
            yield
        finally:
            _SCOPE.pop()
    tf.variable_scope = variable_scope

    def placeholder(dtype, shape=None, name=None):
        value = FEEDS[name]
        t = torch.as_tensor(np.asarray(value))
        if dtype is tf.float32:
            t = t.to(COMPUTE_DTYPE)
        return t
    tf.placeholder = placeholder

    def get_variable(name, shape=None, dtype=None, initializer=None):
        # scope[0] is the model scope ('AtariModel'); variables are keyed by layer scope + name
        key = "/".join([s for s in _SCOPE[1:]] + [name])
        v = VARIABLES[key]
        assert tuple(v.shape) == tuple(shape), (key, v.shape, shape)
        return v
    tf.get_variable = get_variable
    tf.orthogonal_initializer = lambda gain=1.0, dtype=None: ("orthogonal", gain)
    tf.zeros_initializer = lambda dtype=None: ("zeros",)

    tf.shape = lambda x: list(x.shape)
    tf.cast = lambda x, dtype, name=None: _t(x).to(_torch_dtype(dtype))
    tf.reshape = lambda x, shape, name=None: _t(x).reshape([int(s) for s in shape])
    tf.stop_gradient = lambda x, name=None: _t(x).detach()
    tf.expand_dims = lambda x, axis, name=None: _t(x).unsqueeze(axis)
    tf.squeeze = lambda x, axis=None, name=None: _t(x).squeeze(axis) if axis is not None else _t(x).squeeze()
    tf.matmul = lambda a, b, name=None: torch.matmul(_t(a), _t(b).to(_t(a).dtype))
    tf.reduce_mean = lambda x, axis=None, name=None: _t(x).mean() if axis is None else _t(x).mean(axis)
    tf.square = lambda x, name=None: _t(x) * _t(x)
    tf.group = lambda ops, name=None: ops
    tf.no_op = lambda name=None: None

    # --- eager control flow (kfac_utils.py:38-53): everything executes in program order, so a
    # `control_dependencies` block is simply "after what ran before"; predicates read the CURRENT value of a StepVariable
    @contextlib.contextmanager
    def control_dependencies(ops):
        yield
    tf.control_dependencies = control_dependencies
    tf.less = lambda a, b, name=None: _val(a) < _val(b)
    tf.greater = lambda a, b, name=None: _val(a) > _val(b)
    tf.equal = lambda a, b, name=None: _val(a) == _val(b)
    tf.mod = lambda a, b, name=None: _val(a) % _val(b)
    tf.logical_and = lambda a, b, name=None: bool(a) and bool(b)

    def cond(pred, true_fn, false_fn, name=None):
        return true_fn() if bool(pred) else false_fn()
    tf.cond = cond

    def py_func(fn, inputs, dtype, stateful=True, name=None):
        # objectives.py:198,213: inputs reach the Python function as NumPy values of their graph
        # dtype (terminals: bool; discount_factor: a Python float converted to a float32 tensor)
        args = []
        for x in inputs:
            if isinstance(x, torch.Tensor):
                args.append(x.detach().cpu().numpy())
            elif isinstance(x, float):
                args.append(np.float32(x))
            else:
                args.append(np.asarray(x))
        out = fn(*args)
        return torch.as_tensor(np.array(out, dtype=dtype.np, copy=True, order="C").reshape(np.shape(out)).copy()).to(COMPUTE_DTYPE)
    tf.py_func = py_func

    nn = types.ModuleType("tensorflow.nn")

    def conv2d(input, filter, strides, padding, data_format="NHWC", name=None):
        assert data_format == "NHWC" and padding == "VALID" and strides[0] == 1 and strides[3] == 1
        x = _t(input).permute(0, 3, 1, 2)                      # NHWC -> NCHW
        w = _t(filter).permute(3, 2, 0, 1)                     # HWIO -> OIHW (cross-correlation, no flip)
        y = torch.nn.functional.conv2d(x, w, stride=(strides[1], strides[2]))
        return y.permute(0, 2, 3, 1)
    nn.conv2d = conv2d
    nn.relu = lambda x, name=None: torch.relu(_t(x))
    tf.nn = nn

    distributions = types.ModuleType("tensorflow.distributions")

    class Categorical(object):
        def __init__(self, logits, name=None):
            self.logits = logits
            self._logp = torch.log_softmax(logits, dim=-1)

        def sample(self, sample_shape=(), seed=None, name=None):
            g = torch.Generator().manual_seed(0 if seed is None else int(seed))
            flat = torch.softmax(self.logits.detach().reshape(-1, self.logits.shape[-1]), -1)
            return torch.multinomial(flat, 1, generator=g).reshape(self.logits.shape[:-1])

        def mode(self, name=None):
            return torch.argmax(self.logits, dim=-1)

        def entropy(self, name=None):
            return -(torch.exp(self._logp) * self._logp).sum(-1)

        def log_prob(self, value, name=None):
            return torch.gather(self._logp, -1, _t(value).long().unsqueeze(-1)).squeeze(-1)
    distributions.Categorical = Categorical
    tf.distributions = distributions

    train = types.ModuleType("tensorflow.train")

    class Optimizer(object):
        def __init__(self, use_locking=False, name=None):
            pass
    train.Optimizer = Optimizer
    tf.train = train
    return tf


class StepVariable(object):
    """A global_step stand-in: a mutable integer read at the moment an op uses it."""

    def __init__(self, value=0):
        self.value = int(value)

    def assign_add(self, n):
        self.value += int(n)
        return self.value

    def __sub__(self, other):
        return self.value - _val(other)

    def __int__(self):
        return self.value


def _val(x):
    return x.value if isinstance(x, StepVariable) else x


EVENTS = []       # (name, global_step value when it ran) recorded by the kfac / optimizer skeletons below


class RecordingOptimizer(object):
    """Stands in for the cold optimizer (tf.train.MomentumOptimizer behind nn.ClipGlobalNormOptimizer): like every
    tf.train.Optimizer.apply_gradients it increments the global_step it is handed."""

    def __init__(self, name="cold"):
        self.name = name

    def apply_gradients(self, grads_and_vars, global_step=None, name=None):
        EVENTS.append((self.name, _val(global_step)))
        if global_step is not None:
            global_step.assign_add(1)
        return self.name


def _make_kfac():
    kfac = types.ModuleType("kfac")

    class KfacOptimizer(object):
        def __init__(self, **kwargs):
            self.kwargs = kwargs

        def make_vars_and_create_op_thunks(self):
            def cov():
                EVENTS.append(("cov", None))
                return "cov"

            def inv():
                EVENTS.append(("inv", None))
                return "inv"
            return [cov], [inv]

        def apply_gradients(self, grads_and_vars, global_step=None, name=None):
            EVENTS.append(("kfac_apply", _val(global_step)))
            if global_step is not None:
                global_step.assign_add(1)
            return "kfac_apply"

    class LayerCollection(object):
        """Records registrations so the golden file can list what the reference registers."""

        def __init__(self):
            self.calls = []

        def _rec(self, kind, **kw):
            self.calls.append((kind, kw))

        def register_conv2d(self, params, strides, padding, inputs, outputs):
            self._rec("conv2d", params=params, strides=strides, padding=padding, inputs=inputs, outputs=outputs)

        def register_fully_connected(self, params, inputs, outputs):
            self._rec("fully_connected", params=params, inputs=inputs, outputs=outputs)

        def register_categorical_predictive_distribution(self, logits, seed=None):
            self._rec("categorical", logits=logits, seed=seed)

        def register_normal_predictive_distribution(self, mean, var=0.5, seed=None):
            self._rec("normal", mean=mean, var=var, seed=seed)
    kfac.KfacOptimizer, kfac.LayerCollection = KfacOptimizer, LayerCollection
    return kfac


class _Shape(list):
    def as_list(self):
        return list(self)


def install(reference_root="/root/reference"):
    """Put the shims in sys.modules and the reference on sys.path.  Returns (gym, tf, kfac)."""
    torch.Tensor.get_shape = lambda self: _Shape(int(d) for d in self.shape)     # nn.py:125
    gym, spaces = _make_gym()
    tf = _make_tf()
    kfac = _make_kfac()
    sys.modules["gym"] = gym
    sys.modules["gym.spaces"] = spaces
    sys.modules["tensorflow"] = tf
    sys.modules["kfac"] = kfac
    if reference_root not in sys.path:
        sys.path.insert(0, reference_root)
    return gym, tf, kfac
