"""actorcritic/nn.py: parameter initialisers, `linear_decay`, `ClipGlobalNormOptimizer`, plus the two
`tf.train` optimizers the example constructs (a2c_acktr.py:240,250).  The layer functions of the reference
(`conv2d`, `fully_connected`, `flatten`) only exist fused inside the engine's kernels."""
import numpy as np


def _orthogonal(shape, gain, rng):
    rows, cols = int(np.prod(shape[:-1])), int(shape[-1])
    a = rng.standard_normal((max(rows, cols), min(rows, cols)))
    q, r = np.linalg.qr(a)
    q = q * np.sign(np.diag(r))
    if rows < cols:
        q = q.T
    return (gain * q).reshape(shape).astype(np.float32)


def fully_connected_params(input_size, output_size, dtype=np.float32, weights_initializer=None, bias_initializer=None,
                           rng=None, gain=1.0):
    """nn.py:8-33: weights [input_size, output_size], bias [output_size] (zeros)."""
    rng = np.random.default_rng() if rng is None else rng
    return _orthogonal((input_size, output_size), gain, rng).astype(dtype), np.zeros(output_size, dtype)


def conv2d_params(num_input_channels, num_filters, filter_extent, dtype=np.float32, weights_initializer=None,
                  bias_initializer=None, rng=None, gain=1.0):
    """nn.py:55-84: HWIO kernel [k, k, cin, cout], bias [cout]."""
    rng = np.random.default_rng() if rng is None else rng
    shape = (filter_extent, filter_extent, num_input_channels, num_filters)
    return _orthogonal(shape, gain, rng).astype(dtype), np.zeros(num_filters, dtype)


def _fused_only(name):
    def fn(*a, **k):
        raise NotImplementedError("nn.%s is fused into the engine's forward kernels (see envs.atari.model.AtariModel); it "
                                  "has no stand-alone op in this framework" % name)
    fn.__name__ = name
    return fn


fully_connected = _fused_only("fully_connected")
conv2d = _fused_only("conv2d")
flatten = _fused_only("flatten")


class LinearDecay:
    """Value of nn.linear_decay (nn.py:129-156): (start - end) * (1 - min(step, total) / total) + end.  Evaluated on
    the device from the engine's global step (kfac.cu sched_begin_kernel)."""

    def __init__(self, start_value, end_value, step, total_steps):
        self.start_value, self.end_value = float(start_value), float(end_value)
        self.step, self.total_steps = step, float(total_steps)

    def value_at(self, step):
        s = min(float(step), self.total_steps)
        return (self.start_value - self.end_value) * (1.0 - s / self.total_steps) + self.end_value


def linear_decay(start_value, end_value, step, total_steps, name=None):
    return LinearDecay(start_value, end_value, step, total_steps)


class MomentumOptimizer:
    """tf.train.MomentumOptimizer(learning_rate, momentum) record (a2c_acktr.py:240)."""

    def __init__(self, learning_rate, momentum):
        self.learning_rate, self.momentum = learning_rate, momentum


class RMSPropOptimizer:
    """tf.train.RMSPropOptimizer record with the TF-1 defaults (a2c_acktr.py:250)."""

    def __init__(self, learning_rate, decay=0.9, momentum=0.0, epsilon=1e-10):
        if momentum != 0.0:
            raise NotImplementedError("RMSProp momentum is not on the hot path")
        self.learning_rate, self.decay, self.epsilon = learning_rate, decay, epsilon


class ClipGlobalNormOptimizer:
    """nn.py:159-189: clips the gradients by global norm, then applies the wrapped optimizer."""

    def __init__(self, optimizer, clip_norm, name=None):
        self.optimizer, self.clip_norm = optimizer, clip_norm

    def engine_overrides(self):
        opt = self.optimizer
        if not isinstance(opt, RMSPropOptimizer):
            raise NotImplementedError("stand-alone optimizer on the hot path is ClipGlobalNorm(RMSProp) (a2c_acktr.py:250-251)")
        lr = opt.learning_rate
        if not isinstance(lr, LinearDecay):
            lr = LinearDecay(float(lr), float(lr), None, 1.0)
        return dict(acktr=False, lr_start=lr.start_value, lr_end=lr.end_value, lr_decay_steps=lr.total_steps,
                    rms_decay=float(opt.decay), rms_epsilon=float(opt.epsilon), clip_norm=float(self.clip_norm))
