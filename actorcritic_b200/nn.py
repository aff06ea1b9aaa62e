"""actorcritic/nn.py: parameter initialisers, the layer functions `fully_connected`, `conv2d`, `flatten`, `linear_decay`,
`ClipGlobalNormOptimizer`, plus the two `tf.train` optimizers the example constructs (a2c_acktr.py:240,250).

The layer functions take and return fp32 CUDA tensors; the products run on libacx's tcgen05 GEMM (`acx_gemm`, operands split
into three bf16 planes and accumulated over six plane pairs: fp32-class results) or on the implicit-GEMM convolution
(`acx_conv`).  AtariModel's train step does not call them - it runs the same layers fused inside the learner engine - they
exist so that other `ActorCriticModel` subclasses can be built the way the reference builds them (model.py:107-133)."""
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


def _device_f32(x, device=None):
    import torch
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(x))
    if device is None and not x.is_cuda:
        from . import _lib
        raise _lib.AcxError("nn layers operate on CUDA tensors (libacx has no CPU fallback)")
    if device is not None and x.device != device:
        x = x.to(device)
    return x.float() if x.dtype != torch.float32 else x


# noinspection PyShadowingBuiltins
def fully_connected(input, params):
    """nn.py:37-52: `input @ weights + bias` for input [batch, in] (fp32 CUDA tensor), weights [in, out], bias [out]."""
    from . import ops
    x = _device_f32(input)
    weights, bias = params
    w = _device_f32(weights, x.device)
    b = _device_f32(bias, x.device).contiguous()
    if x.dim() != 2 or w.dim() != 2 or x.shape[1] != w.shape[0] or b.numel() != w.shape[1]:
        raise ValueError("fully_connected: input %s, weights %s, bias %s" % (tuple(x.shape), tuple(w.shape), tuple(b.shape)))
    m, k = x.shape
    n = w.shape[1]
    out, _ = ops.gemm(ops.split_planes(x, 3), ops.split_planes(w.t().contiguous(), 3), m, n, k, pairs=ops.PAIRS[6], bias=b)
    return out


# noinspection PyShadowingBuiltins
def conv2d(input, params, stride, padding, impl="gemm"):
    """nn.py:88-110: tf.nn.conv2d(input, weights, (1, stride, stride, 1), padding, 'NHWC') + bias - cross-correlation, HWIO
    weights.  input [batch, h, w, cin] (fp32 CUDA tensor).  impl="gemm": patch rows in (kh, kw, cin) order x the reshaped
    kernel on `acx_gemm`; impl="implicit": `acx_conv` (no patch matrix; VALID, kernel a multiple of the stride,
    stride * cin == 64 - the 4x4/2 and 3x3/1 layers of envs/atari/model.py:180-199)."""
    import torch
    from . import ops
    x = _device_f32(input)
    weights, bias = params
    w = _device_f32(weights, x.device)
    b = _device_f32(bias, x.device).contiguous()
    if x.dim() != 4 or w.dim() != 4 or w.shape[0] != w.shape[1] or w.shape[2] != x.shape[3]:
        raise ValueError("conv2d: input %s (NHWC), weights %s (HWIO)" % (tuple(x.shape), tuple(w.shape)))
    k, cin, cout = int(w.shape[0]), int(w.shape[2]), int(w.shape[3])
    s = int(stride)
    if padding == "SAME":     # TensorFlow's rule: out = ceil(in / stride), the odd pixel of padding goes to the bottom / right
        pads = []
        for size in (x.shape[2], x.shape[1]):
            total = max((-(-size // s) - 1) * s + k - size, 0)
            pads += [total // 2, total - total // 2]
        x = torch.nn.functional.pad(x, (0, 0) + tuple(pads))
    elif padding != "VALID":
        raise ValueError("padding must be 'VALID' or 'SAME'")
    n, h, wd, _ = x.shape
    oh, ow = (h - k) // s + 1, (wd - k) // s + 1
    if impl == "implicit":
        if h != wd:
            raise ValueError("the implicit-GEMM convolution takes square inputs")
        geom = (int(h), cin, k, s, int(oh), cout)
        xp = [p.view(n, h, wd, cin) for p in ops.split_planes(x.reshape(-1, cin), 3)]
        wt = ops.split_planes(w.reshape(k * k * cin, cout).t().contiguous(), 3)
        planes = ops.conv(xp, wt, geom, n, bias=b, relu=False, pairs=ops.PAIRS[6], out_planes=3)
        return planes[0].float() + planes[1].float() + planes[2].float()
    patches = x.unfold(1, k, s).unfold(2, k, s).permute(0, 1, 2, 4, 5, 3).reshape(n * oh * ow, k * k * cin)
    out = fully_connected(patches, (w.reshape(k * k * cin, cout), b))
    return out.view(n, oh, ow, cout)


# noinspection PyShadowingBuiltins
def flatten(input):
    """nn.py:114-126: [batch, d1, ..., dn] -> [batch, d1 * ... * dn] (row-major: (h, w, c) order for NHWC)."""
    return input.reshape(input.shape[0], -1)


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


def standalone_spec(optimizer):
    """(kind, learning-rate LinearDecay, hyper-parameters, clip norm) of an optimizer applied on its own
    (objectives.py:31-54 `optimize_separate`): [ClipGlobalNorm](RMSProp | Momentum)."""
    clip = 3.4e38
    if isinstance(optimizer, ClipGlobalNormOptimizer):
        clip, optimizer = float(optimizer.clip_norm), optimizer.optimizer
    lr = optimizer.learning_rate if hasattr(optimizer, "learning_rate") else None
    if not isinstance(lr, LinearDecay):
        lr = LinearDecay(float(lr), float(lr), None, 1.0)
    if isinstance(optimizer, RMSPropOptimizer):
        return "rmsprop", lr, dict(decay=float(optimizer.decay), epsilon=float(optimizer.epsilon)), clip
    if isinstance(optimizer, MomentumOptimizer):
        return "momentum", lr, dict(momentum=float(optimizer.momentum)), clip
    raise TypeError("unsupported optimizer %r" % (optimizer,))


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
