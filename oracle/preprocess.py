"""Oracle for Atari frame preprocessing and frame stacking (uint8, bit-exact).

Restates, in NumPy:
  * AtariFrameskipWrapper.step's 2-frame max         (wrappers.py:64-65)
  * AtariPreprocessFrameWrapper.observation          (wrappers.py:30-33)
      = cv2.cvtColor(RGB2GRAY) then cv2.resize(84x84, INTER_AREA)
  * FrameStackWrapper.step / reset                   (wrappers.py:224-235)
  * _AutoResetWrapper.step ordering                  (multi_env.py:127-132)

cv2 (OpenCV, the dependency the reference calls) is NOT used here: this is the
restatement of its published algorithm (imgproc color_yuv / resize.cpp
``computeResizeAreaTab`` + ``ResizeArea_``).  tests/test_oracle_preprocess.py
pins it bit-for-bit against the real cv2 and against golden vectors produced by
running the reference's own wrapper classes (tests/golden/make_golden.py).
"""
import math

import numpy as np

RAW_H, RAW_W = 210, 160
OUT_H, OUT_W = 84, 84

# RGB2GRAY fixed point, 15 fractional bits (OpenCV >= 3.4.x / 4.x): R2Y=9798 G2Y=19235 B2Y=3735
_R2Y, _G2Y, _B2Y, _SHIFT = 9798, 19235, 3735, 15


def frame_max(frame_a, frame_b):
    """wrappers.py:64-65  np.amax((frames[-2], frames[-1]), axis=0)."""
    return np.maximum(frame_a, frame_b)


def rgb_to_gray(rgb):
    """cv2.cvtColor(frame, cv2.COLOR_RGB2GRAY) for uint8 (wrappers.py:31)."""
    rgb = np.asarray(rgb)
    assert rgb.dtype == np.uint8 and rgb.shape[-1] == 3
    r = rgb[..., 0].astype(np.int32)
    g = rgb[..., 1].astype(np.int32)
    b = rgb[..., 2].astype(np.int32)
    y = (r * _R2Y + g * _G2Y + b * _B2Y + (1 << (_SHIFT - 1))) >> _SHIFT
    return y.astype(np.uint8)


def area_tab(ssize, dsize):
    """OpenCV computeResizeAreaTab for one axis, cn=1.

    Returns a list of (dst_index, src_index, float32 weight) in table order.
    """
    scale = ssize / dsize  # double
    tab = []
    for d in range(dsize):
        fsx1 = d * scale
        fsx2 = fsx1 + scale
        cell = min(scale, ssize - fsx1)
        sx1 = math.ceil(fsx1)
        sx2 = math.floor(fsx2)
        sx2 = min(sx2, ssize - 1)
        sx1 = min(sx1, sx2)
        if sx1 - fsx1 > 1e-3:
            tab.append((d, sx1 - 1, np.float32((sx1 - fsx1) / cell)))
        for sx in range(sx1, sx2):
            tab.append((d, sx, np.float32(1.0 / cell)))
        if fsx2 - sx2 > 1e-3:
            tab.append((d, sx2, np.float32(min(min(fsx2 - sx2, 1.0), cell) / cell)))
    return tab


_XTAB = area_tab(RAW_W, OUT_W)
_YTAB = area_tab(RAW_H, OUT_H)


def fixed_tap_tables():
    """The per-axis tables padded to 3 taps per output (weight 0 for absent taps).

    Returns (xsrc[84,3] int32, xw[84,3] float32, xn[84] int32, ysrc, yw, yn).  These are what the
    CUDA kernel bakes into constant memory; the number of *real* taps is kept so
    the kernel can skip absent ones (an absent tap must not add +0.0*x, which is
    harmless numerically, but the op ORDER of the present ones must be kept).
    """
    def pack(tab, dsize):
        src = np.zeros((dsize, 3), np.int32)
        w = np.zeros((dsize, 3), np.float32)
        n = np.zeros((dsize,), np.int32)
        for d, s, a in tab:
            src[d, n[d]] = s
            w[d, n[d]] = a
            n[d] += 1
        return src, w, n
    xs, xw, xn = pack(_XTAB, OUT_W)
    ys, yw, yn = pack(_YTAB, OUT_H)
    return xs, xw, xn, ys, yw, yn


def resize_area_84(gray):
    """cv2.resize(gray, (84, 84), interpolation=cv2.INTER_AREA) for a 210x160 uint8 image
    (wrappers.py:32).  fp32, multiply and add are separate roundings (no FMA)."""
    gray = np.asarray(gray)
    assert gray.dtype == np.uint8 and gray.shape[-2:] == (RAW_H, RAW_W)
    lead = gray.shape[:-2]
    src = gray.reshape((-1, RAW_H, RAW_W)).astype(np.float32)
    nimg = src.shape[0]
    # horizontal pass for every source row: buf[dx] = 0; buf[dx] = buf[dx] + S[sx]*alpha (table order)
    buf = np.zeros((nimg, RAW_H, OUT_W), np.float32)
    for d, s, a in _XTAB:
        buf[:, :, d] = buf[:, :, d] + src[:, :, s] * a
    out = np.zeros((nimg, OUT_H, OUT_W), np.float32)
    first = np.ones((OUT_H,), bool)
    for d, s, b in _YTAB:
        term = buf[:, s, :] * b
        if first[d]:
            out[:, d, :] = term
            first[d] = False
        else:
            out[:, d, :] = out[:, d, :] + term
    # saturate_cast<uchar>(float): round half to even (cvRound), clamp
    res = np.clip(np.rint(out), 0, 255).astype(np.uint8)
    return res.reshape(lead + (OUT_H, OUT_W))


def preprocess_frame(rgb):
    """AtariPreprocessFrameWrapper.observation (wrappers.py:30-33): [210,160,3] u8 -> [84,84,1] u8."""
    return resize_area_84(rgb_to_gray(rgb))[..., None]


class FrameStack:
    """FrameStackWrapper (wrappers.py:201-235) for one environment, array part only."""

    def __init__(self, num_stacked_frames=4):
        self.k = num_stacked_frames
        self.stack = np.zeros((OUT_H, OUT_W, num_stacked_frames), np.uint8)

    def step(self, frame, terminal):
        self.stack = np.roll(self.stack, shift=-1, axis=-1)      # :226
        if terminal:
            self.stack.fill(0)                                    # :227-228
        self.stack[..., -1:] = frame                              # :229
        return self.stack

    def reset(self, frame):
        self.stack = np.repeat(frame, self.k, axis=-1)            # :234
        return self.stack


def batched_stack_step(stacks, raw_a, raw_b, terminal, reset_mask=None, reset_raw=None):
    """One MultiEnv.step worth of array work for E environments, as the reference orders it
    (multi_env.py:127-132 then wrappers.py:224-230):

      for each env e:
        if reset_mask[e]:  stack[e] = repeat(preprocess(reset_raw[e]), 4)   # _AutoResetWrapper: previous step was terminal
        frame = preprocess(max(raw_a[e], raw_b[e]))
        stack[e] = roll(stack[e], -1); if terminal[e]: stack[e] = 0; stack[e][..., 3] = frame

    stacks: uint8 [E,84,84,4] (updated copy is returned); raw_*: uint8 [E,210,160,3].
    """
    stacks = np.array(stacks, copy=True)
    e_count = stacks.shape[0]
    if reset_mask is not None:
        for e in range(e_count):
            if reset_mask[e]:
                stacks[e] = np.repeat(preprocess_frame(reset_raw[e]), stacks.shape[-1], axis=-1)
    frames = resize_area_84(rgb_to_gray(np.maximum(raw_a, raw_b)))  # [E,84,84]
    stacks = np.roll(stacks, -1, axis=-1)
    term = np.asarray(terminal, bool)
    stacks[term] = 0
    stacks[..., -1] = frames
    return stacks
